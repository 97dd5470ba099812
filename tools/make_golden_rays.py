#!/usr/bin/env python3
"""Makes tests/golden/level1_<scene>.npz: a fixed batch of rays per scene (seeded camera rays with lens / time jitter, plus
scatter rays leaving the hit points) and the hit records the CPU oracle computes for them: entry id, primitive id, t, front
face. The fixtures are committed; `test_oracle_reproduces_self_generated_golden_rays` (CPU) guards the oracle against drift and
`test_level1_self_generated_golden_fixtures` (GPU) holds the CUDA path to the same records without calling the oracle. Regenerate only when
the oracle is deliberately changed: python tools/make_golden_rays.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
SCENES = {"cornell": 160, "random": 200, "cornell-glossy": 160, "cornell-lucy": 240, "hdri-test": 240, "cornell-smoke": 160, "primitives": 240, "earth": 200}
N_PRIMARY, N_SECONDARY = 1500, 1000


def batch(name, width, grt, orc):
    rng = np.random.default_rng(20261018)
    sc = grt.config_scene(name, width=width, spp=1)
    o = orc.OracleScene(sc.desc_ptr, sc.cam_ptr)
    n = N_PRIMARY
    ij = np.stack([rng.integers(0, sc.width, n), rng.integers(0, sc.height, n)], axis=1).astype(np.int32)
    r, a = np.sqrt(rng.random(n)), rng.random(n) * 2 * np.pi
    rays = o.camera_rays(ij, rng.random((n, 2)) - 0.5, np.stack([r * np.cos(a), r * np.sin(a)], axis=1), rng.random(n))
    h = o.trace_closest(rays)
    hit = np.flatnonzero(h["entry"] >= 0)[:N_SECONDARY]
    u = rng.standard_normal((len(hit), 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    sec = np.concatenate([h["p"][hit], h["normal"][hit] + u, rng.random((len(hit), 1))], axis=1)
    rays = np.concatenate([rays, sec])
    h = o.trace_closest(rays)
    o.close()
    return dict(width=np.int32(width), rays=rays, entry=h["entry"].astype(np.int32), prim=h["prim"].astype(np.int32), t=h["t"], front=h["front"].astype(np.uint8))


def main():
    grt = importlib.import_module("go-raytracing_b200")
    import make_assets
    import oracle_lib as orc
    make_assets.ensure_assets()
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for name, width in SCENES.items():
        b = batch(name, width, grt, orc)
        np.savez_compressed(os.path.join(out, f"level1_{name}.npz"), **b)
        print(name, len(b["rays"]), "rays,", int((b["entry"] >= 0).sum()), "hits")


if __name__ == "__main__":
    sys.exit(main())
