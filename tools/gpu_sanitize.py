"""Small workloads for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family of the library runs once on
a scene small enough for the tools' 10-1000x slowdown. Usage (on the GPU box):
    compute-sanitizer --tool memcheck  python tools/gpu_sanitize.py cornell random lucy20k
    compute-sanitizer --tool racecheck python tools/gpu_sanitize.py lucy20k
`lucy20k` is CornellBoxLucy over a 20,000-triangle stand-in mesh (same generator as the 280K one, coarser grid), built
under a temporary asset root."""
import importlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
grt = importlib.import_module("go-raytracing_b200")
import make_assets  # noqa: E402

make_assets.ensure_assets()


def small_lucy_root():
    root = tempfile.mkdtemp(prefix="rtx_small_")
    make_assets.make_lucy_standin(os.path.join(root, "assets/models/lucy_standin.obj"), n_theta=100, n_rows=100)
    return root


def scene(name, width, spp, depth):
    if name == "lucy20k":
        return grt.NamedScene("cornell-lucy", width, 16.0 / 9.0, spp, depth, asset_root=small_lucy_root())
    return grt.config_scene(name, width=width, spp=spp, depth=depth)


def main():
    names = sys.argv[1:] or ["cornell", "random", "lucy20k"]
    width = int(os.environ.get("SAN_WIDTH", "64"))
    spp = int(os.environ.get("SAN_SPP", "2"))
    nrays = int(os.environ.get("SAN_RAYS", "4096"))
    ctx = grt.Context(0)
    rng = np.random.default_rng(3)
    for name in names:
        depth = 6
        sc = scene(name, width, spp, depth)
        ctx.load(sc)
        n = nrays
        ij = np.stack([rng.integers(0, sc.width, n), rng.integers(0, sc.height, n)], axis=1).astype(np.int32)
        rays = ctx.camera_rays(ij, rng.random((n, 2)) - 0.5, np.zeros((n, 2)), rng.random(n))
        h = ctx.trace_closest(rays)
        for moments in (False, True):
            ctx.enable_moments(moments)
            ctx.clear()
            ctx.render_pass(spp, depth, seed=11)
        ctx.enable_moments(False)
        pix = ctx.resolve_rgba8(spp)
        st = ctx.stats()
        print(f"[sanitize {name}] {sc.width}x{sc.height} {spp} spp depth {depth}: hits {(h['entry'] >= 0).mean():.3f}, "
              f"{st['extension_rays']} ext + {st['shadow_rays']} shadow rays, {st['kernel_launches']} launches, mean byte {pix[..., :3].mean():.1f}", flush=True)
        sc.close()
    ctx.close()
    print("sanitize workload done", flush=True)


if __name__ == "__main__":
    main()
