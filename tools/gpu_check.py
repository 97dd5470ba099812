#!/usr/bin/env python3
"""Developer shake-out on a GPU box: level-1 parity on camera rays, a small render per scene, timings.
Not part of the product; the judged checks live in tests/ (-m gpu) and bench.py."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
grt = importlib.import_module("go-raytracing_b200")
import oracle_lib as orc  # noqa: E402


def camera_batch(sc, n, rng):
    ij = np.stack([rng.integers(0, sc.width, n), rng.integers(0, sc.height, n)], axis=1).astype(np.int32)
    sq = rng.random((n, 2)) - 0.5
    r = np.sqrt(rng.random(n))
    a = rng.random(n) * 2 * np.pi
    disk = np.stack([r * np.cos(a), r * np.sin(a)], axis=1)
    tm = rng.random(n)
    return ij, sq, disk, tm


def main():
    names = sys.argv[1:] or ["cornell", "cornell-glossy", "random", "hdri-test", "cornell-lucy"]
    ctx = grt.Context(0)
    for kv in os.environ.get("RTX_OPTS", "").split(","):  # e.g. RTX_OPTS=blas_leaf=2,pool_paths=2097152
        if "=" in kv:
            k, v = kv.split("=")
            ctx.set_option(k, int(v))
    rng = np.random.default_rng(0)
    for name in names:
        w = 1200 if name == "cornell-lucy" else 400
        sc = grt.config_scene(name, width=w)
        t0 = time.time()
        ctx.load(sc)
        t_up = time.time() - t0
        o = orc.OracleScene(sc.desc_ptr, sc.cam_ptr)
        n = 200000
        ij, sq, disk, tm = camera_batch(sc, n, rng)
        rg = ctx.camera_rays(ij, sq, disk, tm)
        ro = o.camera_rays(ij, sq, disk, tm)
        print(f"[{name}] upload {t_up:.2f}s; camera rays max|diff| = {np.abs(rg - ro).max():.3e}")
        t0 = time.time(); hg = ctx.trace_closest(ro); tg = time.time() - t0
        t0 = time.time(); ho = o.trace_closest(ro); to = time.time() - t0
        hit = ho["entry"] >= 0
        bad_id = (hg["entry"] != ho["entry"]) | (hg["prim"] != ho["prim"])
        dt = np.abs(hg["t"] - ho["t"])[hit & ~bad_id]
        dn = np.abs(hg["normal"] - ho["normal"])[hit & ~bad_id]
        dp = np.abs(hg["p"] - ho["p"])[hit & ~bad_id]
        print(f"   trace {n} rays: gpu {tg:.3f}s oracle {to:.3f}s; hits {hit.mean():.3f}; id mismatches {bad_id.sum()}; "
              f"max|dt| {dt.max() if dt.size else 0:.3e} max|dn| {dn.max() if dn.size else 0:.3e} max|dp| {dp.max() if dp.size else 0:.3e} "
              f"front mismatches {(hg['front'] != ho['front'])[~bad_id].sum()}")
        if bad_id.any():
            k = np.flatnonzero(bad_id)[:5]
            for i in k:
                print("     ray", ro[i], "gpu", hg["entry"][i], hg["prim"][i], hg["t"][i], "orc", ho["entry"][i], ho["prim"][i], ho["t"][i])
        # secondary rays: scatter from oracle hit points
        P = ho["p"][hit][:100000]; N = ho["normal"][hit][:100000]
        u = rng.standard_normal(P.shape); u /= np.linalg.norm(u, axis=1, keepdims=True)
        sec = np.concatenate([P, N + u, rng.random((len(P), 1))], axis=1)
        hg2 = ctx.trace_closest(sec); ho2 = o.trace_closest(sec)
        bad2 = (hg2["entry"] != ho2["entry"]) | (hg2["prim"] != ho2["prim"])
        ok2 = (ho2["entry"] >= 0) & ~bad2
        print(f"   secondary {len(sec)} rays: id mismatches {bad2.sum()}; max|dt| {np.abs(hg2['t'] - ho2['t'])[ok2].max() if ok2.any() else 0:.3e}")
        # small render
        spp, depth = 16, sc.cam.max_depth
        ctx.clear(); ctx.enable_moments(True)
        t0 = time.time(); ctx.render_pass(spp, depth, seed=7); tr = time.time() - t0
        st = ctx.stats()
        s, q, cnt = ctx.resolve_accum(moments=True)
        print(f"   render {sc.width}x{sc.height} {spp}spp depth {depth}: {tr:.3f}s wall, {st['ms_total']:.1f} ms device; paths {st['paths']} "
              f"ext {st['extension_rays']} shadow {st['shadow_rays']} iters {st['wavefront_iterations']} "
              f"-> {st['paths'] / st['ms_total'] / 1e3:.2f} Mpaths/s {(st['extension_rays'] + st['shadow_rays']) / st['ms_total'] / 1e3:.1f} Mrays/s; "
              f"ms gen/ext/shade/conn {st['ms_generate']:.1f}/{st['ms_extend']:.1f}/{st['ms_shade']:.1f}/{st['ms_connect']:.1f}; count ok {np.all(cnt == spp)}")
        if sc.width * sc.height <= 400 * 400:
            ro_ = o.render(spp, depth, seed=3, threads=0)
            mg, mo = s.mean(axis=(0, 1)) / spp, ro_["sum"].mean(axis=(0, 1)) / spp
            print(f"   mean radiance gpu {mg} oracle {mo} (oracle {ro_['seconds']:.2f}s, RayCount {ro_['counters']['RayCount']})")
        pix = ctx.resolve_rgba8(spp)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        grt.host().rth_write_png(os.path.join(ROOT, "gpurun_out", f"gpu_{name}.png").encode(), pix.ctypes.data, sc.width, sc.height)
        o.close(); sc.close()


if __name__ == "__main__":
    main()
