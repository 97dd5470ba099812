"""Host-built vs device-built mesh hierarchies: build time, size, traversal work per ray, throughput (diagnostic)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
grt = importlib.import_module("go-raytracing_b200")
import make_assets
make_assets.ensure_assets()
name = sys.argv[1] if len(sys.argv) > 1 else "cornell-lucy"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
sc = grt.config_scene(name)
depth = sc.cam.max_depth
ref = None
for mode in (0, 1, 1):
    ctx = grt.Context(0)
    ctx.set_option("bvh_device", mode)
    t0 = time.perf_counter(); ctx.load(sc); t_load = time.perf_counter() - t0
    st0 = ctx.stats()
    ctx.set_option("count_stats", 3)
    ctx.clear(); ctx.render_pass(4, depth, seed=5)
    sc_ = ctx.stats()
    rays = sc_["extension_rays"] + sc_["shadow_rays"]
    ctx.set_option("count_stats", 0)
    ctx.clear(); ctx.render_pass(spp, depth, seed=7)
    st = ctx.stats()
    s, _, _ = ctx.resolve_accum()
    if ref is None:
        ref = s
    print(f"[{name}] bvh_device={mode}: load {t_load*1e3:.1f} ms (upload call {st0['ms_scene_upload']:.1f} ms, device build {st0['ms_bvh_build']:.2f} ms, on_device {st0['bvh_on_device']}), "
          f"blas nodes {st0['blas_nodes']} depth {st0['blas_depth']} | nodes/ray {sc_['nodes_visited']/rays:.2f} tris/ray {sc_['tri_tests']/rays:.2f} | "
          f"{st['paths']/st['ms_total']/1e3:.1f} Mpaths/s ext {st['extension_rays']/st['ms_extend']/1e3:.0f} Mr/s conn {st['shadow_rays']/max(st['ms_connect'],1e-9)/1e3:.0f} Mr/s | "
          f"mean radiance {s.mean() / spp:.6f}", flush=True)
    ctx.close()
