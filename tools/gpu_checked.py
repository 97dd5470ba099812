"""The checked build's workload: every trace-kernel family on scenes small and large, hit records compared with the production library,
then the violation counters (must be zero) as one JSON line. Run with RTX_B200_LIB pointing at librtx_b200_checked.so."""
import importlib, json, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
grt = importlib.import_module("go-raytracing_b200")
import make_assets
make_assets.ensure_assets()
ctx = grt.Context(0)
rng = np.random.default_rng(11)
out = {"lib": os.path.basename(grt.LIB_PATH), "scenes": {}}
small_root = tempfile.mkdtemp(prefix="rtx_small_")
make_assets.make_lucy_standin(os.path.join(small_root, "assets/models/lucy_standin.obj"), n_theta=100, n_rows=100)
cases = [("cornell", grt.config_scene("cornell", width=160, spp=4, depth=10)), ("random", grt.config_scene("random", width=200, spp=4, depth=20)),
         ("cornell-smoke", grt.config_scene("cornell-smoke", width=120, spp=4, depth=5)), ("primitives", grt.config_scene("primitives", width=160, spp=4, depth=12)),
         ("lucy20k", grt.NamedScene("cornell-lucy", 240, 16.0 / 9.0, 4, 20, asset_root=small_root)), ("cornell-lucy", grt.config_scene("cornell-lucy", width=400, spp=4, depth=50))]
for name, sc in cases:
    for opts in ({}, {"flat_max_entries": 0}, {"tlas_flat_max": 0, "flat_max_entries": 0}, {"lean": 0, "flat_max_entries": 0}):
        for k, v in {"flat_max_entries": 16, "tlas_flat_max": 16, "lean": 1, **opts}.items():
            ctx.set_option(k, v)
        ctx.load(sc)
        n = 30000
        ij = np.stack([rng.integers(0, sc.width, n), rng.integers(0, sc.height, n)], axis=1).astype(np.int32)
        rays = ctx.camera_rays(ij, rng.random((n, 2)) - 0.5, np.zeros((n, 2)), rng.random(n))
        h = ctx.trace_closest(rays)
        ctx.clear()
        ctx.render_pass(sc.cam.samples_per_pixel, sc.cam.max_depth, seed=5)
        st = ctx.stats()
        out["scenes"].setdefault(name, []).append({"opts": opts, "hits": float((h["entry"] >= 0).mean()), "ext": st["extension_rays"], "shadow": st["shadow_rays"],
                                                   "entry_sum": int(h["entry"].astype(np.int64).sum()), "prim_sum": int(h["prim"].astype(np.int64).sum()), "t_sum": float(h["t"].sum())})
st = ctx.stats()
out["checked_build"], out["violations"], out["by_kind"] = st["checked_build"], st["checked_violations"], st["checked_by_kind"]
if st["checked_build"]:
    ctx.set_option("checked_selftest", 5)
    out["selftest_violations"] = ctx.stats()["checked_violations"] - st["checked_violations"]
print(json.dumps(out))
