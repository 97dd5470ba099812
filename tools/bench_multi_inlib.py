"""Throughput of the in-library multi-GPU path (rtx_create_multi): ONE process, N devices behind one context, the library slices
the samples and reduces with NCCL. Prints one JSON line per device count. Usage: python tools/bench_multi_inlib.py [workload] [spp]"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
grt = importlib.import_module("go-raytracing_b200")
import make_assets
make_assets.ensure_assets()
name = sys.argv[1] if len(sys.argv) > 1 else "cornell-lucy"
sc = grt.config_scene(name)
spp = int(sys.argv[2]) if len(sys.argv) > 2 else sc.cam.samples_per_pixel
depth = sc.cam.max_depth
ndev = grt.device_count()
base = None
for n in [k for k in (1, 2, 4, 8) if k <= ndev]:
    ctx = grt.Context(devices=list(range(n)))
    ctx.load(sc)
    pix = None
    for rep in range(3):                      # two warm-ups, one timed
        ctx.clear()
        t0 = time.perf_counter()
        ctx.render_pass(spp, depth, seed=7 + rep)
        pix = ctx.resolve_rgba8(spp, pix)
        wall = time.perf_counter() - t0
    st = ctx.stats()
    value = sc.width * sc.height * spp / wall / 1e6
    base = base or value
    print(json.dumps({"impl": "in-library multi-GPU (rtx_create_multi)", "workload": name, "n_devices": n, "spp": spp, "depth": depth,
                      "value": value, "unit": "Mpaths/s", "wall_ms": wall * 1e3, "ms_device_max": st["ms_total"], "ms_reduce": st["ms_reduce"],
                      "ms_resolve": st["ms_resolve"], "ms_tail": st["ms_tail"], "efficiency_vs_1": value / (base * n),
                      "timed": "host wall clock around rtx_render_pass + rtx_resolve_rgba8 (RGBA8 to host), scene resident"}), flush=True)
    ctx.close()
