import importlib, sys, time, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
grt = importlib.import_module("go-raytracing_b200")
sc = grt.config_scene("cornell-lucy")
ctx = grt.Context(0)
for i in range(4):
    t0 = time.perf_counter(); ctx.load(sc); print("load %.2f ms" % ((time.perf_counter() - t0) * 1e3), ctx.stats()["ms_scene_upload"], flush=True)
