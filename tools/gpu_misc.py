"""Odds and ends measured on the GPU box: the reference's 3-pass BucketRenderer schedule through the host mirror (image.png's configuration:
hdri-test 800x450, 200 spp, depth 20: 30.61 s in the reference's README stats bar) and the thread scaling of the OBJ text parse."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
grt = importlib.import_module("go-raytracing_b200")
import make_assets
make_assets.ensure_assets()
sc = grt.config_scene("hdri-test", width=800, spp=200, depth=20)
for rep in range(2):
    t0 = time.perf_counter()
    pix, sec = sc.bucket_render(seed=3 + rep)
    print(f"BucketRenderer 3-pass hdri-test {sc.width}x{sc.height} 200 spp depth 20: GetRenderDuration {sec:.3f} s, wall {time.perf_counter() - t0:.3f} s", flush=True)
sc2 = grt.config_scene("cornell-lucy")
for rep in range(2):
    t0 = time.perf_counter()
    pix, sec = sc2.bucket_render(seed=5 + rep)
    print(f"BucketRenderer 3-pass cornell-lucy {sc2.width}x{sc2.height} 500 spp depth 50: GetRenderDuration {sec:.3f} s, wall {time.perf_counter() - t0:.3f} s", flush=True)
path = os.path.join(ROOT, "assets", "models", "lucy_standin.obj")
for th in (1, 2, 4, 8, 16, 0):
    best = min(grt.parse_obj(path, th)[2] for _ in range(3))
    print(f"ParseOBJ lucy_standin.obj (10.6 MB, 280000 triangles) threads={th or 'all'}: {best * 1e3:.1f} ms", flush=True)
os.environ["RT_DEBUG_TIMING"] = "1"
t0 = time.perf_counter(); grt.config_scene("cornell-lucy"); print(f"config_scene cornell-lucy (LoadOBJ + NewBVHNode + flatten): {time.perf_counter() - t0:.3f} s")
