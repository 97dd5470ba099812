"""Wall-clock of each C-ABI call of one bench step, host-built vs device-built BVH (diagnostic)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
grt = importlib.import_module("go-raytracing_b200")
import make_assets
make_assets.ensure_assets()
use_torch = len(sys.argv) > 1 and sys.argv[1] == "torch"
if use_torch:
    import torch
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
sc = grt.config_scene("cornell-lucy")
depth = sc.cam.max_depth
for mode in (0, 1):
    ctx = grt.Context(0)
    if use_torch:
        ctx.set_stream(stream.cuda_stream)
    ctx.set_option("bvh_device", mode)
    pix = None
    for rep in range(4):
        t = [time.perf_counter()]
        ctx.load(sc); t.append(time.perf_counter())
        up = ctx.stats()
        ctx.clear(); t.append(time.perf_counter())
        ctx.render_pass(8, depth, seed=rep); t.append(time.perf_counter())
        st = ctx.stats()
        pix = ctx.resolve_rgba8(8, pix); t.append(time.perf_counter())
        d = [(t[i + 1] - t[i]) * 1e3 for i in range(4)]
        print(f"mode {mode} torch {use_torch} rep {rep}: load {d[0]:.1f} (upload {up['ms_scene_upload']:.1f}, bvh {up['ms_bvh_build']:.1f}) clear {d[1]:.1f} render {d[2]:.1f} (device {st['ms_total']:.1f}) resolve {d[3]:.1f} ms", flush=True)
    ctx.close()
