"""Experiment: two wavefronts on ONE device (two contexts, two host threads), each rendering half of the samples, against one context
rendering all of them. If the second wavefront fills the SMs the first leaves idle while its launches wind down, the pair is faster."""
import importlib, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
grt = importlib.import_module("go-raytracing_b200")
import make_assets
make_assets.ensure_assets()
name, spp = sys.argv[1], int(sys.argv[2])
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
sc = grt.config_scene(name)
depth = grt.CONFIGS[name]["depth"]
ctxs = [grt.Context(0) for _ in range(lanes)]
for c in ctxs:
    c.load(sc); c.clear()
def one(c, n, base):
    c.render_pass(n, depth, seed=1, sample_base=base)
for rep in range(3):
    t0 = time.perf_counter(); one(ctxs[0], spp, 0); t1 = time.perf_counter()
    th = [threading.Thread(target=one, args=(c, spp // lanes, i * (spp // lanes))) for i, c in enumerate(ctxs)]
    t2 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    t3 = time.perf_counter()
    print(f"[{name} {spp} spp] one context {1e3*(t1-t0):.1f} ms | {lanes} contexts x {spp//lanes} spp concurrently {1e3*(t3-t2):.1f} ms")
