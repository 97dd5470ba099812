"""GPU-vs-oracle mean radiance at increasing depth: localises a coherent bias to a bounce / material (diagnostic)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
grt = importlib.import_module("go-raytracing_b200")
import oracle_lib as orc

name = sys.argv[1] if len(sys.argv) > 1 else "random"
width = int(sys.argv[2]) if len(sys.argv) > 2 else 120
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 512
ctx = grt.Context(0)
for depth in [int(x) for x in (sys.argv[4].split(",") if len(sys.argv) > 4 else "1,2,3,4,8,50".split(","))]:
    sc = grt.config_scene(name, width=width, spp=spp, depth=depth)
    ctx.load(sc)
    o = orc.OracleScene(sc.desc_ptr, sc.cam_ptr)
    ctx.clear(); ctx.enable_moments(True)
    ctx.render_pass(spp, depth, seed=11)
    sg, qg, _ = ctx.resolve_accum(moments=True)
    r = o.render(spp, depth, seed=12, threads=0)
    r2 = o.render(spp, depth, seed=13, threads=0)
    def z(sa, qa, sb, qb):
        ma, mb = sa / spp, sb / spp
        va = np.maximum(qa / spp - ma * ma, 0); vb = np.maximum(qb / spp - mb * mb, 0)
        se = np.sqrt((va + vb) / (spp - 1))
        live = se > 1e-9
        return ((ma - mb)[live] / se[live]).mean(), (ma.mean() / mb.mean() - 1)
    print(f"{name} depth {depth}: gpu-vs-oracle mean z {z(sg.astype(np.float64), qg.astype(np.float64), r['sum'], r['sumsq'])} ; "
          f"oracle-vs-oracle {z(r2['sum'], r2['sumsq'], r['sum'], r['sumsq'])}", flush=True)
    if os.environ.get("BIAS_DUMP"):
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"bias_{name}_d{depth}.npz"), sg=sg, qg=qg, so=r["sum"], qo=r["sumsq"], so2=r2["sum"], qo2=r2["sumsq"], spp=spp)
