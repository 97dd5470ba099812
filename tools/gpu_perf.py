"""Steady-state throughput probe: render `spp` samples of a configured scene, print per-kernel device times (diagnostic)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
grt = importlib.import_module("go-raytracing_b200")
import make_assets
make_assets.ensure_assets()
name = sys.argv[1] if len(sys.argv) > 1 else "cornell-lucy"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ctx = grt.Context(0)
for kv in os.environ.get("RTX_OPTS", "").split(","):
    if "=" in kv:
        k, v = kv.split("="); ctx.set_option(k, int(v))
sc = grt.config_scene(name)
ctx.load(sc)
depth = int(os.environ.get('RTX_DEPTH', sc.cam.max_depth))
for rep in range(2):
    ctx.clear()
    ctx.render_pass(spp, depth, seed=7 + rep)
    st = ctx.stats()
    rays = st["extension_rays"] + st["shadow_rays"]
    print(f"[{name} {sc.width}x{sc.height} {spp}spp d{depth} {os.environ.get('RTX_OPTS','')} {os.path.basename(os.environ.get('RTX_B200_LIB','default'))}] "
          f"{st['ms_total']:.1f} ms: {st['paths']/st['ms_total']/1e3:.1f} Mpaths/s {rays/st['ms_total']/1e3:.0f} Mrays/s | "
          f"ms gen/ext/shade/conn {st['ms_generate']:.1f}/{st['ms_extend']:.1f}/{st['ms_shade']:.1f}/{st['ms_connect']:.1f} | "
          f"tail {st['ms_tail']:.2f} ms / {st['tail_iterations']} it | ext {st['extension_rays']/max(st['ms_extend'],1e-9)/1e3:.0f} Mr/s conn {st['shadow_rays']/max(st['ms_connect'],1e-9)/1e3:.0f} Mr/s iters {st['wavefront_iterations']} nodes/ray {st['nodes_visited']/max(rays,1):.2f} tris/ray {st['tri_tests']/max(rays,1):.2f} quads/ray {st['quad_tests']/max(rays,1):.2f} spheres/ray {st['sphere_tests']/max(rays,1):.3f} planes/ray {st['plane_tests']/max(rays,1):.3f}", flush=True)
