"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/rtx_b200.h declares, it refuses
to run without a device (no fallback), and the C++ host mirror of the Go rt API builds / flattens the configured
scenes the way the reference's scene functions describe them (rt/scenes.go)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_header_symbols_exported(grt):
    header = open(os.path.join(ROOT, "include", "rtx_b200.h")).read()
    declared = set(re.findall(r"\b(rtx_[a-z0-9_]+)\s*\(", header))
    assert declared == set(grt.ABI_SYMBOLS), declared ^ set(grt.ABI_SYMBOLS)
    L = grt.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert L.rtx_abi_version() == grt.RTX_ABI_VERSION


def test_struct_layout_matches_header(grt):
    # sizes follow from the field lists in include/rtx_b200.h (LP64): catches a drifting ctypes mirror
    assert C.sizeof(grt.CameraDesc) == 8 + 4 * 3 + 4 + 8 + 9 * 8 + 16 + 48 + 8 + 48 + 8 + 8 + 7 * 24 + 24
    assert C.sizeof(grt.Stats) == 10 * 8 + 5 * 8 + 4 * 4 + 2 * 4 + 2 * 8 + 3 * 8 + 8 + 2 * 4 + 8 + 8 * 8


def test_no_cpu_fallback(grt):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(grt.RtxError) as e:
        grt.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_scene_shapes(grt):
    # object counts of the scene functions (rt/scenes.go)
    c = grt.config_scene("cornell")
    d = c.desc
    assert (c.width, c.height) == (400, 225)
    assert d.n_entries == 9 and d.n_quads == 6 + 3 * 6 and d.n_groups == 3 and d.n_volumes == 1 and d.n_lights == 1
    kinds = [d.entry_geom_kind[i] for i in range(9)]
    assert kinds == [grt.GEOM_QUAD] * 6 + [grt.GEOM_LIST] * 3
    # boxes: Translate outermost, then RotateY (scale 1 is skipped, rt/transform.go:27)
    assert [d.entry_xf_count[i] for i in range(9)] == [0] * 6 + [2, 2, 0]
    xb = d.entry_xf_begin[6]
    assert d.xf_type[xb] == grt.XF_TRANSLATE and d.xf_type[xb + 1] == grt.XF_ROTATE_Y
    assert (d.xf_a[3 * xb], d.xf_a[3 * xb + 1], d.xf_a[3 * xb + 2]) == (265.0, 0.0, 295.0)
    assert d.entry_volume[8] == 0 and d.vol_neg_inv_density[0] == -1.0 / 0.001
    assert d.light_quad[0] == 0 and d.world_is_bvh == 1
    assert (c.cam.samples_per_pixel, c.cam.max_depth, c.cam.vfov) == (10, 10, 40.0)

    g = grt.config_scene("cornell-glossy")
    assert g.desc.n_entries == 10 and g.desc.n_spheres == 4 and g.desc.n_quads == 6 and (g.width, g.height) == (600, 337)
    assert g.desc.light_quad[0] == 5  # the light is the 6th quad added (rt/scenes.go:674-680)

    r = grt.config_scene("random")
    assert r.desc.n_planes == 1 and r.desc.n_spheres == r.desc.n_entries - 1 and 330 < r.desc.n_spheres < 380
    assert r.cam.defocus_angle == 0.6 and r.cam.use_sky_gradient == 1 and r.desc.n_lights == 0
    # seeded: same seed -> identical geometry, different seed -> different
    r2 = grt.config_scene("random")
    n = 3 * r.desc.n_spheres
    assert np.array_equal(np.ctypeslib.as_array(r.desc.sph_center, (n,)), np.ctypeslib.as_array(r2.desc.sph_center, (n,)))
    r3 = grt.config_scene("random", seed=7)
    assert r3.desc.n_spheres != r.desc.n_spheres or not np.array_equal(np.ctypeslib.as_array(r.desc.sph_center, (n,)),
                                                                    np.ctypeslib.as_array(r3.desc.sph_center, (n,)))

    h = grt.config_scene("hdri-test", width=800)
    assert h.desc.n_spheres == 5 and h.desc.n_planes == 1 and h.desc.n_lights == 0      # no AddLight => no NEE
    assert h.desc.env_width == 1024 and h.desc.env_height == 512 and h.cam.phantom_hdri == 1
    assert (h.width, h.height) == (800, 450)


def test_other_scene_shapes(grt):
    # the scene functions beyond BASELINE's five (rt/scenes.go:132-311, :564-604, :820-925)
    q = grt.config_scene("quads")
    assert (q.width, q.height) == (400, 400) and q.desc.n_quads == 5 and q.desc.n_entries == 5 and q.cam.use_sky_gradient == 1 and q.cam.vfov == 80.0
    s_ = grt.config_scene("simple")
    assert s_.desc.n_spheres == 4 and s_.desc.n_planes == 1 and (s_.width, s_.height) == (400, 225)
    iors = sorted(s_.desc.mat_ior[i] for i in range(s_.desc.n_materials) if s_.desc.mat_type[i] == grt.MAT_DIELECTRIC)
    assert iors == [1.0 / 1.5, 1.5]                                     # hollow glass sphere: bubble of ior 1/1.5 inside
    ck = grt.config_scene("checkered")
    assert ck.desc.n_spheres == 2 and any(ck.desc.tex_type[i] == grt.TEX_CHECKER for i in range(ck.desc.n_textures))
    gm = grt.config_scene("glossy-metal")
    assert gm.desc.n_lights == 1 and gm.desc.light_quad[0] == 0 and gm.desc.n_spheres == 3 and (gm.width, gm.height) == (640, 360)
    sm = grt.config_scene("cornell-smoke")
    d = sm.desc
    assert d.n_entries == 8 and d.n_volumes == 2 and d.n_quads == 6 + 12 and d.n_lights == 1
    assert [d.entry_volume[i] for i in range(8)] == [-1] * 6 + [0, 1] and [d.entry_xf_count[i] for i in range(8)] == [0] * 6 + [2, 2]
    assert d.vol_neg_inv_density[0] == -1.0 / 0.01 and (sm.cam.samples_per_pixel, sm.cam.max_depth) == (150, 5)
    pr = grt.config_scene("primitives")
    d = pr.desc
    assert d.n_circles == 1 and d.n_tris == 4 and d.n_entries == 7 and d.n_groups == 2 and d.n_lights == 1 and (pr.width, pr.height) == (800, 450)
    assert [d.entry_geom_kind[i] for i in range(7)] == [grt.GEOM_PLANE, grt.GEOM_CIRCLE, grt.GEOM_LIST, grt.GEOM_SPHERE, grt.GEOM_LIST, grt.GEOM_QUAD, grt.GEOM_SPHERE]
    kinds = [d.list_item_kind[d.group_begin[0] + i] for i in range(d.group_count[0])]
    assert kinds == [grt.GEOM_QUAD] + [grt.GEOM_TRIANGLE] * 4                # Pyramid: base quad + four sides (rt/primitives.go:39-71)
    pe = grt.config_scene("perlin")
    assert pe.desc.n_perlin == 1 and any(pe.desc.tex_type[i] == grt.TEX_NOISE for i in range(pe.desc.n_textures))
    perm = np.ctypeslib.as_array(pe.desc.perlin_perm, (768,))
    assert all(sorted(perm[256 * a:256 * a + 256]) == list(range(256)) for a in range(3))   # three permutations of 0..255
    vec = np.ctypeslib.as_array(pe.desc.perlin_vec, (768,)).reshape(256, 3)
    assert np.allclose((vec * vec).sum(axis=1), 1.0)
    ea = grt.config_scene("earth")                                       # ImageTexture over the PPM conversion of earthmap.jpg (or the stand-in)
    assert ea.desc.n_images == 1 and ea.desc.n_spheres == 1 and ea.desc.image_width[0] >= 512 and (ea.width, ea.height) == (800, 450)
    px = np.ctypeslib.as_array(ea.desc.image_rgb, (3 * ea.desc.image_width[0] * ea.desc.image_height[0],))
    assert 0.0 <= px.min() and px.max() <= 1.0 and px.mean() > 0.2      # sqrt(v / 255): the reference's load-time gamma


def test_lucy_instances(grt):
    s = grt.config_scene("cornell-lucy", width=120, spp=1)
    d = s.desc
    assert d.n_entries == 16 and d.n_groups == 1 and d.group_kind[0] == grt.GEOM_MESH
    assert d.group_count[0] == d.n_tris and d.n_tris > 250000
    inst = [i for i in range(16) if d.entry_geom_kind[i] == grt.GEOM_MESH]
    assert len(inst) == 10 and all(d.entry_geom_index[i] == 0 for i in inst)            # one mesh shared by ten instances
    # Scale -> RotateY -> Translate, RotateY skipped when the angle is 0 (rt/transform.go:24-46, rt/scenes.go:776-790)
    assert [d.entry_xf_count[i] for i in inst] == [3, 3, 3, 3, 2, 3, 3, 3, 2, 3]
    xb = d.entry_xf_begin[inst[0]]
    assert [d.xf_type[xb + k] for k in range(3)] == [grt.XF_TRANSLATE, grt.XF_ROTATE_Y, grt.XF_SCALE]
    assert d.xf_a[3 * (xb + 2)] == 0.15 and d.xf_b[3 * (xb + 2)] == 1.0 / 0.15
    # LoadOBJ defers the reference-order mesh tree (the library derives the test order on the device): tri_rank is NULL ...
    assert not d.tri_rank
    # ... unless the tree is asked for (RT_EAGER_BVH=1 / SetEagerMeshBVH): then tri_rank is a permutation, the DFS leaf order of the tree
    H = grt.host()
    H.rth_set_eager_mesh_bvh(1)
    try:
        s2 = grt.config_scene("cornell-lucy", width=120, spp=1)
        ranks = np.ctypeslib.as_array(s2.desc.tri_rank, (s2.desc.n_tris,))
        assert np.array_equal(np.sort(ranks), np.arange(s2.desc.n_tris))
        assert np.array_equal(np.ctypeslib.as_array(s2.desc.tri_v0, (3 * d.n_tris,)), np.ctypeslib.as_array(d.tri_v0, (3 * d.n_tris,)))
        assert np.array_equal(np.ctypeslib.as_array(s2.desc.tri_v2, (3 * d.n_tris,)), np.ctypeslib.as_array(d.tri_v2, (3 * d.n_tris,)))
    finally:
        H.rth_set_eager_mesh_bvh(0)


def test_unsupported_objects_are_flatten_errors(grt):
    H = grt.host()
    # asset root without the image: EarthScene cannot load its texture -> an error (the reference would render its cyan debug colour)
    assert H.rth_scene_named(b"earth", b"/nonexistent-root", 1, 1, 0, 1.0, 0, 0) is None
    assert b"cannot load" in H.rth_last_error()
    assert H.rth_scene_named(b"no-such-scene", b".", 1, 1, 0, 1.0, 0, 0) is None
    assert b"unknown scene" in H.rth_last_error()


def test_png_writer(grt, tmp_path):
    from PIL import Image
    img = (np.arange(4 * 5 * 7, dtype=np.uint32).reshape(5, 7, 4) * 9 % 256).astype(np.uint8)
    img[..., 3] = 255
    p = str(tmp_path / "t.png")
    assert grt.host().rth_write_png(p.encode(), img.ctypes.data, 7, 5) == 0
    assert np.array_equal(np.asarray(Image.open(p)), img)


def test_stats_bar_matches_reference_image(grt):
    """drawStatsToFramebuffer (rt/bucket_renderer.go:375-407): the bottom 30 rows of the reference's committed image.png are a
    black bar with the stats line in basicfont.Face7x13 — the one bit-exact golden in the reference tree
    (tests/golden/image_png_stats_bar.json, tools/make_golden.py). Layout, text and every glyph used must match bit for bit."""
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "image_png_stats_bar.json")))
    W, H = g["width"], g["height"]
    pix = np.full((H, W, 4), 77, np.uint8)
    text = grt.stats_bar(pix, g["spp"], g["depth"], g["seconds"], g["workers"])
    assert text == g["text"]
    want = np.unpackbits(np.frombuffer(bytes.fromhex(g["bits_hex"]), np.uint8))[:30 * W].reshape(30, W).astype(bool)
    bar = pix[H - 30:]
    assert np.all(bar[..., 3] == 255) and np.all(pix[:H - 30] == 77)            # only the bar is touched
    assert np.array_equal(bar[..., 0] == 255, want) and np.all((bar[..., :3] == 0) | (bar[..., :3] == 255))
    assert np.array_equal(bar[..., 0], bar[..., 1]) and np.array_equal(bar[..., 0], bar[..., 2])
    # FormatDuration (rt/utils.go:50-61) and clipping of images shorter than the bar (image.RGBA.Set ignores outside points)
    small = np.zeros((20, 64, 4), np.uint8)
    assert grt.stats_bar(small, 10, 5, 3725.9, 8) == "64x20 | SPP:10 | Depth:5 | 100.0% | 1h 2m 5s | Workers: 8"
    assert np.all(small[..., 3] == 255)
    assert grt.stats_bar(small, 10, 5, 125.5, 8).split(" | ")[4] == "2m 5s"
    assert grt.stats_bar(small, 10, 5, 0.066, 8).split(" | ")[4] == "0.07s"


# ---- LoadOBJ: the parallel text parse gives what the reference's sequential scan gives (rt/obj_loader.go:15-102) ----
def _write_obj(path, n_blocks, rng, crlf=False, tail=""):
    """Blocks of 4 vertices followed by faces in every index style the reference accepts; returns (vertices, triangles)."""
    eol = "\r\n" if crlf else "\n"
    lines, verts, tris = ["# test mesh", "", "   # indented comment", "o thing", "vn 0 1 0", "vt 0.5 0.5"], [], []
    for b in range(n_blocks):
        base = len(verts)
        for k in range(4):
            # shortest round-trip, fixed, scientific, integer: the short forms take the parser's exact fast path, repr goes to strtod
            fmt = ("{!r}", "{:.6f}", "{:+.4e}", "{:.0f}")[(b + k) % 4]
            toks = [fmt.format(float(c)) for c in rng.normal(size=3) * (10 if k else 1e-3)]
            lines.append(f"v {toks[0]} {toks[1]}\t{toks[2]}" + (" 1.0" if k == 0 else ""))   # a 4th field (w) is ignored
            verts.append(tuple(float(t) for t in toks))
        style = b % 5
        if style == 0:
            lines.append(f"f {base + 1} {base + 2} {base + 3}"); tris.append((base, base + 1, base + 2))
        elif style == 1:   # negative indices count from the vertices read so far
            lines.append("f -4 -3 -2"); tris.append((base, base + 1, base + 2))
        elif style == 2:   # v/vt/vn, quad -> fan of two
            lines.append(f"  f {base + 1}/1/1 {base + 2}/1/1 {base + 3}//1 {base + 4}/1")
            tris += [(base, base + 1, base + 2), (base, base + 2, base + 3)]
        elif style == 3:   # reference to a much earlier vertex, mixed signs
            lines.append(f"f 1 -1 {base + 2}"); tris.append((0, base + 3, base + 1))
        else:              # fewer than three indices: skipped without a word (rt/obj_loader.go:56-58)
            lines.append(f"f {base + 1} {base + 2}")
    with open(path, "w", newline="") as f:
        f.write(eol.join(lines) + eol + tail)
    return np.array(verts), np.array(tris, dtype=np.uint32)


@pytest.mark.parametrize("crlf", [False, True])
def test_obj_parse_parallel_equals_sequential(grt, tmp_path, crlf):
    rng = np.random.default_rng(5)
    path = str(tmp_path / "m.obj")
    verts, tris = _write_obj(path, 12000, rng, crlf=crlf, tail="f 1 2 3")   # > 2 MB of text: several chunks; last line without a newline
    tris = np.vstack([tris, [[0, 1, 2]]]).astype(np.uint32)
    v1, t1, _ = grt.parse_obj(path, threads=1)
    assert np.array_equal(v1, verts) and np.array_equal(t1, tris)          # bit-exact doubles (repr round-trips), file order
    for th in (2, 3, 8):
        v, t, _ = grt.parse_obj(path, threads=th)
        assert np.array_equal(v, v1) and np.array_equal(t, t1), th


def test_obj_parse_errors_are_the_first_by_line(grt, tmp_path):
    rng = np.random.default_rng(6)
    path = str(tmp_path / "e.obj")
    _write_obj(path, 12000, rng)
    text = open(path).read().split("\n")
    n = len(text)

    def run(edit):
        t = list(text)
        for line, s in edit.items():
            t[line - 1] = s
        p = str(tmp_path / "e2.obj")
        open(p, "w").write("\n".join(t))
        msgs = set()
        for th in (1, 8):
            with pytest.raises(RuntimeError) as e:
                grt.parse_obj(p, threads=th)
            msgs.add(str(e.value))
        assert len(msgs) == 1, msgs
        return msgs.pop()

    late, early = n - 100, 200
    assert run({late: "v 1.0 2.0"}) == f"invalid vertex at line {late}"
    assert run({late: "v 1.0 2.0 3.0x"}) == f"invalid vertex coordinates at line {late}"
    assert run({late: "f 1 2 x3"}) == f"invalid face index at line {late}"
    assert run({early: f"f 1 2 {10 ** 6}"}) == f"vertex index out of bounds at line {early}"
    # a face may only use vertices that precede it in the file (the bounds check runs against the vertices read so far)
    assert run({early: "f 1 2 40000"}) == f"vertex index out of bounds at line {early}"
    assert run({early: "f 1 2 -100000"}) == f"vertex index out of bounds at line {early}"
    # two errors in different chunks: the earlier line wins, whichever kind it is
    assert run({early: "f 1 2 40000", late: "v oops 1 2"}) == f"vertex index out of bounds at line {early}"
    assert run({early: "v oops 1 2", late: "f 1 2 40000"}) == f"invalid vertex coordinates at line {early}"
    with pytest.raises(RuntimeError, match="failed to open OBJ file"):
        grt.parse_obj(str(tmp_path / "missing.obj"))


def test_obj_parse_of_the_benchmark_mesh(grt):
    import make_assets
    make_assets.ensure_assets()
    path = os.path.join(ROOT, "assets", "models", "lucy_standin.obj")
    v1, t1, s1 = grt.parse_obj(path, threads=1)
    v8, t8, s8 = grt.parse_obj(path, threads=0)
    assert len(t1) >= 270000 and np.array_equal(v1, v8) and np.array_equal(t1, t8)
    print(f"lucy stand-in: {len(v1)} vertices, {len(t1)} triangles; parse {s1 * 1e3:.0f} ms on 1 thread, {s8 * 1e3:.0f} ms on all")


def test_tile_path_order_visits_every_sample_once():
    """generate_path (csrc/rtx_kernels.cuh), pixel_major = 1: path p of a pass of `spp` samples over `npix` pixels is pixel
    32 (p / 32 spp) + p % 32 at sample (p / 32) % spp, the last npix % 32 pixels handled as one narrower tile. Restated here to pin the
    claim the kernel's comment makes: a bijection onto (pixel, sample), a warp of 32 consecutive paths = 32 neighbouring pixels at one
    sample. (The device code itself is exercised by test_full_size_lucy_properties: per-pixel sample counts on a frame of 810 000 pixels.)"""
    def order(p, npix, spp):
        per_tile, full = 32 * spp, npix // 32
        tile = p // per_tile
        if tile < full:
            within = p - tile * per_tile
            return tile * 32 + (within & 31), within >> 5
        rest = npix - full * 32
        q = p - full * per_tile
        return full * 32 + q % rest, q // rest
    for npix, spp in ((64, 3), (90000, 2), (810000 % 4096 + 4096, 5), (31, 7), (33, 1), (1200 * 3, 37)):
        seen = np.zeros((npix, spp), dtype=np.int32)
        for p in range(npix * spp):
            px, s = order(p, npix, spp)
            seen[px, s] += 1
        assert (seen == 1).all(), (npix, spp)
        if npix >= 64:
            first = [order(p, npix, spp) for p in range(32)]
            assert [px for px, _ in first] == list(range(32)) and {s for _, s in first} == {0}
