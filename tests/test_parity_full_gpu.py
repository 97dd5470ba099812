"""GPU parity at BASELINE.json's own sizes, converged bias bounds, and the camera branches no shipped scene takes.

(a) level-2 statistical parity at the exact W x H / spp / depth of the configs that fit a test run — cornell 400x225/10/10,
    random 600x337/100/50, cornell-glossy 600x337/256/5 — and cornell-lucy at its own depth 50 on a reduced frame, against the
    oracle in its no-atomics mode on all host threads (the with-atomics mode computes the same image 15x slower);
(b) a converged-bias test per config on a 96x54 frame: GPU >= 16k spp against oracle >= 4k spp; the global mean of every channel
    within 0.3 % and the 8x8 block means within 3 sigma. This is the north star's "image RMSE under 1 % at matched SPP" in a
    well-posed form: at the configs' own spp two independent renders of the SAME estimator differ by far more than 1 %;
(c) level-1 and level-2 with camera motion (SetMotion) and the free camera (EnableFreeCamera), rt/camera.go:390-434.
"""
import ctypes as C

import numpy as np
import pytest

from test_parity_gpu import assert_level1, camera_batch, check_statistical, secondary_rays, z_scores

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------------------------
# (a) the configs at their own size
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["cornell", "random", "cornell-glossy"])
def test_level2_baseline_config_at_full_size(grt, orc, ctx, name):
    cfg = grt.CONFIGS[name]
    sc = grt.config_scene(name)                     # W, aspect, spp, depth exactly as BASELINE.json / SURVEY 8d
    assert (sc.width, sc.cam.samples_per_pixel, sc.cam.max_depth) == (cfg["width"], cfg["spp"], cfg["depth"])
    ctx.load(sc)
    o = orc.OracleScene(sc.desc_ptr, sc.cam_ptr)
    spp, depth = cfg["spp"], cfg["depth"]
    ctx.clear(); ctx.enable_moments(True)
    ctx.render_pass(spp, depth, seed=21)
    sg, qg, cnt = ctx.resolve_accum(moments=True)
    ctx.enable_moments(False)
    assert np.all(cnt == spp)
    r = o.render(spp, depth, seed=22, threads=0, use_atomics=False)
    mg, mo, se = z_scores(sg.astype(np.float64), qg.astype(np.float64), spp, r["sum"], r["sumsq"], spp)
    live = se > 1e-5 * np.maximum(1e-3, np.abs(mo))
    assert np.allclose(mg[~live], mo[~live], rtol=1e-4, atol=1e-5)
    z = (mg[live] - mo[live]) / se[live]
    # per-pixel: the fraction beyond 3 sigma (Gaussian 0.27 %; heavy-tailed pixels at 10 spp need the slack), no coherent offset
    assert np.mean(np.abs(z) > 3) < (0.03 if spp < 32 else 0.012), np.mean(np.abs(z) > 3)
    assert abs(np.mean(z)) < 0.02, f"mean z = {np.mean(z):.4f}"
    # image level: global mean of every channel inside its own 3.5-sigma band, and that band is narrow at these sizes
    npx = mg.shape[0] * mg.shape[1]
    gm, om = mg.mean(axis=(0, 1)), mo.mean(axis=(0, 1))
    gse = np.sqrt((se ** 2).sum(axis=(0, 1))) / npx
    assert np.all(np.abs(gm - om) <= 3.5 * gse), (name, gm, om, gse)
    # (cornell's 0.9 M heavy-tailed samples give a 1-sigma band of 1.0-1.2 %; the two larger configs are well inside 1 %)
    assert np.all(gse <= (0.015 if name == "cornell" else 0.01) * np.maximum(om, 1e-3)), f"{name}: the global-mean band {gse / om} is too wide"
    # 16x16 block means in noise units
    H, W, _ = mg.shape
    bh, bw = H // 16 * 16, W // 16 * 16
    blk = lambda a: a[:bh, :bw].reshape(bh // 16, 16, bw // 16, 16, 3).mean(axis=(1, 3))
    bse = np.sqrt(blk(se ** 2) / 256)
    ok = bse > 1e-9
    bz = (blk(mg) - blk(mo))[ok] / bse[ok]
    assert np.mean(np.abs(bz) > 3) < 0.012 and np.abs(bz).max() < 6.0, (np.mean(np.abs(bz) > 3), np.abs(bz).max())


def test_level2_lucy_at_depth_50(grt, orc, ctx):
    """cornell-lucy at the config's depth 50 (10 instances of the 280K-triangle mesh, area-light NEE, no Russian roulette) on a
    240x135 frame."""
    sc = grt.config_scene("cornell-lucy", width=240, spp=64, depth=50)
    assert sc.cam.max_depth == grt.CONFIGS["cornell-lucy"]["depth"] == 50
    ctx.load(sc)
    o = orc.OracleScene(sc.desc_ptr, sc.cam_ptr)
    ctx.clear(); ctx.enable_moments(True)
    ctx.render_pass(64, 50, seed=31)
    sg, qg, cnt = ctx.resolve_accum(moments=True)
    ctx.enable_moments(False)
    st = ctx.stats()
    assert st["wavefront_iterations"] >= 50, "paths of a closed box run to the depth limit"
    r = o.render(64, 50, seed=32, threads=0, use_atomics=False)
    mg, mo, se = z_scores(sg.astype(np.float64), qg.astype(np.float64), 64, r["sum"], r["sumsq"], 64)
    live = se > 1e-5 * np.maximum(1e-3, np.abs(mo))
    z = (mg[live] - mo[live]) / se[live]
    assert np.mean(np.abs(z) > 3) < 0.015 and abs(np.mean(z)) < 0.03, (np.mean(np.abs(z) > 3), np.mean(z))
    gm, om = mg.mean(axis=(0, 1)), mo.mean(axis=(0, 1))
    gse = np.sqrt((se ** 2).sum(axis=(0, 1))) / (mg.shape[0] * mg.shape[1])
    assert np.all(np.abs(gm - om) <= 3.5 * gse) and np.all(gse < 0.01 * om), (gm, om, gse)


# ------------------------------------------------------------------------------------------------------------
# (b) converged bias
# ------------------------------------------------------------------------------------------------------------
def _accumulate_gpu(ctx, spp_total, depth, chunk, seed):
    """Sum and sum of squares of `spp_total` samples per pixel, rendered in chunks (the moments buffer holds one chunk)."""
    ctx.clear(); ctx.enable_moments(True)
    done = 0
    while done < spp_total:
        n = min(chunk, spp_total - done)
        ctx.render_pass(n, depth, seed=seed, sample_base=done)
        done += n
    s, q, cnt = ctx.resolve_accum(moments=True)
    ctx.enable_moments(False)
    assert np.all(cnt == spp_total)
    return s.astype(np.float64), q.astype(np.float64)


@pytest.mark.parametrize("name,depth,spp_g,spp_o", [("cornell", 10, 32768, 12288), ("cornell-glossy", 5, 16384, 4096), ("random", 50, 16384, 4096),
                                                   ("hdri-test", 20, 16384, 4096), ("cornell-lucy", 50, 8192, 2048)])
def test_converged_bias_bound(grt, orc, ctx, name, depth, spp_g, spp_o):
    """|global mean difference| <= 0.3 % per channel and 8x8 block means within 3 sigma, GPU >= 16k spp (8k for the mesh scene) against
    the oracle at >= 4k (2k) spp on a 96x54 frame (cornell, whose fog and small light make the heaviest tails: 32k against 12k). A 1 % integrator bias fails this test by more than three of its own sigmas."""
    sc = grt.config_scene(name, width=96, spp=1, depth=depth)
    ctx.load(sc)
    o = orc.OracleScene(sc.desc_ptr, sc.cam_ptr)
    sg, qg = _accumulate_gpu(ctx, spp_g, depth, 4096, seed=77)
    r = o.render(spp_o, depth, seed=78, threads=0, use_atomics=False)
    mg, mo, se = z_scores(sg, qg, spp_g, r["sum"], r["sumsq"], spp_o)
    npx = mg.shape[0] * mg.shape[1]
    gm, om = mg.mean(axis=(0, 1)), mo.mean(axis=(0, 1))
    gse = np.sqrt((se ** 2).sum(axis=(0, 1))) / npx
    rel = np.abs(gm - om) / np.maximum(om, 1e-9)
    print(f"[bias {name}] mean gpu {gm} oracle {om}; relative difference {rel}; 1-sigma band {gse / np.maximum(om, 1e-9)}")
    assert np.all(gse <= 0.0015 * om), f"{name}: the test is not sharp enough: 1-sigma band {gse / om}"
    assert np.all(rel <= 0.003), f"{name}: global mean differs by {rel} (> 0.3 %)"
    H, W, _ = mg.shape
    bh, bw = H // 8 * 8, W // 8 * 8
    blk = lambda a: a[:bh, :bw].reshape(bh // 8, 8, bw // 8, 8, 3).mean(axis=(1, 3))
    bse = np.sqrt(blk(se ** 2) / 64)
    ok = bse > 1e-7 * np.maximum(blk(mo), 1e-3)
    bz = (blk(mg) - blk(mo))[ok] / bse[ok]
    frac3 = np.mean(np.abs(bz) > 3)
    print(f"[bias {name}] {ok.sum()} live block channels: fraction beyond 3 sigma {frac3:.4f}, max |z| {np.abs(bz).max():.2f}, rms z {np.sqrt(np.mean(bz ** 2)):.3f}")
    # 48 blocks x 3 strongly correlated channels: one block beyond 3 sigma shows up as 3 of 144 (2.1 %), which a heavy-tailed scene
    # (cornell: fog, a small bright light) produces in about one run in eight with identical estimators. So: at most one such block, none
    # beyond 5 sigma, and the bulk of the distribution — the median |z| (0.674 for a Gaussian) and the rms — where a common expectation puts it.
    assert frac3 <= 0.022 and np.abs(bz).max() < 5.0, f"{name}: block means outside 3 sigma: {frac3:.4f}, max {np.abs(bz).max():.2f}"
    assert np.median(np.abs(bz)) < 0.85, f"{name}: median |z| of the block means {np.median(np.abs(bz)):.3f} (0.674 = same expectation)"
    assert np.sqrt(np.mean(bz ** 2)) < 1.25, f"{name}: rms block z {np.sqrt(np.mean(bz ** 2)):.3f} (1 = same expectation)"


# ------------------------------------------------------------------------------------------------------------
# (c) camera motion and free camera (rt/camera.go:390-434)
# ------------------------------------------------------------------------------------------------------------
def _cameras(grt, w):
    moving = grt.make_camera(w, 16.0 / 9.0, 32, 8, 40, (278, 278, -800), (278, 278, 0), motion=((340, 300, -760), (250, 260, 30)))
    moving_dof = grt.make_camera(w, 16.0 / 9.0, 32, 8, 40, (278, 278, -800), (278, 278, 0), defocus_angle=1.5, focus_dist=900.0,
                                 motion=((200, 330, -820), (300, 250, -20)))
    free = grt.make_camera(w, 16.0 / 9.0, 32, 8, 55, (278, 320, -700), (0, 0, 0), free_forward=(0.1, -0.08, 1.0))
    free_moving = grt.make_camera(w, 16.0 / 9.0, 32, 8, 55, (278, 320, -700), (0, 0, 0), free_forward=(-0.2, -0.05, 1.0),
                                  motion=((320, 280, -650), (0, 0, 0)))
    return {"motion": moving, "motion+dof": moving_dof, "free": free, "free+motion": free_moving}


@pytest.mark.parametrize("which", ["motion", "motion+dof", "free", "free+motion"])
def test_camera_motion_and_free_camera(grt, orc, ctx, which):
    rng = np.random.default_rng(5)
    sc = grt.config_scene("cornell", width=160)
    cam = _cameras(grt, 160)[which]
    ctx.upload(sc.desc_ptr); ctx.set_camera(cam)
    o = orc.OracleScene(sc.desc_ptr, C.pointer(cam))
    ij, sq, disk, tm = camera_batch(ctx.width, ctx.height, 80000, rng)
    rg, ro = ctx.camera_rays(ij, sq, disk, tm), o.camera_rays(ij, sq, disk, tm)
    assert np.array_equal(rg, ro), f"{which}: camera rays differ (max {np.abs(rg - ro).max()})"
    # independent of the oracle: a moving camera's ray origin is LookFrom + time * (LookFrom2 - LookFrom) (centerMotion.At, rt/camera.go:391)
    # when the lens is a pinhole; the free camera looks along Forward whatever LookAt says
    lf = np.array(list(cam.look_from)); lf2 = np.array(list(cam.look_from2))
    if cam.defocus_angle <= 0:
        want = lf + (tm[:, None] * (lf2 - lf) if cam.camera_motion else 0.0)
        assert np.allclose(rg[:, :3], want, rtol=0, atol=1e-9)
    if cam.free_camera:
        centre = ctx.camera_rays(np.array([[ctx.width // 2, ctx.height // 2]], np.int32), np.zeros((1, 2)), np.zeros((1, 2)), np.array([0.3]))[0]
        d = centre[3:6] / np.linalg.norm(centre[3:6])
        assert np.dot(d, np.array(list(cam.forward))) > 0.999
    assert_level1(ctx.trace_closest(ro), o.trace_closest(ro), f"{which} primary")
    scatter, shadow = secondary_rays(o.trace_closest(ro), rng, 30000)
    assert_level1(ctx.trace_closest(scatter), o.trace_closest(scatter), f"{which} scatter")
    check_statistical(ctx, o, 96, 96, 8, frac_limit=0.025)


def test_camera_motion_in_a_hierarchy_world(grt, orc, ctx):
    """RandomScene (moving spheres, DOF) through a moving camera: the slow GetRay branch feeding the persistent trace kernels."""
    rng = np.random.default_rng(6)
    sc = grt.config_scene("random", width=160)
    cam = grt.make_camera(160, 16.0 / 9.0, 32, 12, 20, (13, 2, 3), (0, 0, 0), defocus_angle=0.6, focus_dist=10.0, sky=True,
                          motion=((12, 2.6, 4.5), (0.5, 0.2, -0.4)))
    ctx.upload(sc.desc_ptr); ctx.set_camera(cam)
    o = orc.OracleScene(sc.desc_ptr, C.pointer(cam))
    ij, sq, disk, tm = camera_batch(ctx.width, ctx.height, 60000, rng)
    ro = o.camera_rays(ij, sq, disk, tm)
    assert np.array_equal(ctx.camera_rays(ij, sq, disk, tm), ro)
    assert_level1(ctx.trace_closest(ro), o.trace_closest(ro), "random, moving camera")
    check_statistical(ctx, o, 64, 64, 12, frac_limit=0.025)
