#!/usr/bin/env python3
"""Generates the binary assets the configured scenes need (deterministic, numpy only, no network):

  assets/models/lucy_standin.obj   ~280,000-triangle closed, bumpy, statue-proportioned mesh. The reference's
                                   assets/models/lucy_low.obj is a 133-byte Git-LFS pointer, so the real Lucy scan is
                                   not available; bounds follow rt/scenes.go:765
                                   ([-465,-0.025,-267] .. [465,1597,267]). The real file is used instead when
                                   assets/models/lucy_low.obj exists and is > 1 KB.
  assets/hdri/synthetic_hall_1k.hdr  1024x512 Radiance RGBE stand-in (dim hall + bright windows + a small very
                                   bright lamp) used only when the reference's HDRI is not available.
  assets/hdri/abandoned_hall_01_1k.hdr  copied from /root/reference/assets/hdri when that tree exists (it does in
                                   the build container, not on the GPU box; the copy is git-ignored and travels
                                   with the gpurun snapshot exactly like the built .so files).
"""
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_HDR = "/root/reference/assets/hdri/abandoned_hall_01_1k.hdr"
REF_EARTH = "/root/reference/assets/images/earthmap.jpg"


def make_lucy_standin(path, n_theta=400, n_rows=350, seed=1234):
    rng = np.random.default_rng(seed)
    H, XR, ZR = 1597.0, 465.0, 267.0
    rows = np.linspace(0.0, 1.0, n_rows)
    th = np.linspace(0.0, 2 * np.pi, n_theta, endpoint=False)
    # statue-like profile: pedestal, narrow ankles, body, shoulders, head; "wings" widen x at mid-height
    def profile(s):
        ped = 0.55 * np.exp(-((s - 0.02) / 0.05) ** 2)
        body = 0.33 * np.exp(-((s - 0.45) / 0.28) ** 2)
        head = 0.12 * np.exp(-((s - 0.93) / 0.05) ** 2)
        return 0.04 + ped + body + head
    prof = profile(rows)
    wings = 0.75 * np.exp(-((rows - 0.62) / 0.16) ** 2)
    S, T = np.meshgrid(rows, th, indexing="ij")
    rad = np.repeat(prof[:, None], n_theta, axis=1)
    wing = np.repeat(wings[:, None], n_theta, axis=1) * np.abs(np.cos(T)) ** 6
    # bumps: a few octaves of separable sinusoids + per-vertex jitter -> incoherent normals like a scanned statue
    bump = np.zeros_like(rad)
    for k, (ft, fs, amp) in enumerate([(7, 11, 0.035), (19, 29, 0.02), (43, 61, 0.012), (97, 131, 0.006)]):
        ph = rng.uniform(0, 2 * np.pi, size=2)
        bump += amp * np.sin(ft * T + ph[0]) * np.sin(fs * np.pi * S + ph[1])
    bump += 0.004 * rng.standard_normal(rad.shape)
    rx = np.clip(rad + wing * 0.6 + bump, 0.01, None)
    rz = np.clip(rad + bump, 0.01, None)
    x = rx * np.cos(T)
    z = rz * np.sin(T)
    x *= XR / np.abs(x).max()
    z *= ZR / np.abs(z).max()
    y = -0.025 + S * (H + 0.025)
    verts = np.stack([x, y, z], axis=-1).reshape(-1, 3)
    bottom = len(verts)
    top = bottom + 1
    verts = np.vstack([verts, [[0.0, -0.025, 0.0]], [[0.0, H, 0.0]]])
    faces = []
    idx = np.arange(n_rows * n_theta).reshape(n_rows, n_theta)
    a = idx[:-1, :]
    b = np.roll(idx, -1, axis=1)[:-1, :]
    c = idx[1:, :]
    d = np.roll(idx, -1, axis=1)[1:, :]
    faces.append(np.stack([a, c, b], axis=-1).reshape(-1, 3))
    faces.append(np.stack([b, c, d], axis=-1).reshape(-1, 3))
    r0, r1 = idx[0], idx[-1]
    faces.append(np.stack([np.full(n_theta, bottom), r0, np.roll(r0, -1)], axis=-1))
    faces.append(np.stack([np.full(n_theta, top), np.roll(r1, -1), r1], axis=-1))
    faces = np.vstack(faces) + 1
    # shuffle face order a little like a real scan export (strips interleaved), deterministic
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write("# procedural stand-in for lucy_low.obj (see tools/make_assets.py)\n")
        f.write("".join("v %.6f %.6f %.6f\n" % tuple(v) for v in verts))
        f.write("".join("f %d %d %d\n" % tuple(t) for t in faces))
    return len(verts), len(faces)


def float_to_rgbe(rgb):
    rgb = np.maximum(rgb, 0.0)
    m = rgb.max(axis=-1)
    e = np.zeros(m.shape, dtype=np.int32)
    nz = m > 1e-32
    mant, ex = np.frexp(m[nz])
    scale = np.zeros_like(m)
    scale[nz] = mant * 256.0 / m[nz]
    e[nz] = ex + 128
    out = np.zeros(rgb.shape[:-1] + (4,), dtype=np.uint8)
    out[..., :3] = np.clip(rgb * scale[..., None], 0, 255).astype(np.uint8)
    out[..., 3] = np.clip(e, 0, 255).astype(np.uint8)
    out[~nz] = 0
    return out


def make_synthetic_hdr(path, W=1024, H=512, seed=7):
    rng = np.random.default_rng(seed)
    v = (np.arange(H) + 0.5) / H
    u = (np.arange(W) + 0.5) / W
    U, V = np.meshgrid(u, v)
    img = np.zeros((H, W, 3))
    img += np.array([0.55, 0.5, 0.45]) * (0.25 + 0.6 * (1 - V))[..., None]          # walls / ceiling gradient
    img += np.array([0.18, 0.16, 0.14]) * (V > 0.55)[..., None]                        # floor
    for cu in (0.12, 0.37, 0.62, 0.87):                                               # tall windows
        win = (np.abs(U - cu) < 0.035) & (V > 0.18) & (V < 0.5)
        img[win] += np.array([9.0, 10.0, 12.0]) * (0.6 + 0.4 * rng.random())
    lamp = ((U - 0.5) ** 2 * 4 + (V - 0.08) ** 2) < 0.0009                             # small very bright source
    img[lamp] += np.array([70.0, 66.0, 60.0])
    img *= 1.0 + 0.05 * rng.standard_normal((H, W, 1))
    img = np.clip(img, 0.02, None)
    rgbe = float_to_rgbe(img)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n" % (H, W))
        for y in range(H):  # new-RLE scanlines made of literal runs only
            f.write(bytes([2, 2, W >> 8, W & 255]))
            for c in range(4):
                row = rgbe[y, :, c].tobytes()
                for s in range(0, W, 128):
                    chunk = row[s:s + 128]
                    f.write(bytes([len(chunk)]) + chunk)


def write_ppm(path, rgb8):
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (rgb8.shape[1], rgb8.shape[0]))
        f.write(np.ascontiguousarray(rgb8, dtype=np.uint8).tobytes())


def ensure_assets(verbose=False):
    made = []
    obj = os.path.join(ROOT, "assets/models/lucy_standin.obj")
    if not os.path.exists(obj):
        nv, nf = make_lucy_standin(obj)
        made.append(f"{obj}: {nv} vertices, {nf} triangles")
    syn = os.path.join(ROOT, "assets/hdri/synthetic_hall_1k.hdr")
    if not os.path.exists(syn):
        make_synthetic_hdr(syn)
        made.append(syn)
    real = os.path.join(ROOT, "assets/hdri/abandoned_hall_01_1k.hdr")
    if not os.path.exists(real) and os.path.exists(REF_HDR):
        shutil.copyfile(REF_HDR, real)
        made.append(real + " (copied from the reference checkout)")
    # ImageTexture of EarthScene: the host mirror reads binary PPM (no JPEG decoder in the toolchain). The reference's
    # earthmap.jpg is converted when the checkout is present; a synthetic continents-and-oceans map stands in otherwise.
    os.makedirs(os.path.join(ROOT, "assets/images"), exist_ok=True)
    earth = os.path.join(ROOT, "assets/images/earthmap.ppm")
    if not os.path.exists(earth) and os.path.exists(REF_EARTH):
        try:
            from PIL import Image
            im = np.asarray(Image.open(REF_EARTH).convert("RGB"), dtype=np.uint8)
            write_ppm(earth, im)
            made.append(f"{earth} ({im.shape[1]}x{im.shape[0]}, decoded from the reference's earthmap.jpg)")
        except Exception as e:  # noqa: BLE001
            made.append(f"earthmap.jpg not converted: {e}")
    syn_earth = os.path.join(ROOT, "assets/images/synthetic_earth.ppm")
    if not os.path.exists(syn_earth):
        yy, xx = np.mgrid[0:256, 0:512]
        lat, lon = (yy / 256.0 - 0.5) * np.pi, xx / 512.0 * 2 * np.pi
        land = (np.sin(3 * lon + 1.0) * np.cos(2 * lat) + 0.5 * np.sin(7 * lon - 2 * lat) + 0.3 * np.cos(5 * lat + lon)) > 0.25
        im = np.zeros((256, 512, 3), dtype=np.uint8)
        im[...] = (20, 60, 150)
        im[land] = (40, 140, 50)
        im[np.abs(lat) > 1.25] = (235, 240, 245)
        write_ppm(syn_earth, im)
        made.append(syn_earth)
    if verbose:
        for m in made:
            print("made", m)
    return made


if __name__ == "__main__":
    ensure_assets(verbose=True)
    sys.exit(0)
