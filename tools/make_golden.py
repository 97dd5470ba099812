#!/usr/bin/env python3
"""Makes tests/golden/image_png_region_means.json from the reference's committed image.png
(/root/reference/image.png: hdri-test 800x450, 200 spp, depth 20; rows 420..449 hold the burnt-in stats bar and are
excluded, rt/bucket_renderer.go:383-391). Run in the build container, where /root/reference exists; the JSON is
committed so the test can run on the GPU box. Only 30x30 block means of the linearised pixels ((b+0.5)/256)^2 are stored, not the image."""
import json
import os
import sys

import numpy as np
from PIL import Image

REF = "/root/reference/image.png"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "image_png_region_means.json")
OUT_BAR = os.path.join(os.path.dirname(OUT), "image_png_stats_bar.json")


def main():
    im = np.asarray(Image.open(REF).convert("RGB"), dtype=np.float64)
    im = ((im + 0.5) / 256.0) ** 2  # undo uint8(256*sqrt(x)): back to (clamped) linear radiance per pixel
    H, W, _ = im.shape
    bs = 30
    rows, cols = 420 // bs, W // bs
    means = [[im[by * bs:(by + 1) * bs, bx * bs:(bx + 1) * bs].mean(axis=(0, 1)).round(6).tolist() for bx in range(cols)] for by in range(rows)]
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    json.dump(dict(source="byvfx/go-raytracing image.png (hdri-test 800x450, 200 spp, depth 20)", width=W, height=H, block=bs, rows=rows,
                   cols=cols, means=means), open(OUT, "w"))
    print("wrote", OUT, rows, "x", cols, "blocks")
    # The stats bar of the same image (rows 420..449: drawStatsToFramebuffer, rt/bucket_renderer.go:375-407) is pure black and white:
    # one bit per pixel, the only bit-exact golden the reference tree offers. Pins DrawStatsBar's layout and 7x13 glyph rows.
    raw = np.asarray(Image.open(REF).convert("RGB"))[420:450]
    assert set(np.unique(raw)) <= {0, 255}
    bits = np.packbits((raw[..., 0] == 255).astype(np.uint8), axis=None)
    json.dump(dict(source="byvfx/go-raytracing image.png rows 420..449 (white = 1), np.packbits row-major", width=W, height=H, bar_rows=30,
                   text="800x450 | SPP:200 | Depth:20 | 100.0% | 30.61s | Workers: 32", spp=200, depth=20, seconds=30.61, workers=32,
                   bits_hex=bits.tobytes().hex()), open(OUT_BAR, "w"))
    print("wrote", OUT_BAR)


if __name__ == "__main__":
    sys.exit(main())
