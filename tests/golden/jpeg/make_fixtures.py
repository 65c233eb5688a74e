#!/usr/bin/env python
"""Regenerates tests/golden/jpeg/*: small JPEG files in the coding modes the decoder supports, and for each the pixels the REFERENCE's
decoder (stb_image, external/include/stb_image.h, called as in src/sceneStructs.h:199: stbi_load(file, &w, &h, &c, 0)) returns.

Needs /root/reference (stb_image is compiled from where it lies, nothing is copied), g++, Pillow. Run in the build container:
    python tests/golden/jpeg/make_fixtures.py
Outputs: <name>.jpg and <name>.rgb (raw bytes, h*w*c) + index.json {name: [w, h, c]}. The test that reads them needs neither.
"""
import json, os, subprocess, sys, tempfile
import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

def picture(w, h, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.stack([127 + 120 * np.sin(x / 5.0 + seed), 127 + 120 * np.cos(y / 7.0), 255 * ((x + y) % 17 > 8)], axis=2)
    img += rng.normal(0, 12, img.shape)           # texture: exercises the AC coefficients
    return Image.fromarray(np.clip(img, 0, 255).astype(np.uint8))

CASES = {   # name: (w, h, PIL save options, grayscale)
    "base444": (40, 24, dict(quality=90, subsampling=0), False),
    "base420_odd": (37, 29, dict(quality=75, subsampling=2), False),           # 4:2:0, size not a multiple of the MCU
    "base422": (33, 16, dict(quality=85, subsampling=1), False),
    "prog444": (24, 40, dict(quality=92, subsampling=0, progressive=True), False),
    "prog420_odd": (45, 31, dict(quality=60, subsampling=2, progressive=True), False),
    "gray": (19, 23, dict(quality=80), True),
    "gray_prog": (32, 32, dict(quality=70, progressive=True), True),
    "optimized_huffman": (48, 20, dict(quality=50, subsampling=2, optimize=True), False),
    "low_quality": (31, 31, dict(quality=5, subsampling=2), False),            # coarse tables, long zero runs
    "one_pixel": (1, 1, dict(quality=90), False),
}

def main():
    helper_src = r'''
#define STB_IMAGE_IMPLEMENTATION
#include <stb_image.h>
#include <cstdio>
int main(int argc, char **argv) {
    int w, h, c; unsigned char *p = stbi_load(argv[1], &w, &h, &c, 0);
    if (!p) { std::fprintf(stderr, "stbi_load failed: %s\n", stbi_failure_reason()); return 1; }
    FILE *f = std::fopen(argv[2], "wb"); std::fwrite(p, 1, (size_t)w * h * c, f); std::fclose(f);
    std::printf("%d %d %d\n", w, h, c); return 0;
}'''
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "stb_decode.cpp"); exe = os.path.join(td, "stb_decode")
        open(src, "w").write(helper_src)
        subprocess.check_call(["g++", "-O2", "-I", os.path.join(REF, "external", "include"), src, "-o", exe])
        index = {}
        for i, (name, (w, h, opts, gray)) in enumerate(sorted(CASES.items())):
            im = picture(w, h, i)
            if gray: im = im.convert("L")
            jpg = os.path.join(HERE, name + ".jpg")
            im.save(jpg, "JPEG", **opts)
            out = subprocess.check_output([exe, jpg, os.path.join(HERE, name + ".rgb")]).split()
            index[name] = [int(v) for v in out]
        # restart markers: Pillow cannot write them; cv2 can
        try:
            import cv2
            im = np.array(picture(50, 34, 99))[:, :, ::-1].copy()
            jpg = os.path.join(HERE, "restart_interval.jpg")
            cv2.imwrite(jpg, im, [cv2.IMWRITE_JPEG_QUALITY, 80, cv2.IMWRITE_JPEG_RST_INTERVAL, 2])
            out = subprocess.check_output([exe, jpg, os.path.join(HERE, "restart_interval.rgb")]).split()
            index["restart_interval"] = [int(v) for v in out]
        except ImportError:
            pass
        json.dump(index, open(os.path.join(HERE, "index.json"), "w"), indent=1, sort_keys=True)
        print(index)

if __name__ == "__main__":
    sys.exit(main())
