"""Pins the CPU oracle (oracle/svgf_oracle.cpp) against the REFERENCE'S OWN CODE executed on the CPU.

oracle/_ref/libref_cpu*.so are src/pathtrace.cu + src/denoise.cu of /root/reference compiled by g++ through the
CUDA-on-host shim (oracle/ref/cuda_emu), i.e. the reference's kernels, line for line, run sequentially.
The oracle must reproduce every intermediate BIT FOR BIT (same libm, -ffp-contract=off on both sides):
   libref_cpu.so         <-> oracle variance_mode = INPLACE_SEQ (the in-place variance update in emulator order)
   libref_cpu_jacobi.so  <-> oracle variance_mode = JACOBI      (race-free double buffer)
Each case runs in a subprocess because the reference keeps its state in file statics.
Skipped where the reference binaries are absent (they are built from /root/reference by `make -C oracle ref_cpu`).
"""
import json
import os
import subprocess
import sys

import pytest

import refh

WORKER = r'''
import sys, json
sys.path.insert(0, %(oracle)r)
import numpy as np, refh, orc
variant, mode, scene, W, H, moving, nframes, over = json.loads(sys.argv[1])
h = refh.RefHarness(variant); h.load_blob(scene, W, H); h.set_params(**refh.ALL_ON); h.set_params(**over)
if moving:
    h.set_params(automate_camera=1, camera_speed_x=0.05, camera_speed_y=0.02, camera_speed_z=0.02,
                 camera_speed_theta=0.02, camera_speed_phi=0.05)
sc = orc.Scene(scene); o = orc.Oracle(sc, W, H); P = orc.default_params(**over)
drv = orc.CameraDriver(sc, W, H, automate=bool(moving))
bad = []
for f in range(nframes):
    h.frame()
    cam_ref = h.fetch("camera"); cam = drv.step()
    if not np.array_equal(cam_ref.view(np.uint32), cam.as_array().view(np.uint32)):
        bad.append([f, "camera"])
    o.frame(cam, P, f, mode, 2)
    keys = ["image", "gbuffer", "color_acc", "moment_acc", "history_length", "variance", "color_history", "denoised",
            "pbo", "host_image"]
    if not over.get("temporal_enable", 1) or not over.get("denoise_enable", 1):
        # BackProjection never runs: the reference leaves these cudaMalloc'ed buffers unwritten (denoise.cu:41-53)
        keys = [k for k in keys if k not in ("color_acc", "moment_acc", "history_length")]
    if not over.get("denoise_enable", 1):
        keys = ["image", "gbuffer", "denoised", "pbo", "host_image"]
    for k in keys:
        if not np.array_equal(h.fetch(k).view(np.uint8), o.fetch(k).view(np.uint8)):
            bad.append([f, k])
print(json.dumps(bad))
'''

CASES = [
    ("cpu", 1, "cornell", 64, 64, 0, 5, {}),
    ("cpu_jacobi", 0, "cornell", 64, 64, 0, 5, {}),
    ("cpu_jacobi", 0, "cornell", 96, 64, 1, 4, {}),                       # non-square: aspect reprojection quirk
    ("cpu", 1, "room", 48, 48, 0, 3, {}),                                 # textures + 819-node BVH, 4 mesh geoms
    ("cpu_jacobi", 0, "bunny", 64, 48, 1, 4, {}),                         # moving camera (C5)
    ("cpu_jacobi", 0, "diamond", 48, 48, 0, 3, {}),                       # refraction
    ("cpu_jacobi", 0, "cornell", 40, 40, 0, 3, {"atrous_nlevel": 3, "history_level": 3}),
    ("cpu_jacobi", 0, "cornell", 40, 40, 0, 3, {"temporal_enable": 0}),
    ("cpu_jacobi", 0, "cornell", 40, 40, 0, 3, {"spatial_enable": 0}),
    ("cpu_jacobi", 0, "cornell", 40, 40, 0, 3, {"sepcolor": 0, "addcolor": 0, "blurvariance": 0}),
    ("cpu_jacobi", 0, "cornell", 40, 40, 0, 3, {"denoise_enable": 0}),    # running-mean image (pathtrace.cu:398)
    ("cpu_jacobi", 0, "cornell", 40, 40, 0, 3, {"shadowray": 0, "tracedepth": 6}),
    ("cpu_jacobi", 0, "cornell", 40, 40, 0, 3, {"reducevar": 0, "right_view_option": 2}),
    ("cpu_jacobi", 0, "cornell", 40, 40, 0, 3, {"right_view_option": 1, "atrous_nlevel": 7}),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%s-%dx%d-%s" % (c[0], c[2], c[3], c[4], "-".join("%s%s" % kv for kv in c[7].items()) or "allon"))
def test_oracle_bit_exact_vs_reference_cpu(case):
    if not refh.available(case[0]):
        pytest.skip("oracle/_ref/libref_%s.so not built (needs /root/reference)" % case[0])
    code = WORKER % {"oracle": os.path.dirname(os.path.abspath(refh.__file__))}
    r = subprocess.run([sys.executable, "-c", code, json.dumps(case)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    bad = json.loads(r.stdout.strip().splitlines()[-1])
    assert bad == [], "buffers differing from the reference (frame, name): %r" % bad


DENOISE_WORKER = r'''
import sys, json
sys.path.insert(0, %(oracle)r)
import numpy as np, refh, orc
variant, mode, scene, W, H, nframes, over = json.loads(sys.argv[1])
# inputs: frames of the oracle's own path tracer with a moving camera, then DISTORTED (shuffled rows, scaled colour), so that
# the entry point is exercised on buffers no pathtrace() call would have produced in this order
sc = orc.Scene(scene); src = orc.Oracle(sc, W, H); P = orc.default_params(**over)
drv = orc.CameraDriver(sc, W, H, automate=True)
rng = np.random.default_rng(5)
h = refh.RefHarness(variant); h.load_blob(scene, W, H); h.set_params(**refh.ALL_ON); h.set_params(**over); h.init()
o = orc.Oracle(sc, W, H)
bad = []
for f in range(nframes):
    cam = drv.step()
    src.frame(cam, P, f, mode, 2)
    img = src.fetch("image") * np.float32(1.0 + 0.25 * f); g = src.fetch("gbuffer")
    if f == 2:
        img = img[::-1].copy(); g = g[::-1].copy()
    h.set_camera(cam.as_array())
    a = h.denoise(img, g); b = o.denoise(img, g, cam, P, mode, 2)
    if not np.array_equal(a.view(np.uint8), b.view(np.uint8)):
        bad.append([f, "output"])
    for k in ["variance", "color_history", "moment_history", "history_length", "gbuffer_prev"]:
        if not over.get("temporal_enable", 1) and k in ("moment_history", "history_length"):
            continue
        if not np.array_equal(h.fetch(k).view(np.uint8), o.fetch(k).view(np.uint8)):
            bad.append([f, k])
print(json.dumps(bad))
'''


@pytest.mark.parametrize("case", [("cpu_jacobi", 0, "cornell", 48, 40, 4, {}), ("cpu", 1, "bunny", 40, 40, 3, {"atrous_nlevel": 3}),
                                  ("cpu_jacobi", 0, "cornell", 40, 40, 3, {"temporal_enable": 0}),
                                  ("cpu_jacobi", 0, "room", 40, 40, 3, {"history_level": 2, "sepcolor": 0})],
                         ids=lambda c: "%s-%s-%s" % (c[0], c[2], "-".join("%s%s" % kv for kv in c[6].items()) or "allon"))
def test_oracle_denoise_entry_point_bit_exact(case):
    """The public denoise(output, input, gbuffer) entry (src/denoise.h:8, denoise.cu:349-402) on its own: the oracle's
    orc_denoise against the reference's own denoise() executed on the CPU, on caller-supplied buffers, every history it
    leaves behind included."""
    if not refh.available(case[0]):
        pytest.skip("oracle/_ref/libref_%s.so not built (needs /root/reference)" % case[0])
    code = DENOISE_WORKER % {"oracle": os.path.dirname(os.path.abspath(refh.__file__))}
    r = subprocess.run([sys.executable, "-c", code, json.dumps(case)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    bad = json.loads(r.stdout.strip().splitlines()[-1])
    assert bad == [], "buffers differing from the reference's denoise() (frame, name): %r" % bad
