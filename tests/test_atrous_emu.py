"""CPU: the a-trous tile kernels' per-thread code -- csrc/atrous_tile_core.h (the production kernel) and
csrc/atrous_pair_core.h (the symmetric two-phase variant): the very functions the CUDA kernels call -- emulated thread by thread
on the host (tests/emu/*.cpp) against the oracle's restatement of ATrousFilter (src/denoise.cu:77-170). Checks what a GPU is
not needed for: the packed pair arithmetic, which taps (and which unordered pair) every centre uses, the aprons, ragged image
borders, strips of a sharded frame, every tile shape of the kernel's table. Bar as for the CUDA kernels: 1e-4 relative."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from util import ROOT, synthetic_planes, assert_close, COLOR_FLOOR, VAR_FLOOR
import orc

CSRC = os.path.join(ROOT, "cuda-path-tracer-denoising_b200", "csrc")


def build_emu(name):
    src, lib = os.path.join(ROOT, "tests", "emu", name + ".cpp"), os.path.join(ROOT, "tests", "emu", "lib" + name + ".so")
    hdrs = [os.path.join(CSRC, "atrous_pair_core.h"), os.path.join(CSRC, "atrous_tile_core.h"), os.path.join(CSRC, "atrous_slide_core.h")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas", src, "-o", lib], check=True)
    L = ctypes.CDLL(lib)
    fn = getattr(L, name + "_level")
    fn.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 6 + [ctypes.c_void_p]
    return fn


@pytest.fixture(scope="module")
def emu():
    return build_emu("pair_emu")


@pytest.fixture(scope="module")
def tile_emu():
    return build_emu("tile_emu")


@pytest.fixture(scope="module")
def slide_emu():
    return build_emu("slide_emu")


def device_planes(color, var, g, P):
    """The planes the CUDA path stages (csrc/svgf_internal.h): cv, lv, the pre-scaled G-buffer view, and the pre-pass's kl."""
    H, W = var.shape
    cv = np.concatenate([color, var[..., None]], axis=2).astype(np.float32)
    lum = (0.2126 * color[..., 0].astype(np.float64) + 0.7152 * color[..., 1] + 0.0722 * color[..., 2]).astype(np.float32)
    lv = np.stack([lum, var], axis=2).astype(np.float32)
    log2e = 1.4426950408889634
    kn = np.float32(log2e / (np.float64(np.float32(P.sigman)) + 1e-6)); kx = np.float32(log2e / (np.float64(np.float32(P.sigmax)) + 1e-6))
    n, p = g[..., 0:3], g[..., 3:6]
    gnp = np.stack([kn * n[..., 0], kx * p[..., 0], kn * n[..., 1], kx * p[..., 1]], axis=2).astype(np.float32)
    gzl = np.stack([kn * n[..., 2], kx * p[..., 2]], axis=2).astype(np.float32)
    # 3x3 Gaussian of the variance, renormalised at the border (denoise.cu:102-115), then kl = log2(e) / (sqrt(v) * sigma_l + 1e-6)
    if P.blurvariance:
        pad = np.pad(var.astype(np.float32), 1)
        ones = np.pad(np.ones_like(var, np.float32), 1)
        acc = np.zeros_like(var, np.float32); wsum = np.zeros_like(var, np.float32)
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                gw = np.float32((0.5 if dx == 0 else 0.25) * (0.5 if dy == 0 else 0.25))
                acc += gw * pad[1 + dy:1 + dy + H, 1 + dx:1 + dx + W]; wsum += gw * ones[1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
        v = np.maximum(acc / wsum, 0)
    else:
        v = np.maximum(var, 0)
    kl = (np.float32(log2e) / (np.sqrt(v, dtype=np.float32) * np.float32(P.sigmal) + np.float32(1e-6))).astype(np.float32)
    return cv, gnp, gzl, lv, kl


def emu_level(L, color, var, g, level, P, shape=0, rows=None):
    H, W = var.shape
    cv, gnp, gzl, lv, kl = (np.ascontiguousarray(a) for a in device_planes(color, var, g, P))
    out = np.full((H, W, 4), np.nan, np.float32)
    r0, r1 = rows if rows else (0, H)
    rc = L(cv.ctypes.data, gnp.ctypes.data, gzl.ctypes.data, lv.ctypes.data, kl.ctypes.data, W, H, 1 << level, r0, r1, shape, out.ctypes.data)
    assert rc == 0
    return out


@pytest.mark.parametrize("shape", [0, 1, 2, 3])
@pytest.mark.parametrize("case", [(96, 80, 1), (96, 80, 2), (61, 45, 1), (61, 45, 3), (50, 70, 4), (33, 17, 5), (5, 3, 1), (1, 1, 2), (130, 40, 7)])
def test_emulated_pair_kernel_matches_oracle(emu, case, shape):
    W, H, level = case
    color, var, g = synthetic_planes(W, H, seed=100 + W + level)
    P = orc.default_params()
    out = emu_level(emu, color, var, g, level, P, shape)
    assert np.isfinite(out).all(), "a pixel was not written, or a pair that phase 1 never produced was read"
    oc, ov = orc.atrous_level(color, var, g, level, False, P)
    assert_close(out[..., 0:3], oc, COLOR_FLOOR, "colour %dx%d L%d" % case)
    assert_close(out[..., 3], ov, VAR_FLOOR, "variance %dx%d L%d" % case)


def test_emulated_strip_and_nan_normals(emu):
    """One rank's strip writes exactly its rows (taps come from the whole frame), and NaN normals get weight 1."""
    W, H, level = 72, 64, 2
    color, var, g = synthetic_planes(W, H, seed=5)
    g[10:30, 20:50, 0:3] = np.nan
    P = orc.default_params()
    oc, ov = orc.atrous_level(color, var, g, level, False, P)
    out = emu_level(emu, color, var, g, level, P, 0, rows=(19, 41))
    assert np.isnan(out[:19]).all() and np.isnan(out[41:]).all() and np.isfinite(out[19:41]).all()
    assert_close(out[19:41, :, 0:3], oc[19:41], COLOR_FLOOR, "strip colour")
    assert_close(out[19:41, :, 3], ov[19:41], VAR_FLOOR, "strip variance")


@pytest.mark.parametrize("shape", list(range(11)) + [13, 14])
@pytest.mark.parametrize("case", [(96, 80, 1), (61, 45, 2), (50, 70, 4), (33, 17, 5), (5, 3, 1), (130, 40, 7)])
def test_emulated_production_kernel_matches_oracle(tile_emu, case, shape):
    W, H, level = case
    color, var, g = synthetic_planes(W, H, seed=300 + W + level)
    if W > 60:
        g[H // 4:H // 2, W // 3:W // 2, 0:3] = np.nan          # NaN normals get weight 1 (denoise.cu:144)
    P = orc.default_params()
    out = emu_level(tile_emu, color, var, g, level, P, shape)
    assert np.isfinite(out).all(), "a pixel was not written"
    oc, ov = orc.atrous_level(color, var, g, level, False, P)
    assert_close(out[..., 0:3], oc, COLOR_FLOOR, "colour %dx%d L%d shape %d" % (case + (shape,)))
    assert_close(out[..., 3], ov, VAR_FLOOR, "variance %dx%d L%d shape %d" % (case + (shape,)))


def test_emulated_production_kernel_strip(tile_emu):
    W, H, level = 72, 64, 3
    color, var, g = synthetic_planes(W, H, seed=9)
    P = orc.default_params(blurvariance=0)
    oc, ov = orc.atrous_level(color, var, g, level, False, P)
    out = emu_level(tile_emu, color, var, g, level, P, 9, rows=(23, 52))
    assert np.isnan(out[:23]).all() and np.isnan(out[52:]).all()
    assert_close(out[23:52, :, 0:3], oc[23:52], COLOR_FLOOR, "strip colour")
    assert_close(out[23:52, :, 3], ov[23:52], VAR_FLOOR, "strip variance")


@pytest.mark.parametrize("bands", [1, 3])
@pytest.mark.parametrize("case", [(96, 80, 1), (61, 45, 2), (50, 70, 4), (33, 17, 5), (5, 3, 1), (1, 1, 2), (130, 40, 7), (200, 36, 3), (26, 90, 1)])
def test_emulated_sliding_kernel_matches_oracle(slide_emu, case, bands):
    """csrc/atrous_slide_core.h: every unordered pair's distance computed by one lane and mirrored to the other (the shuffles are
    emulated by reads of the other lane's registers), five centres in flight per lane through the five-phase register rotation,
    rows staged like the TMA row loads stage them. `bands` is the `shape` argument of the other emulations."""
    W, H, level = case
    color, var, g = synthetic_planes(W, H, seed=500 + W + level)
    if W > 60:
        g[H // 4:H // 2, W // 3:W // 2, 0:3] = np.nan          # NaN normals get weight 1 (denoise.cu:144)
    P = orc.default_params()
    out = emu_level(slide_emu, color, var, g, level, P, bands)
    assert np.isfinite(out).all(), "a pixel was not written"
    oc, ov = orc.atrous_level(color, var, g, level, False, P)
    assert_close(out[..., 0:3], oc, COLOR_FLOOR, "colour %dx%d L%d bands %d" % (case + (bands,)))
    assert_close(out[..., 3], ov, VAR_FLOOR, "variance %dx%d L%d bands %d" % (case + (bands,)))


def test_emulated_sliding_kernel_strip(slide_emu):
    W, H, level = 72, 64, 2
    color, var, g = synthetic_planes(W, H, seed=11)
    P = orc.default_params(blurvariance=0)
    oc, ov = orc.atrous_level(color, var, g, level, False, P)
    out = emu_level(slide_emu, color, var, g, level, P, 2, rows=(23, 52))
    assert np.isnan(out[:23]).all() and np.isnan(out[52:]).all()
    assert_close(out[23:52, :, 0:3], oc[23:52], COLOR_FLOOR, "strip colour")
    assert_close(out[23:52, :, 3], ov[23:52], VAR_FLOOR, "strip variance")
