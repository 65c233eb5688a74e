"""Shared helpers for the parity tests."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

PKG = "cuda-path-tracer-denoising_b200"


def svgf():
    return importlib.import_module(PKG)


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# ---- tolerances -------------------------------------------------------------------------------------
# north_star: "pixel-for-pixel on object-ID/history-length integers and within 1e-4 relative on colour/variance
# floats". Relative error is taken against max(|reference|, floor): `floor` keeps the measure meaningful for
# values near zero (black pixels; variance = m2 - m1^2 is a cancellation whose absolute error is set by the
# moments' magnitude, not by its own).
REL_TOL = 1e-4
COLOR_FLOOR = 1e-2
VAR_FLOOR = 1e-2


def rel_err(a, b, floor):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def frac_bad(a, b, floor, tol=REL_TOL):
    return float((rel_err(a, b, floor) > tol).mean())


# Every comparison made through assert_close / report() is recorded; tests/conftest.py writes the list to
# gpurun_out/parity_report.json at the end of a session (copied to profiles/ by hand): the MEASURED fractions and maxima,
# not just "below budget".
REPORT = []


def report(what, **kv):
    REPORT.append(dict(what=what, **kv))


def assert_close(a, b, floor, what, tol=REL_TOL, max_bad_frac=0.0):
    r = rel_err(a, b, floor)
    bad = float((r > tol).mean()) if r.size else 0.0
    report(what, n=int(r.size), tol=tol, floor=floor, frac_beyond_tol=bad, budget=max_bad_frac, max_rel_err=float(r.max()) if r.size else 0.0,
           p999_rel_err=float(np.quantile(r, 0.999)) if r.size else 0.0)
    assert bad <= max_bad_frac, "%s: %.4f%% of values exceed %g relative (max %.3g)" % (what, 100 * bad, tol, r.max())


def synthetic_planes(W, H, seed=1234):
    """SURVEY.md 8(d) synthetic a-trous input: noisy colour, U[0,0.25) variance, 4 vertical bands of geometry."""
    rng = np.random.default_rng(seed)
    color = (0.5 + 0.5 * rng.uniform(-1, 1, (H, W, 3))).astype(np.float32)
    variance = rng.uniform(0, 0.25, (H, W)).astype(np.float32)
    band = np.minimum((np.arange(W) * 4) // max(W, 1), 3)
    normals = np.array([[0, 0, 1], [1, 0, 0], [0, 1, 0], [0, 0, -1]], np.float32)
    g = np.zeros((H, W, 13), np.float32)
    g[..., 0:3] = normals[band][None, :, :]
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    g[..., 3] = xs * 0.01; g[..., 4] = ys * 0.01; g[..., 5] = band[None, :].astype(np.float32)
    g[..., 6:9] = rng.uniform(0.3, 1.0, (H, W, 3)).astype(np.float32)
    g[..., 9:12] = 1.0
    g[..., 12] = np.broadcast_to(band[None, :].astype(np.int32), (H, W)).view(np.float32)
    return color, variance, g


def write_scene_blob(scene_file, path, textures=()):
    """Serialise a natively ingested scene (SceneFile) in the blob format the oracle reads (same layout as the blobs the
    reference's loader exported, see cuda-path-tracer-denoising_b200/__init__.py:SceneBlob)."""
    import struct
    m = svgf()
    a = scene_file.arrays()
    hdr = b"SVGFSCN1" + struct.pack("<6i", a["geoms"].size // 248, a["materials"].size // 56, a["triangles"].size // 136,
                                    a["bvh"].size // 40, a["boxes"].shape[0], len(textures)) + struct.pack("<f", scene_file.fovy) + b"\0" * 4
    cam = m.Camera()
    cam.position[:] = list(scene_file.eye); cam.lookAt[:] = list(scene_file.lookat); cam.up[:] = list(scene_file.up)
    out = hdr + bytes(cam) + a["geoms"].tobytes() + a["materials"].tobytes() + a["triangles"].tobytes() + a["bvh"].tobytes() + a["boxes"].tobytes()
    for t in textures:
        t = np.ascontiguousarray(t, np.uint8)
        out += struct.pack("<3i", t.shape[1], t.shape[0], t.shape[2]) + t.tobytes()
    with open(path, "wb") as f:
        f.write(out)
