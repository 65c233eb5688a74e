#!/usr/bin/env python
"""Diagnostic (GPU): tests/test_gpu_denoise_entry.py's room / temporal-off case, piece by piece: the product through
svgf_denoise_host, the shim's denoise() behind the reference's harness (twice: is it deterministic?), and the reference,
with the location and size of every disagreement."""
import importlib, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_denoise_entry as T
m = importlib.import_module("cuda-path-tracer-denoising_b200")

scene, W, H, nl, nframes, over = "room", 200, 120, 4, 3, {"temporal_enable": 0}
tmp = tempfile.mkdtemp()
inp = os.path.join(tmp, "ref.npz")
T._run(T.DUMP, [scene, W, H, nl, nframes, over, inp])
d = np.load(inp)
blob, R = m.open_scene(scene, W, H)
P = m.default_params(atrous_nlevel=nl, **over)
mine = []
for f in range(nframes):
    mine.append(R.denoise(d["f%d_image" % f], d["f%d_gbuffer" % f], m.Camera.from_array(d["f%d_camera" % f]), P).copy())
R.close()
# the same frames in another order on a fresh context: is frame 1 a function of its inputs only?
blob, R = m.open_scene(scene, W, H)
alone = R.denoise(d["f1_image"], d["f1_gbuffer"], m.Camera.from_array(d["f1_camera"]), P).copy()
R.close()
print("product f1 after f0 vs f1 alone: differing", int((mine[1].view(np.uint32) != alone.view(np.uint32)).any(axis=2).sum()))
shim = []
for k in range(2):
    out = os.path.join(tmp, "shim%d.npz" % k)
    T._run(T.ENTRY, ["shim", scene, W, H, nl, nframes, over, inp, out])
    shim.append(np.load(out))
def where(a, b, tag):
    dd = (a.view(np.uint32) != b.view(np.uint32)).any(axis=2)
    ys, xs = np.nonzero(dd)
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-2)
    print(tag, "differing pixels", int(dd.sum()), "rows", (int(ys.min()), int(ys.max())) if ys.size else None, "cols", (int(xs.min()), int(xs.max())) if xs.size else None, "max rel %.3g" % float(rel.max()), "frac > 1e-4: %.4f" % float((rel > 1e-4).mean()))
for f in range(nframes):
    where(shim[0]["f%d_denoised" % f], shim[1]["f%d_denoised" % f], "f%d shim run0 vs run1" % f)
    where(shim[0]["f%d_denoised" % f], mine[f], "f%d shim vs product" % f)
    where(mine[f], d["f%d_denoised" % f], "f%d product vs reference" % f)
    where(shim[0]["f%d_denoised" % f], d["f%d_denoised" % f], "f%d shim vs reference" % f)
