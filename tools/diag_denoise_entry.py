#!/usr/bin/env python
"""Diagnostic (GPU): is svgf_denoise's output a pure function of its inputs? Runs the same three frames (temporal off / on) in
this process twice: once on fresh memory, once after device memory was filled with NaN bit patterns and freed, and reports where
the outputs differ. A difference means some kernel reads memory nobody initialised."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
m = importlib.import_module("cuda-path-tracer-denoising_b200")
import orc
import torch

def frames(scene, W, H, n, nl, **over):
    sc = orc.Scene(scene); o = orc.Oracle(sc, W, H); P = orc.default_params(atrous_nlevel=nl, **over)
    drv = orc.CameraDriver(sc, W, H, automate=True)
    out = []
    for f in range(n):
        cam = drv.step(); o.frame(cam, P, f, orc.VAR_JACOBI, 0)
        out.append((o.fetch("image").copy(), o.fetch("gbuffer").copy(), cam.as_array().copy()))
    return out

def run(scene, W, H, fr, nl, reset=False, **over):
    blob, R = m.open_scene(scene, W, H)
    if reset:
        R.reset()
    P = m.default_params(atrous_nlevel=nl, **over)
    res = [R.denoise(i, g, m.Camera.from_array(c), P) for i, g, c in fr]
    R.close()
    return res

for over in ({"temporal_enable": 0}, {}):
    scene, W, H, nl = "room", 200, 120, 4
    fr = frames(scene, W, H, 3, nl, **over)
    a = run(scene, W, H, fr, nl, **over)
    junk = torch.full((1 << 28,), float("nan"), device="cuda"); torch.cuda.synchronize(); del junk; torch.cuda.empty_cache()
    b = run(scene, W, H, fr, nl, **over)
    c = run(scene, W, H, fr, nl, reset=True, **over)     # the shim's order: svgf_create, then svgf_reset (denoiseInit)
    for tag, b in (("nan-filled", b), ("reset-first", c)):
        for f in range(3):
            d = (a[f].view(np.uint32) != b[f].view(np.uint32)).any(axis=2)
            rows = np.nonzero(d.any(axis=1))[0]
            print(over, tag, "frame", f, "differing pixels", int(d.sum()), "rows", rows[:5], "...", rows[-5:] if rows.size else "", "nan in b", int(np.isnan(b[f]).sum()))
