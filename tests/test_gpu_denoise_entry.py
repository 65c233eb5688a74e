"""The public denoise(output, input, gbuffer) entry point on its own (-m gpu): /root/reference/src/denoise.h:8,
denoise.cu:349-402 <-> svgf_denoise (device AoS buffers), svgf_denoise_host, and the drop-in shim's denoise().

Inputs are the reference's OWN per-frame `dev_image` and `dev_gbuffer` dumps (its CUDA build, moving camera), fed frame by
frame to
  (1) the product through svgf_denoise_host                       -> must match the reference's denoised / variance /
                                                                     history length of the same frame (exact integers, 1e-4 floats)
  (2) the product through svgf_denoise on caller DEVICE buffers    -> bit-identical to (1)
  (3) the shim's mangled denoise(vec3*, vec3*, GBufferTexel*) behind the reference's harness -> bit-identical to (1)
  (4) the reference's denoise() called directly (not via pathtrace) -> bit-identical to its own pathtrace() path
  (5) the CPU oracle's orc_denoise (pinned bit for bit against the reference's denoise() on the CPU, tests/test_oracle_vs_reference.py)
                                                                   -> 1e-4: same inputs, so no path-tracing noise in between
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from util import ROOT, svgf, assert_close, report, COLOR_FLOOR, VAR_FLOOR
import orc
import refh

pytestmark = pytest.mark.gpu

DUMP = r'''
import sys, json
sys.path.insert(0, %(oracle)r)
import numpy as np, refh
scene, W, H, nl, nframes, over, out = json.loads(sys.argv[1])
h = refh.RefHarness("gpu_jacobi"); h.load_blob(scene, W, H); h.set_params(**refh.ALL_ON); h.set_params(atrous_nlevel=nl); h.set_params(**over)
h.set_params(automate_camera=1, camera_speed_x=0.05, camera_speed_y=0.02, camera_speed_z=0.02, camera_speed_theta=0.02, camera_speed_phi=0.05)
res = {}
for f in range(nframes):
    h.frame()
    for k in ["image", "gbuffer", "camera", "denoised", "variance", "history_length"]:
        res["f%%d_%%s" %% (f, k)] = h.fetch(k)
np.savez(out, **res)
'''

ENTRY = r'''
import sys, json
sys.path.insert(0, %(oracle)r)
import numpy as np, refh
variant, scene, W, H, nl, nframes, over, inp, out = json.loads(sys.argv[1])
h = refh.RefHarness(variant); h.load_blob(scene, W, H); h.set_params(**refh.ALL_ON); h.set_params(atrous_nlevel=nl); h.set_params(**over)
h.init()
d = np.load(inp); res = {}
for f in range(nframes):
    h.set_camera(d["f%%d_camera" %% f])
    res["f%%d_denoised" %% f] = h.denoise(d["f%%d_image" %% f], d["f%%d_gbuffer" %% f])
    res["f%%d_variance" %% f] = h.fetch("variance"); res["f%%d_history_length" %% f] = h.fetch("history_length")
np.savez(out, **res)
'''


def _run(code, args):
    r = subprocess.run([sys.executable, "-c", code % {"oracle": os.path.join(ROOT, "oracle")}, json.dumps(args)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.parametrize("case", [("bunny", 256, 192, 5, 4, {}), ("cornell", 320, 180, 3, 3, {"history_level": 2}), ("room", 200, 120, 4, 3, {"temporal_enable": 0})],
                         ids=lambda c: "%s-%dx%d-%s" % (c[0], c[1], c[2], "-".join("%s%s" % kv for kv in c[5].items()) or "allon"))
def test_denoise_entry_point(case, tmp_path):
    if not (refh.available("gpu_jacobi") and refh.available("shim")):
        pytest.skip("oracle/_ref/libref_gpu_jacobi.so / libref_shim.so not present on this box")
    import ctypes
    import torch
    scene, W, H, nl, nframes, over = case
    m = svgf()
    inp = str(tmp_path / "ref_frames.npz")
    _run(DUMP, [scene, W, H, nl, nframes, over, inp])
    d = np.load(inp)
    what = "denoise() %s %dx%d" % (scene, W, H)
    temporal = over.get("temporal_enable", 1)

    # (1) svgf_denoise_host against the reference's frames, (5) against the oracle on the same inputs
    blob, R = m.open_scene(scene, W, H)
    P = m.default_params(atrous_nlevel=nl, **over)
    osc = orc.Scene(scene); O = orc.Oracle(osc, W, H); OP = orc.default_params(atrous_nlevel=nl, **over)
    mine = {}
    for f in range(nframes):
        cam = m.Camera.from_array(d["f%d_camera" % f])
        out = R.denoise(d["f%d_image" % f], d["f%d_gbuffer" % f], cam, P)
        mine[f] = (out, R.fetch("variance"), R.fetch("history_length"))
        oo = O.denoise(d["f%d_image" % f], d["f%d_gbuffer" % f], orc.Camera.from_array(d["f%d_camera" % f]), OP, orc.VAR_JACOBI, 0)
        if temporal:
            assert np.array_equal(mine[f][2], d["f%d_history_length" % f]), "%s: history length f%d vs the reference" % (what, f)
            # host (no FMA, x86 libm) and device round the reprojected coordinate differently for a handful of pixels
            hd = float((mine[f][2] != O.fetch("history_length")).mean())
            report("%s vs oracle: history length f%d" % (what, f), frac_differ=hd)
            assert hd < 5e-3, "%s: history length f%d differs from the oracle on %.4f of the pixels" % (what, f, hd)
        assert_close(out, d["f%d_denoised" % f], COLOR_FLOOR, "%s vs reference: denoised f%d" % (what, f), max_bad_frac=1e-3)
        assert_close(mine[f][1], d["f%d_variance" % f], VAR_FLOOR, "%s vs reference: variance f%d" % (what, f), max_bad_frac=1e-2)
        # (the few pixels whose reprojection rounds differently spread through the filter's footprint)
        assert_close(out, oo, COLOR_FLOOR, "%s vs oracle: denoised f%d" % (what, f), max_bad_frac=2e-2)
        assert_close(mine[f][1], O.fetch("variance"), VAR_FLOOR, "%s vs oracle: variance f%d" % (what, f), max_bad_frac=5e-2)
    R.close()

    # (2) svgf_denoise on caller-owned DEVICE buffers in the reference's AoS layouts
    blob, R = m.open_scene(scene, W, H)
    for f in range(nframes):
        cam = m.Camera.from_array(d["f%d_camera" % f])
        t_in = torch.from_numpy(d["f%d_image" % f]).cuda(); t_g = torch.from_numpy(d["f%d_gbuffer" % f]).cuda()
        t_out = torch.empty_like(t_in)
        R._ck(m.lib().svgf_denoise(R.h, ctypes.c_void_p(t_out.data_ptr()), ctypes.c_void_p(t_in.data_ptr()), ctypes.c_void_p(t_g.data_ptr()),
                                   ctypes.byref(cam), ctypes.byref(P)), "svgf_denoise")
        assert np.array_equal(t_out.cpu().numpy().view(np.uint32), mine[f][0].view(np.uint32)), "%s: device-buffer entry differs from the host-buffer entry, f%d" % (what, f)
        assert np.array_equal(R.fetch("history_length"), mine[f][2])
    # error behaviour of the entry point: NULL buffers and a wrong resolution are rejected, not dereferenced
    assert m.lib().svgf_denoise(R.h, None, None, None, ctypes.byref(cam), ctypes.byref(P)) != 0
    bad = m.Camera.from_array(d["f0_camera"]); bad.resolution[0] = W + 1
    with pytest.raises(m.SvgfError, match="resolution"):
        R.denoise(d["f0_image"], d["f0_gbuffer"], bad, P)
    R.close()

    # (3) the shim's denoise() behind the reference's harness, (4) the reference's denoise() called directly
    for variant in ("shim", "gpu_jacobi"):
        out = str(tmp_path / ("entry_%s.npz" % variant))
        _run(ENTRY, [variant, scene, W, H, nl, nframes, over, inp, out])
        e = np.load(out)
        for f in range(nframes):
            if variant == "shim":
                assert np.array_equal(e["f%d_denoised" % f].view(np.uint32), mine[f][0].view(np.uint32)), "%s: shim denoise() differs from svgf_denoise_host, f%d" % (what, f)
                assert np.array_equal(e["f%d_variance" % f].view(np.uint32), mine[f][1].view(np.uint32))
                if temporal:
                    assert np.array_equal(e["f%d_history_length" % f], mine[f][2])
            else:
                assert np.array_equal(e["f%d_denoised" % f].view(np.uint32), d["f%d_denoised" % f].view(np.uint32)), "the reference's denoise() called directly differs from its pathtrace() path, f%d" % f
