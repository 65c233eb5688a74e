"""CPU: the oracle against the committed golden vectors of the REFERENCE'S OWN CUDA BUILD (tests/golden/ref_gpu,
produced on a B200 by oracle/gen_golden_gpu.py). Host and device arithmetic differ (x86 libm and no FMA vs CUDA libdevice
and FMA contraction) and path tracing turns one ulp into a different surface at silhouettes, so this comparison is
statistical; it documents the CPU<->GPU noise floor that tests/test_gpu_parity.py::test_against_cpu_oracle allows for.
(The oracle's exactness is pinned separately, bit for bit, by tests/test_oracle_vs_reference.py.)"""
import os

import numpy as np
import pytest

from util import ROOT, rel_err, COLOR_FLOOR
import orc

GOLD = os.path.join(ROOT, "tests", "golden", "ref_gpu")


@pytest.mark.parametrize("case", [("ref_gpu_jacobi_cornell_64x64.npz", "cornell", 64, 64, 3, [0, 1, 4], False),
                                  ("ref_gpu_jacobi_room_64x64.npz", "room", 64, 64, 5, [0, 2], False),
                                  ("ref_gpu_jacobi_bunny_moving_64x64.npz", "bunny", 64, 64, 5, [0, 3], True)],
                         ids=lambda c: c[0][:-4])
def test_oracle_close_to_reference_gpu_goldens(case):
    fname, scene, W, H, nl, frames, moving = case
    gold = np.load(os.path.join(GOLD, fname))
    sc = orc.Scene(scene); O = orc.Oracle(sc, W, H); P = orc.default_params(atrous_nlevel=nl)
    drv = orc.CameraDriver(sc, W, H, automate=moving)
    for f in range(max(frames) + 1):
        cam = drv.step()
        O.frame(cam, P, f, orc.VAR_JACOBI, 0)
        if f not in frames:
            continue
        assert np.array_equal(gold["f%d_camera" % f].view(np.uint32), cam.as_array().view(np.uint32)), "host camera logic f%d" % f
        g, og = gold["f%d_gbuffer" % f], O.fetch("gbuffer")
        same = g[..., 12].view(np.int32) == og[..., 12].view(np.int32)
        assert same.mean() > 0.985, "geomId agreement %.4f" % same.mean()
        r = rel_err(og[..., :12][same], g[..., :12][same], 1e-2)
        assert (r > 1e-4).mean() < 5e-3
        ri = rel_err(O.fetch("image"), gold["f%d_image" % f], COLOR_FLOOR)
        assert (ri > 1e-4).mean() < 0.02, "1-spp image: %.4f of the values differ" % (ri > 1e-4).mean()
        rd = rel_err(O.fetch("denoised"), gold["f%d_denoised" % f], COLOR_FLOOR)
        assert np.median(rd) < 1e-3 and (rd < 3e-2).mean() > 0.90, "denoised: median %g, within 3e-2: %.4f" % (np.median(rd), (rd < 3e-2).mean())
        assert (O.fetch("history_length") != gold["f%d_history_length" % f]).mean() < 0.05
