"""The sliding a-trous kernel (-m gpu; SVGF_ATROUS_VARIANT=5, csrc/atrous_slide_core.h) against the CPU oracle's restatement of
src/denoise.cu:77-170 and against the tiled kernel in the frame path. (Its lane program also runs on the CPU in lock-step
emulation, tests/test_atrous_emu.py.)"""
import numpy as np
import pytest

from util import svgf, synthetic_planes, assert_close, COLOR_FLOOR, VAR_FLOOR
import orc

pytestmark = pytest.mark.gpu


def ctx_for(W, H):
    m = svgf()
    blob, R = m.open_scene("cornell", W, H)
    return m, R


@pytest.mark.parametrize("bands", [0, 1, 7])
@pytest.mark.parametrize("size", [(256, 256), (333, 77), (334, 141), (31, 5), (1, 1), (1920, 64)])
def test_sliding_variant_matches_oracle(size, bands, monkeypatch):
    """SVGF_ATROUS_VARIANT=5: the sliding kernel (csrc/atrous_slide_core.h: pair distances once per unordered pair, mirrored by
    register shuffles; per-warp TMA row ring; its lane program is also run in lock step on the CPU, tests/test_atrous_emu.py).
    Same bar as the default kernel: even and odd widths (TMA rows / per-lane loads), band counts, NaN normals, last level."""
    monkeypatch.setenv("SVGF_ATROUS_VARIANT", "5")
    monkeypatch.setenv("SVGF_ATROUS_BANDS", str(bands))
    W, H = size
    m, R = ctx_for(W, H)
    color, var, g = synthetic_planes(W, H, seed=177 + W)
    if W > 100:
        g[H // 3:H // 2, W // 4:W // 2, 0:3] = np.nan
    for level, last in ((1, False), (2, False), (3, False), (4, True), (5, False), (7, False)):
        co, vo = R.atrous_level(color, var, g, level, last, m.default_params())
        oc, ov = orc.atrous_level(color, var, g, level, last, orc.default_params())
        assert_close(co, oc, COLOR_FLOOR, "sliding variant colour %dx%d L%d bands %d" % (W, H, level, bands))
        assert_close(vo, ov, VAR_FLOOR, "sliding variant variance %dx%d L%d bands %d" % (W, H, level, bands))
    R.close()


def test_sliding_variant_in_the_frame_path(monkeypatch):
    """Whole frames through the sliding kernel against the default kernel: same definition, different summation order, so
    within the parity tolerance of each other (not bit-identical); sharded strips included."""
    outs = []
    for variant in ("2", "5"):
        monkeypatch.setenv("SVGF_ATROUS_VARIANT", variant)
        m = svgf()
        blob, R = m.open_scene("cornell", 320, 180)
        P = m.default_params(atrous_nlevel=5)
        drv = blob.camera_driver(320, 180, automate=True)
        host = np.zeros((180, 320, 3), np.float32)
        for f in range(4):
            R.pathtrace(drv.step(), P, f, host_image=host)
        outs.append((host.copy(), R.fetch("variance"), R.fetch("history_length")))
        R.close()
    assert np.array_equal(outs[0][2], outs[1][2])
    assert_close(outs[1][0], outs[0][0], COLOR_FLOOR, "frame through the sliding kernel vs the tiled kernel: denoised", max_bad_frac=1e-3)
    assert_close(outs[1][1], outs[0][1], VAR_FLOOR, "frame through the sliding kernel vs the tiled kernel: variance", max_bad_frac=1e-2)
