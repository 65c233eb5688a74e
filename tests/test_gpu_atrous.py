"""A-trous level parity (-m gpu): svgf_atrous_host (the CUDA ATrousFilter replacement, through the C ABI with HOST
buffers) against the CPU oracle's restatement of src/denoise.cu:77-170 on seeded synthetic planes
and on ragged / tiny / degenerate sizes, plus size-independent properties at BASELINE.json's full sizes."""
import numpy as np
import pytest

from util import svgf, synthetic_planes, assert_close, COLOR_FLOOR, VAR_FLOOR
import orc

pytestmark = pytest.mark.gpu


def ctx_for(W, H):
    m = svgf()
    blob, R = m.open_scene("cornell", W, H)
    return m, R


@pytest.mark.parametrize("level", [1, 2, 3, 4, 5, 7])
@pytest.mark.parametrize("size", [(256, 256), (333, 77)])
def test_level_matches_oracle(level, size):
    W, H = size
    m, R = ctx_for(W, H)
    color, var, g = synthetic_planes(W, H)
    P = m.default_params()
    co, vo = R.atrous_level(color, var, g, level, False, P)
    oc, ov = orc.atrous_level(color, var, g, level, False, orc.default_params())
    assert_close(co, oc, COLOR_FLOOR, "colour L%d" % level)
    assert_close(vo, ov, VAR_FLOOR, "variance L%d" % level)
    R.close()


@pytest.mark.parametrize("shape", range(15))
def test_every_tile_shape_matches_oracle(shape, monkeypatch):
    """The lattice-tiled kernel is instantiated for several tile shapes / patch heights (csrc/atrous.cu, g_at_shapes);
    launch_atrous picks one per level. Force each of them (SVGF_ATROUS_SHAPE is read at svgf_create) and check it against
    the oracle on a ragged size, at a fine and a coarse level, with the last-level albedo modulation on the coarse one."""
    monkeypatch.setenv("SVGF_ATROUS_SHAPE", str(shape))
    W, H = 334, 141
    m, R = ctx_for(W, H)
    color, var, g = synthetic_planes(W, H, seed=40 + shape)
    for level, last in ((1, False), (4, True)):
        co, vo = R.atrous_level(color, var, g, level, last, m.default_params())
        oc, ov = orc.atrous_level(color, var, g, level, last, orc.default_params())
        assert_close(co, oc, COLOR_FLOOR, "shape %d colour L%d" % (shape, level))
        assert_close(vo, ov, VAR_FLOOR, "shape %d variance L%d" % (shape, level))
    R.close()


@pytest.mark.parametrize("rows", [1, 2])
@pytest.mark.parametrize("forced_shape", [None, 2, 9])
@pytest.mark.parametrize("size", [(256, 256), (333, 77), (31, 5), (1, 1)])
def test_pair_variant_matches_oracle(size, forced_shape, rows, monkeypatch):
    """SVGF_ATROUS_VARIANT=4: the symmetric two-phase kernel (csrc/atrous_pair_core.h; indexing also checked on the CPU by
    tests/test_atrous_emu.py) meets the same bar as the default kernel, both tile shapes, NaN normals included."""
    monkeypatch.setenv("SVGF_ATROUS_VARIANT", "4")
    monkeypatch.setenv("SVGF_ATROUS_PAIR_ROWS", str(rows))
    if forced_shape is not None:
        monkeypatch.setenv("SVGF_ATROUS_SHAPE", str(forced_shape))
    W, H = size
    m, R = ctx_for(W, H)
    color, var, g = synthetic_planes(W, H, seed=77 + W)
    if W > 100:
        g[H // 3:H // 2, W // 4:W // 2, 0:3] = np.nan
    for level, last in ((1, False), (2, False), (4, True), (6, False)):
        co, vo = R.atrous_level(color, var, g, level, last, m.default_params())
        oc, ov = orc.atrous_level(color, var, g, level, last, orc.default_params())
        assert_close(co, oc, COLOR_FLOOR, "pair variant colour %dx%d L%d" % (W, H, level))
        assert_close(vo, ov, VAR_FLOOR, "pair variant variance %dx%d L%d" % (W, H, level))
    R.close()


@pytest.mark.parametrize("shape", [2, 6])
def test_tile_shapes_in_the_frame_path(shape, monkeypatch):
    """Whole frames with a forced tile shape equal the default shape choice bit for bit (same arithmetic per pixel; only
    the tiling differs), odd width included (cp.async loader instead of TMA)."""
    out = []
    for forced in (None, shape):
        if forced is None:
            monkeypatch.delenv("SVGF_ATROUS_SHAPE", raising=False)
        else:
            monkeypatch.setenv("SVGF_ATROUS_SHAPE", str(forced))
        m = svgf()
        blob, R = m.open_scene("cornell", 131, 77)
        P = m.default_params(atrous_nlevel=5)
        drv = blob.camera_driver(131, 77)
        host = np.zeros((77, 131, 3), np.float32)
        for f in range(3):
            R.pathtrace(drv.step(), P, f, host_image=host)
        out.append((host.copy(), R.fetch("variance")))
        R.close()
    assert np.array_equal(out[0][0].view(np.uint32), out[1][0].view(np.uint32))
    assert np.array_equal(out[0][1].view(np.uint32), out[1][1].view(np.uint32))


@pytest.mark.parametrize("case", [("cornell", 640, 360, 5, {}, None, False), ("cornell", 1920, 1080, 5, {}, None, False), ("bunny", 512, 288, 5, {"history_level": 3}, None, True),
                                  ("room", 320, 200, 3, {"blurvariance": 0}, None, False), ("cornell", 256, 256, 1, {}, None, False),
                                  ("cornell", 1280, 720, 5, {}, (300, 471), False), ("cornell", 96, 64, 7, {}, None, False)],
                         ids=lambda c: "%s-%dx%d-L%d%s%s" % (c[0], c[1], c[2], c[3], "-strip" if c[5] else "", "-moving" if c[6] else ""))
def test_stage_in_one_launch_equals_level_by_level(case, monkeypatch):
    """atrous_stage_kernel (SVGF_ATROUS_FUSED=1: all levels of the a-trous stage as one launch of persistent blocks over a
    dependency-ordered work queue, csrc/atrous.cu; measured slower than the default, kept for A/B) runs the per-level kernels' own
    tile code on the same operands: every buffer of every frame must be bit-identical to the level-by-level launches, for tall and
    short lattices (both tile shapes of the
    stage kernel), one to seven levels, a strip of a frame, a moving camera and the history taken from a middle level."""
    scene, W, H, nl, over, strip, moving = case
    out = []
    for fused in ("1", "0"):
        monkeypatch.setenv("SVGF_ATROUS_FUSED", fused)
        m = svgf()
        blob, R = m.open_scene(scene, W, H)
        if strip:
            R.set_shard(0, 1, strip[0], strip[1])
        P = m.default_params(atrous_nlevel=nl, **over)
        drv = blob.camera_driver(W, H, automate=moving)
        got = []
        for f in range(4):
            R.pathtrace(drv.step(), P, f)
            got.append((R.fetch("denoised"), R.fetch("variance"), R.fetch("history_length")))
        out.append(got); R.close()
    r0, r1 = (strip if strip else (0, H))
    for f, (a, b) in enumerate(zip(*out)):
        for x, y, what in zip(a, b, ("denoised", "variance", "history_length")):
            assert np.array_equal(x[r0:r1].view(np.uint32), y[r0:r1].view(np.uint32)), "frame %d: %s differs between the one-launch stage and the per-level launches" % (f, what)


@pytest.mark.parametrize("over", [{}, {"blurvariance": 0}, {"addcolor": 0}, {"sigmal": 2.0, "sigman": 1.0, "sigmax": 1.0},
                                  {"sigmal": 0.01, "sigman": 0.01, "sigmax": 0.01}])
def test_last_level_and_parameter_variants(over):
    W, H = 160, 120
    m, R = ctx_for(W, H)
    color, var, g = synthetic_planes(W, H, seed=7)
    co, vo = R.atrous_level(color, var, g, 3, True, m.default_params(**over))
    oc, ov = orc.atrous_level(color, var, g, 3, True, orc.default_params(**over))
    assert_close(co, oc, COLOR_FLOOR, "colour %r" % over)
    assert_close(vo, ov, VAR_FLOOR, "variance %r" % over)
    R.close()


@pytest.mark.parametrize("size", [(1, 1), (1, 9), (9, 1), (3, 2), (5, 5), (8, 8), (17, 33), (31, 5), (65, 3)])
def test_tiny_and_ragged_sizes(size):
    """Images smaller than the stencil reach (+-2*step): every out-of-image tap is skipped, the 3x3 variance blur
    renormalises at the border (denoise.cu:110-115, 134)."""
    W, H = size
    m, R = ctx_for(W, H)
    color, var, g = synthetic_planes(W, H, seed=W * 100 + H)
    for level in (1, 3, 5):
        co, vo = R.atrous_level(color, var, g, level, level == 5, m.default_params())
        oc, ov = orc.atrous_level(color, var, g, level, level == 5, orc.default_params())
        assert_close(co, oc, COLOR_FLOOR, "colour %dx%d L%d" % (W, H, level))
        assert_close(vo, ov, VAR_FLOOR, "variance %dx%d L%d" % (W, H, level))
    R.close()


def test_zero_variance_and_extreme_inputs():
    """variance == 0 makes the luminance weight exp(-|dl|/1e-6): exercises the fp64 luminance/denominator path."""
    W, H = 96, 64
    m, R = ctx_for(W, H)
    color, var, g = synthetic_planes(W, H, seed=3)
    var[:] = 0.0
    color[:, : W // 2] = 0.25          # flat half: all weights 1
    co, vo = R.atrous_level(color, var, g, 2, False, m.default_params())
    oc, ov = orc.atrous_level(color, var, g, 2, False, orc.default_params())
    assert_close(co, oc, COLOR_FLOOR, "colour var=0")
    assert_close(vo, ov, VAR_FLOOR, "variance var=0")
    var[:] = 100.0                      # "no history" variance (denoise.cu:315)
    co, vo = R.atrous_level(color, var, g, 1, False, m.default_params())
    oc, ov = orc.atrous_level(color, var, g, 1, False, orc.default_params())
    assert_close(co, oc, COLOR_FLOOR, "colour var=100")
    assert_close(vo, ov, 1.0, "variance var=100")
    R.close()


def test_nan_normals_get_weight_one():
    """min(1.0f, expf(-NaN)) = 1 in the reference (CUDA's min drops the NaN, denoise.cu:144-145): pixels whose shading normal
    is NaN (meshes without vertex normals, sceneStructs.h:168-172) are filtered as if all normals agreed; colour and variance
    stay finite and match the oracle."""
    W, H = 96, 72
    m, R = ctx_for(W, H)
    color, var, g = synthetic_planes(W, H, seed=21)
    g[20:50, 30:70, 0:3] = np.nan
    for level in (1, 3):
        co, vo = R.atrous_level(color, var, g, level, level == 3, m.default_params())
        oc, ov = orc.atrous_level(color, var, g, level, level == 3, orc.default_params())
        assert np.isfinite(oc).all() and np.isfinite(co).all() and np.isfinite(vo).all()
        assert_close(co, oc, COLOR_FLOOR, "colour L%d with NaN normals" % level)
        assert_close(vo, ov, VAR_FLOOR, "variance L%d with NaN normals" % level)
    R.close()


@pytest.mark.parametrize("size", [(1920, 1080), (3840, 2160)])
def test_full_size_properties(size):
    """Size-independent properties at BASELINE.json's sizes (the oracle would take minutes here):
    a constant image is a fixed point; the filter is linear in colour for fixed edge-stopping weights only when the
    luminance weight is disabled, so use the partition-of-unity property instead: output lies within the min/max of
    the 5x5 dilated neighbourhood; and two runs are bit-identical."""
    W, H = size
    m, R = ctx_for(W, H)
    color, var, g = synthetic_planes(W, H, seed=11)
    P = m.default_params()
    flat = np.full_like(color, 0.375)
    for level in (1, 5):
        co, vo = R.atrous_level(flat, var, g, level, False, P)
        assert np.abs(co - 0.375).max() < 1e-6, "constant image is not a fixed point at level %d" % level
    co, vo = R.atrous_level(color, var, g, 3, False, P)
    co2, vo2 = R.atrous_level(color, var, g, 3, False, P)
    assert np.array_equal(co.view(np.uint32), co2.view(np.uint32)) and np.array_equal(vo.view(np.uint32), vo2.view(np.uint32))
    assert co.min() >= color.min() - 1e-6 and co.max() <= color.max() + 1e-6
    assert vo.min() >= 0.0 and vo.max() <= var.max() + 1e-6
    # spot-check 64 random rows' worth of pixels against the oracle on a cropped window that contains their stencil
    rng = np.random.default_rng(5)
    step = 8
    for _ in range(4):
        x0 = int(rng.integers(0, W - 200)); y0 = int(rng.integers(0, H - 200))
        sl = (slice(y0, y0 + 200), slice(x0, x0 + 200))
        oc, ov = orc.atrous_level(color[sl], var[sl], g[sl], 3, False, orc.default_params())
        inner = (slice(2 * step + 1, 200 - 2 * step - 1),) * 2
        assert_close(co[sl][inner], oc[inner], COLOR_FLOOR, "window colour")
        assert_close(vo[sl][inner], ov[inner], VAR_FLOOR, "window variance")
    R.close()
