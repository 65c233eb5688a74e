"""Exact identities the CUDA kernels rely on where they replace a reference expression by a cheaper one (CPU, numpy fp32 = IEEE)."""
import numpy as np


def test_normal_distance_threshold_on_squares_is_exact():
    """csrc/denoise.cu, reprj_valid: isReprjValid (src/denoise.cu:172-182) rejects a tap when glm::distance(n_prev, n_cur) > 1e-1f,
    i.e. sqrtf(d2) > 0.1f. The kernel compares d2 with 0x3c23d70b instead. sqrtf is correctly rounded and monotonic, so the two
    agree for every float iff the constant is the largest float whose root is still <= 0.1f: checked on 2^17 neighbours of the
    constant, on a sweep of all exponents and on the specials."""
    T = np.array([0x3c23d70b], dtype=np.uint32).view(np.float32)[0]
    tenth = np.float32(0.1)
    assert np.sqrt(T) <= tenth and np.sqrt(np.nextafter(T, np.float32(1))) > tenth
    near = (np.arange(-(1 << 16), 1 << 16, dtype=np.int64) + 0x3c23d70b).astype(np.uint32).view(np.float32)
    sweep = np.arange(0, 0x7f800000, 9973, dtype=np.uint32).view(np.float32)         # all non-negative finite floats, strided
    special = np.array([0.0, -0.0, np.inf, np.nan, 1e-45, 3.4e38], dtype=np.float32)
    for d2 in (near, sweep, special):
        with np.errstate(invalid="ignore"):
            assert np.array_equal(np.sqrt(d2) > tenth, d2 > T)


def test_division_by_a_weight_sum_of_one_is_the_identity():
    """csrc/denoise.cu, temporal_pixel: the six divides by the bilinear weight sum are skipped when the sum is exactly 1.0f."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal(1 << 16).astype(np.float32) * np.float32(1e3)
    x = np.concatenate([x, np.array([0.0, -0.0, np.inf, -np.inf, 1e-45, 3.4e38], dtype=np.float32)])
    assert np.array_equal((x / np.float32(1.0)).view(np.uint32), x.view(np.uint32))


def test_tap_row_equals_index_over_width_below_2_pow_24_pixels():
    """csrc/denoise.cu, TapOwner: the reference forms a tap's linear index in float, (int)(px + py * W) (src/denoise.cu:176); the
    kernel takes the tap's row as (int)py instead of dividing that index by W. Exact while W * H < 2^24 (every product and sum is
    an integer below 2^24); larger frames keep the division."""
    rng = np.random.default_rng(5)
    for W, H in ((1920, 1080), (3840, 2160), (4095, 4095), (7, 3), (1, 1)):
        assert W * H < (1 << 24)
        px = rng.integers(0, W, 200000).astype(np.float32)
        py = rng.integers(0, H, 200000).astype(np.float32)
        px[:4] = [0, W - 1, 0, W - 1]; py[:4] = [0, 0, H - 1, H - 1]
        q = (px + py * np.float32(W)).astype(np.int32)
        assert np.array_equal(q // W, py.astype(np.int32))
        assert np.array_equal(q, px.astype(np.int64) + py.astype(np.int64) * W)
