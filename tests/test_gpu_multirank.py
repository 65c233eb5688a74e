"""Row-strip sharding (-m gpu): several ranks wired in ONE process on one GPU (svgf_peer_connect_local) must reproduce
the unsharded frame BIT FOR BIT -- every rank computes only its rows and reads the other strips' rows (history for the
reprojection, colour/variance/G-buffer aprons for the a-trous levels) in place from their owner, ordered by the
per-stage flags. Bit-exactness holds because the variance update is race-free (Jacobi) and the path tracer's RNG is
keyed by the global pixel index. The multi-process / multi-GPU transport (CUDA IPC handles all-gathered over
torch.distributed) is covered by tests/test_multirank_host.py (gloo, CPU) and by bench.py --gpus N on the box."""
import numpy as np
import pytest

from util import svgf

pytestmark = pytest.mark.gpu

KEYS = ["image", "gbuffer", "history_length", "moment_acc", "color_history", "variance", "denoised", "pbo"]


def render_single(scene, W, H, nl, nframes, moving, **over):
    m = svgf()
    blob, R = m.open_scene(scene, W, H)
    P = m.default_params(atrous_nlevel=nl, **over)
    drv = blob.camera_driver(W, H, automate=moving)
    cams = []
    for f in range(nframes):
        cam = drv.step(); cams.append(m.Camera.from_array(cam.as_array()))
        R.pathtrace(cam, P, f)
    out = {k: R.fetch(k) for k in KEYS}
    R.close()
    return out, cams


@pytest.mark.parametrize("case", [("cornell", 96, 96, 5, 2, False, {}), ("cornell", 160, 120, 5, 3, False, {}),
                                  ("bunny", 128, 96, 5, 4, True, {}), ("room", 96, 64, 3, 8, False, {}),
                                  ("cornell", 64, 64, 5, 4, False, {"history_level": 3}),
                                  ("cornell", 64, 48, 5, 3, False, {"temporal_enable": 0})],
                         ids=lambda c: "%s-%dx%d-w%d" % (c[0], c[1], c[2], c[4]))
def test_sharded_equals_unsharded_bitwise(case):
    scene, W, H, nl, world, moving, over = case
    m = svgf()
    nframes = 4
    ref, cams = render_single(scene, W, H, nl, nframes, moving, **over)
    blob = m.SceneBlob(m.scene_path(scene))
    ranks = [m.Renderer(blob.desc(W, H), W, H) for _ in range(world)]
    rs = m.row_partition(H, world)
    m.connect_local(ranks, rs)
    P = m.default_params(atrous_nlevel=nl, **over)
    for f in range(nframes):
        for R in ranks:                     # asynchronous: each call only queues the rank's frame on its stream
            R.pathtrace(cams[f], P, f)
    for R in ranks:
        R.sync()
        assert R.peer_error() == 0, "a cross-rank wait timed out"
    for k in KEYS:
        parts = [R.fetch(k) for R in ranks]
        got = np.concatenate([parts[r][rs[r]:rs[r + 1]] for r in range(world)], axis=0)
        assert got.shape == ref[k].shape
        assert np.array_equal(got.view(np.uint8), ref[k].view(np.uint8)), "%s differs between %d strips and the whole frame" % (k, world)
    for R in ranks:
        R.close()


def test_uneven_partition_and_empty_strip():
    """Strips need not be equal; a rank may even own zero rows (more ranks than useful work)."""
    m = svgf()
    W, H, nl = 64, 40, 4
    ref, cams = render_single("cornell", W, H, nl, 3, False)
    blob = m.SceneBlob(m.scene_path("cornell"))
    ranks = [m.Renderer(blob.desc(W, H), W, H) for _ in range(4)]
    rs = [0, 3, 3, 29, 40]
    m.connect_local(ranks, rs)
    P = m.default_params(atrous_nlevel=nl)
    for f in range(3):
        for R in ranks:
            R.pathtrace(cams[f], P, f)
    for R in ranks:
        R.sync(); assert R.peer_error() == 0
    for k in ["denoised", "history_length", "variance"]:
        parts = [R.fetch(k) for R in ranks]
        got = np.concatenate([parts[r][rs[r]:rs[r + 1]] for r in range(4)], axis=0)
        assert np.array_equal(got.view(np.uint8), ref[k].view(np.uint8)), k
    for R in ranks:
        R.close()
