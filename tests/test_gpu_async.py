"""Pipelined readback (-m gpu): svgf_render_async + svgf_wait_image deliver, frame for frame, the very bits the blocking
svgf_render(host_image) call delivers (the reference's pathtrace() ends in a blocking D2H of scene->state.image,
src/pathtrace.cu:450), including across a reset and when blocking and pipelined calls are mixed."""
import numpy as np
import pytest

from util import svgf

pytestmark = pytest.mark.gpu


def frames_blocking(scene, W, H, nlevel, n, moving):
    m = svgf()
    blob, R = m.open_scene(scene, W, H)
    P = m.default_params(atrous_nlevel=nlevel)
    drv = blob.camera_driver(W, H, automate=moving)
    host = np.zeros((H, W, 3), np.float32)
    out = []
    for f in range(n):
        R.pathtrace(drv.step(), P, f, host_image=host)
        out.append(host.copy())
    R.close()
    return out


@pytest.mark.parametrize("case", [("cornell", 96, 64, 3, False), ("bunny", 64, 64, 5, True)], ids=lambda c: c[0])
def test_async_equals_blocking(case):
    scene, W, H, nl, moving = case
    n = 7
    want = frames_blocking(scene, W, H, nl, n, moving)
    m = svgf()
    blob, R = m.open_scene(scene, W, H)
    P = m.default_params(atrous_nlevel=nl)
    drv = blob.camera_driver(W, H, automate=moving)
    bufs = [np.zeros((H, W, 3), np.float32) for _ in range(2)]
    got = []
    for f in range(n):
        R.pathtrace_async(drv.step(), P, f, bufs[f & 1])
        if f >= 1:      # consume the previous frame while this one renders
            R.wait_image(bufs[(f - 1) & 1])
            got.append(bufs[(f - 1) & 1].copy())
    R.wait_image(bufs[(n - 1) & 1])
    got.append(bufs[(n - 1) & 1].copy())
    for f in range(n):
        assert np.array_equal(got[f].view(np.uint32), want[f].view(np.uint32)), "frame %d differs" % f
    # the introspection view follows the buffer rotation
    assert np.array_equal(R.fetch("denoised").view(np.uint32), want[-1].view(np.uint32))
    R.close()


def test_mixed_blocking_and_async_and_reset():
    scene, W, H, nl = "cornell", 64, 48, 3
    want = frames_blocking(scene, W, H, nl, 6, False)
    m = svgf()
    blob, R = m.open_scene(scene, W, H)
    P = m.default_params(atrous_nlevel=nl)
    bufs = [np.zeros((H, W, 3), np.float32) for _ in range(3)]
    for rnd in range(2):        # second round: after a reset the same six frames come out again
        drv = blob.camera_driver(W, H)
        got = []
        for f in range(6):
            cam = drv.step()
            if f in (2, 5):     # a blocking call in between must not overtake the copies still in flight
                R.pathtrace(cam, P, f, host_image=bufs[2])
                R.wait_image(None)
                got.extend(b.copy() for b in pending)
                pending = []
                got.append(bufs[2].copy())
            else:
                if f in (0, 3):
                    pending = []
                R.pathtrace_async(cam, P, f, bufs[len(pending)])
                pending.append(bufs[len(pending)])
        for f in range(6):
            assert np.array_equal(got[f].view(np.uint32), want[f].view(np.uint32)), "round %d frame %d differs" % (rnd, f)
        R.reset()
    R.close()


@pytest.mark.parametrize("scene,moving", [("cornell", False), ("bunny", True)])
def test_cuda_graph_frames_are_bit_identical(scene, moving):
    """with the "cuda_graph" option the frame's launches are captured, the instantiated graph is updated in
    place and launched as one unit. Same kernels, same arguments: every buffer bit-identical to the plain launches, across a
    parameter change (another kernel sequence: the graph is re-instantiated) and a reset."""
    m = svgf()
    W, H = 160, 96
    outs = []
    for on in (0, 1):
        blob, R = m.open_scene(scene, W, H)
        R.set_option("cuda_graph", on)
        drv = blob.camera_driver(W, H, automate=moving)
        host = np.zeros((H, W, 3), np.float32)
        got = []
        for f in range(9):
            P = m.default_params(atrous_nlevel=5 if f < 5 else 3, history_level=1 if f < 7 else 2)
            if f == 4:
                R.pathtrace_async(drv.step(), P, f, host); R.wait_image(host)
            else:
                R.pathtrace(drv.step(), P, f, host_image=host)
            got.append((host.copy(), R.fetch("variance"), R.fetch("history_length"), R.fetch("pbo")))
        R.reset()
        P = m.default_params()
        drv = blob.camera_driver(W, H, automate=moving)
        for f in range(4):
            R.pathtrace(drv.step(), P, f, host_image=host)
        got.append((host.copy(), R.fetch("variance"), R.fetch("history_length"), R.fetch("pbo")))
        outs.append(got); R.close()
    for f, (a, b) in enumerate(zip(*outs)):
        for x, y, what in zip(a, b, ("image", "variance", "history_length", "pbo")):
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), "frame %d: %s differs with the CUDA graph" % (f, what)


@pytest.mark.parametrize("case", [("cornell", 640, 360, 5, False, {}), ("bunny", 512, 288, 5, True, {}), ("room", 320, 200, 3, True, {"temporal_enable": 0}),
                                  ("cornell", 256, 256, 5, False, {"toggle": 1})], ids=lambda c: "%s-%dx%d%s" % (c[0], c[1], c[2], "-".join([""] + list(c[5]))))
def test_cross_frame_overlap_is_bit_identical(case, monkeypatch):
    """Cross-frame overlap (csrc/api.cu: frame_body): on a single GPU the path tracer of frame N + 1 runs on its own stream next to
    the a-trous stage of frame N, on a second set of the buffers both touch. Every frame's image, and the state the last frame
    leaves, must be the bits of the serial schedule (SVGF_FRAME_OVERLAP=0) -- pipelined calls, frames queued back to back without
    any synchronisation, a parameter change in the middle (denoiser off and on again: the running-mean path does not overlap),
    a reset."""
    scene, W, H, nl, moving, over = case
    toggle = over.pop("toggle", 0) if isinstance(over, dict) else 0
    over = {k: v for k, v in over.items() if k != "toggle"}
    outs = []
    for on in ("1", "0"):
        monkeypatch.setenv("SVGF_FRAME_OVERLAP", on)
        m = svgf()
        blob, R = m.open_scene(scene, W, H)
        drv = blob.camera_driver(W, H, automate=moving)
        bufs = [np.zeros((H, W, 3), np.float32) for _ in range(2)]
        got = []
        n = 9
        for f in range(n):
            P = m.default_params(atrous_nlevel=nl, **over)
            if toggle and f in (4, 5):
                P = m.default_params(atrous_nlevel=nl, denoise_enable=0)
            R.pathtrace_async(drv.step(), P, f, bufs[f & 1])
            if f >= 1:
                R.wait_image(bufs[(f - 1) & 1]); got.append(bufs[(f - 1) & 1].copy())
        R.wait_image(bufs[(n - 1) & 1]); got.append(bufs[(n - 1) & 1].copy())
        # frames queued back to back, nothing waited for until the end
        P = m.default_params(atrous_nlevel=nl, **over)
        for f in range(n, n + 6):
            R.pathtrace(drv.step(), P, f)
        got.append(R.fetch("denoised")); got.append(R.fetch("variance")); got.append(R.fetch("history_length")); got.append(R.fetch("image")); got.append(R.fetch("gbuffer"))
        R.reset()
        drv = blob.camera_driver(W, H, automate=moving)
        for f in range(3):
            R.pathtrace(drv.step(), P, f)
        got.append(R.fetch("denoised")); got.append(R.fetch("pbo"))
        outs.append(got); R.close()
    assert len(outs[0]) == len(outs[1])
    for i, (a, b) in enumerate(zip(*outs)):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), "output %d differs between the overlapped and the serial schedule" % i


def test_overlapped_frames_wait_for_a_bvh_refit(monkeypatch):
    """With frame_overlap the next frame's path tracer starts on its own stream as soon as the previous frame's temporal pass is
    done. A svgf_refit_bvh queued between two frames writes the triangle records and the tree that path tracer reads: it must be
    ordered before it (csrc/lbvh.cu invalidates the hand-over event). Frames queued back to back, meshes moved every frame,
    nothing synchronised in between: the bits of the serial schedule."""
    outs = []
    for on in ("1", "0"):
        monkeypatch.setenv("SVGF_FRAME_OVERLAP", on)
        m = svgf()
        W, H = 320, 200
        blob, R = m.open_scene("bunny", W, H)
        n = blob.counts["tris"]
        base = np.ascontiguousarray(blob.triangles).view(np.uint8).reshape(n, 136).copy()
        drv = blob.camera_driver(W, H)
        P = m.default_params(atrous_nlevel=3)
        for f in range(8):
            moved = base.view(np.float32).reshape(n, 34).copy()
            moved[:, [1, 9, 17]] += 0.05 * f
            R.refit_bvh(moved.view(np.uint8))
            R.pathtrace(drv.step(), P, f)
        outs.append([R.fetch("denoised"), R.fetch("image"), R.fetch("gbuffer"), R.fetch("history_length")])
        R.close()
    for a, b, what in zip(outs[0], outs[1], ("denoised", "image", "gbuffer", "history_length")):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), "%s differs between the overlapped and the serial schedule" % what
