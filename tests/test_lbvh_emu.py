"""CPU: the linear BVH of svgf_rebuild_bvh. The build steps (csrc/lbvh_core.h -- the functions the CUDA
kernels call) are run index by index on the host (tests/emu/lbvh_emu.cpp) over the reference scenes' triangles; checked are
the tree itself (every triangle in exactly one leaf, pre-order layout with left child = index + 1, boxes that contain their
subtree) and, through the ORACLE's traversal (the restatement of IntersectBVH, intersections.h:265-329), that frames rendered
with it equal frames rendered with the reference's own SAH tree: closest-hit results do not depend on the tree."""
import ctypes
import os
import struct
import subprocess

import numpy as np
import pytest

from util import ROOT, svgf
import orc

SRC = os.path.join(ROOT, "tests", "emu", "lbvh_emu.cpp")
LIB = os.path.join(ROOT, "tests", "emu", "liblbvh_emu.so")
HDR = os.path.join(ROOT, "cuda-path-tracer-denoising_b200", "csrc", "lbvh_core.h")


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", SRC, "-o", LIB], check=True)
    L = ctypes.CDLL(LIB)
    L.lbvh_emu_build.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return L


def hot_records(blob):
    """{v0, id} {e1, 0} {e2, 0} per triangle, as svgf_create packs them (csrc/api.cu:upload_scene)."""
    t = blob.triangles.reshape(-1, 136)
    n = t.shape[0]
    ids = t[:, 0:4].copy().view(np.int32)[:, 0]
    pos = t[:, 4:100].copy().view(np.float32).reshape(n, 3, 8)[:, :, 0:3]
    hot = np.zeros((n, 3, 4), np.float32)
    hot[:, 0, 0:3] = pos[:, 0]; hot[:, 0, 3] = ids.view(np.float32)
    hot[:, 1, 0:3] = pos[:, 1] - pos[:, 0]; hot[:, 2, 0:3] = pos[:, 2] - pos[:, 0]
    # the triangle as the intersector sees it: v0, v0 + e1, v0 + e2 with the ROUNDED edges (exact sums, float64)
    v0 = hot[:, 0, 0:3].astype(np.float64)
    eff = np.stack([v0, v0 + hot[:, 1, 0:3], v0 + hot[:, 2, 0:3]], axis=1)
    return hot, eff


def emu_build(L, hot):
    n = hot.shape[0]
    nodes = np.zeros((2 * n - 1, 2, 4), np.float32); order = np.zeros(n, np.int32)
    nn = L.lbvh_emu_build(np.ascontiguousarray(hot).ctypes.data, n, nodes.ctypes.data, order.ctypes.data)
    assert nn == 2 * n - 1
    return nodes, order


@pytest.mark.parametrize("name", ["cornell", "room", "bunny", "diamond"])
def test_tree_invariants(emu, name):
    m = svgf()
    blob = m.SceneBlob(m.scene_path(name))
    hot, pos = hot_records(blob)
    n = hot.shape[0]
    nodes, order = emu_build(emu, hot)
    assert sorted(order.tolist()) == list(range(n)), "every triangle exactly once"
    meta = nodes[:, 0, 3].copy().view(np.int32); off = nodes[:, 1, 3].copy().view(np.int32)
    lo, hi = nodes[:, 0, 0:3], nodes[:, 1, 0:3]
    seen = np.zeros(n, np.int32)

    def walk(i, depth):     # returns (subtree size, box of the triangles below)
        assert depth < 64, "the traversal stack is 64 deep (intersections.h:265)"
        if meta[i] & 0xFFFF:
            assert meta[i] == 1 and 0 <= off[i] < n
            seen[off[i]] += 1
            p = pos[order[off[i]]]
            assert (p >= lo[i]).all() and (p <= hi[i]).all(), "leaf box must contain its triangle"
            return 1, p.min(0), p.max(0)
        assert 0 <= (meta[i] >> 16) <= 2
        sl, l0, l1 = walk(i + 1, depth + 1)
        assert off[i] == i + 1 + sl, "right child follows the left subtree"
        sr, r0, r1 = walk(off[i], depth + 1)
        b0, b1 = np.minimum(l0, r0), np.maximum(l1, r1)
        assert (b0 >= lo[i]).all() and (b1 <= hi[i]).all(), "interior box must contain its subtree"
        return 1 + sl + sr, b0, b1

    import sys
    sys.setrecursionlimit(10000)
    size, _, _ = walk(0, 0)
    assert size == 2 * n - 1 and (seen == 1).all()


def test_single_triangle_and_tiny_inputs(emu):
    for n in (1, 2, 3):
        rng = np.random.default_rng(n)
        hot = np.zeros((n, 3, 4), np.float32)
        hot[:, 0, 0:3] = rng.uniform(-1, 1, (n, 3)); hot[:, 1, 0:3] = rng.uniform(0, 1, (n, 3)); hot[:, 2, 0:3] = rng.uniform(0, 1, (n, 3))
        nodes, order = emu_build(emu, hot)
        leaves = (nodes[:, 0, 3].copy().view(np.int32) & 0xFFFF) > 0
        assert leaves.sum() == n and sorted(order.tolist()) == list(range(n))


@pytest.mark.parametrize("name", ["room", "bunny"])
def test_oracle_renders_the_same_through_the_linear_bvh(emu, name, tmp_path):
    m = svgf()
    blob = m.SceneBlob(m.scene_path(name))
    hot, _ = hot_records(blob)
    n = hot.shape[0]
    nodes, order = emu_build(emu, hot)
    # the blob again, with the linear BVH (as BVH_ArrNode records) and the triangles in its order
    meta = nodes[:, 0, 3].copy().view(np.int32); off = nodes[:, 1, 3].copy().view(np.int32)
    rec = np.zeros((2 * n - 1, 10), np.float32)
    rec[:, 0:3] = nodes[:, 0, 0:3]; rec[:, 3:6] = nodes[:, 1, 0:3]
    ri = rec.view(np.int32)
    leaf = (meta & 0xFFFF) > 0
    ri[:, 6] = np.where(leaf, meta & 0xFFFF, 0); ri[:, 7] = np.where(leaf, 0, meta >> 16)
    ri[:, 8] = np.where(leaf, off, 0); ri[:, 9] = np.where(leaf, 0, off)
    raw = np.fromfile(m.scene_path(name), np.uint8)
    ng, nm, nt, nb, nx, ntex = (int(v) for v in raw[8:32].view(np.int32))
    o_tri = 40 + 84 + ng * 248 + nm * 56
    o_bvh = o_tri + nt * 136
    tris = raw[o_tri:o_bvh].reshape(nt, 136)[order]
    hdr = raw[:40].copy(); hdr[8:32].view(np.int32)[3] = 2 * n - 1
    out = np.concatenate([hdr, raw[40:o_tri], tris.reshape(-1), rec.view(np.uint8).reshape(-1), raw[o_bvh + nb * 40:]])
    path = str(tmp_path / (name + "_lbvh.scene"))
    out.tofile(path)
    W, H = 96, 64
    res = []
    for p in (m.scene_path(name), path):
        sc = orc.Scene(p); O = orc.Oracle(sc, W, H); P = orc.default_params(atrous_nlevel=2)
        drv = orc.CameraDriver(sc, W, H)
        for f in range(2):
            O.frame(drv.step(), P, f, orc.VAR_JACOBI, 0)
        res.append((O.fetch("gbuffer"), O.fetch("image")))
    g0, g1 = res[0][0], res[1][0]
    same = g0[..., 12].view(np.int32) == g1[..., 12].view(np.int32)
    assert same.mean() > 0.999, "geomId agreement %.5f" % same.mean()
    # where both trees found the same geom, position and radiance agree (different triangle of a tie: same point)
    assert np.allclose(g0[..., 3:6][same], g1[..., 3:6][same], atol=1e-4)
    d = np.abs(res[0][1] - res[1][1]).max(axis=2)
    assert (d > 1e-4).mean() < 5e-3, "%.4f of the 1-spp pixels differ" % (d > 1e-4).mean()
