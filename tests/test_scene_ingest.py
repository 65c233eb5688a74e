"""Scene ingest (CPU, no GPU needed): svgf_scene_load -- the reference's text scene format, OBJ meshes,
world-space transform and SAH BVH build restated in csrc/scene_ingest.cpp -- against the arrays the REFERENCE'S OWN loader
produced (tests/golden/scenes/*.scene, exported by oracle/ref/harness.cpp:refh_export_scene from src/scene.cpp +
src/bvhtree.cpp + tinyobjloader, with the fields the reference leaves uninitialised zeroed). Bar: byte for byte -- geoms
(matrices included), materials, BVH-ordered triangles, BVH nodes, mesh boxes, camera block.

The scene text files and OBJ models are the reference's data (scenes/*.txt, scenes/Models/*.obj); they are read where they
lie, and these tests skip when /root/reference is absent (the GPU box). A scene written for this repo (tests/golden/
scenes_txt) covers the parser's branches everywhere."""
import os

import numpy as np
import pytest

from util import ROOT, svgf

REF_SCENES = "/root/reference/scenes"
OWN_SCENES = os.path.join(ROOT, "tests", "golden", "scenes_txt")
needs_reference = pytest.mark.skipif(not os.path.isdir(REF_SCENES), reason="reference scene files not present")


@needs_reference
@pytest.mark.parametrize("name", ["cornell", "room", "bunny", "diamond"])
def test_ingest_equals_reference_loader_bytes(name):
    m = svgf()
    blob = m.SceneBlob(m.scene_path(name))
    sc = m.SceneFile(os.path.join(REF_SCENES, name + ".txt"))
    for i, (w, h, c, px) in enumerate(blob.textures):      # decoded pixels come from outside (stb_image in the reference)
        sc.set_texture(i, px.reshape(h, w, c))
    a = sc.arrays()
    assert len(sc.texture_files) == len(blob.textures)
    for key, ref, sz in (("geoms", blob.geoms, 248), ("materials", blob.materials, 56), ("triangles", blob.triangles, 136), ("bvh", blob.bvh, 40)):
        got = a[key]
        assert got.size == ref.size, "%s: %d vs %d records" % (key, got.size // sz, ref.size // sz)
        if not np.array_equal(got, ref):
            bad = np.nonzero(got.reshape(-1, sz) != ref.reshape(-1, sz))
            raise AssertionError("%s differs: first at record %d byte %d (%d bytes in all)" % (key, bad[0][0], bad[1][0], bad[0].size))
    lc = blob.loader_camera
    assert np.array_equal(sc.eye, np.array(lc.position[:], np.float32)) and np.array_equal(sc.lookat, np.array(lc.lookAt[:], np.float32))
    assert np.array_equal(sc.up, np.array(lc.up[:], np.float32)) and sc.fovy == blob.fovy
    # the camera the two sources drive through resetCamera / camchanged is the same, bit for bit
    c0, c1 = blob.camera_driver(96, 64).step(), sc.camera_driver(96, 64).step()
    assert np.array_equal(c0.as_array().view(np.uint32), c1.as_array().view(np.uint32))
    sc.close()


@needs_reference
def test_mesh_boxes_follow_the_reference_quirk():
    """Scene::BoudningBoxs starts its maxima at FLT_MIN (a tiny POSITIVE number, scene.cpp:255), kept for fidelity."""
    m = svgf()
    raw = np.fromfile(m.scene_path("bunny"), np.uint8)
    hdr = raw[8:32].view(np.int32)
    ng, nm, nt, nb, nx, ntex = (int(v) for v in hdr)
    off = 40 + 84 + ng * 248 + nm * 56 + nt * 136 + nb * 40
    ref_boxes = raw[off:off + nx * 24].view(np.float32).reshape(nx, 6)
    sc = m.SceneFile(os.path.join(REF_SCENES, "bunny.txt"))
    assert np.array_equal(sc.arrays()["boxes"].view(np.uint32), ref_boxes.view(np.uint32))
    sc.close()


def test_own_scene_parses_and_is_consistent():
    """A scene written for this repo: CRLF line ends, comments, a polygon face with negative indices, a v/vt/vn mesh."""
    m = svgf()
    sc = m.SceneFile(os.path.join(OWN_SCENES, "two_meshes.txt"))
    assert sc.res == (320, 200) and sc.fovy == 40.0 and sc.texture_files == ["checker.jpg"]
    with pytest.raises(m.SvgfError, match="no pixels"):
        sc.desc(8, 8)                       # a TEXTURE line was seen but nobody attached decoded pixels yet
    sc.set_texture(0, np.zeros((2, 2, 3), np.uint8))
    assert sc.desc(8, 8).n_textures == 1
    a = sc.arrays()
    geoms = a["geoms"].reshape(-1, 248); tris = a["triangles"].reshape(-1, 136); bvh = a["bvh"].reshape(-1, 40)
    assert geoms.shape[0] == 4 and a["materials"].size == 3 * 56
    types = geoms[:, 0:4].copy().view(np.int32)[:, 0]
    assert list(types) == [1, 0, 2, 2]
    # quad (fan of 2) + pyramid (4 triangles + a quad base = 6)
    assert tris.shape[0] == 8
    ids = sorted(tris[:, 0:4].copy().view(np.int32)[:, 0].tolist())
    assert ids == list(range(8)), "triangle ids are the load order, each exactly once"
    rng = geoms[:, 236:248].copy().view(np.int32)
    assert rng[2, 0] == 0 and rng[2, 1] == 2 and rng[3, 0] == 2 and rng[3, 1] == 8
    # every triangle is referenced by exactly one leaf, children follow their parent, bounds contain the triangles
    n = bvh.shape[0]
    f = bvh.copy().view(np.float32).reshape(n, 10); q = bvh.copy().view(np.int32).reshape(n, 10)
    seen = np.zeros(tris.shape[0], np.int32)
    pos = tris[:, 4:100].copy().view(np.float32).reshape(-1, 3, 8)[:, :, 0:3]
    for i in range(n):
        if q[i, 6] > 0:
            sl = slice(q[i, 8], q[i, 8] + q[i, 6]); seen[sl] += 1
            assert (pos[sl] >= f[i, 0:3] - 1e-6).all() and (pos[sl] <= f[i, 3:6] + 1e-6).all()
        else:
            assert i + 1 < n and i < q[i, 9] < n and 0 <= q[i, 7] <= 2
    assert (seen == 1).all()
    # the identity-transformed quad keeps its coordinates (whichever slots the BVH order gave its two triangles)
    quad = pos[np.isin(tris[:, 0:4].copy().view(np.int32)[:, 0], [0, 1])]
    assert set(np.unique(quad).tolist()) == {-4.0, 0.0, 4.0}
    sc.close()


def test_errors_are_reported():
    m = svgf()
    with pytest.raises(m.SvgfError, match="cannot open scene file"):
        m.SceneFile("/nonexistent/scene.txt")
    with pytest.raises(m.SvgfError, match="cannot open OBJ"):
        m.SceneFile(os.path.join(OWN_SCENES, "missing_mesh.txt"))


@needs_reference
@pytest.mark.parametrize("name", ["cornell", "room"])
def test_textures_decoded_natively_equal_the_reference_loaders_bytes(name):
    """The reference decodes its JPEG textures with stb_image (src/sceneStructs.h:198-199); svgf_scene_load_textures uses the
    library's own decoder (csrc/jpeg_decode.cpp: baseline and progressive Huffman, stb_image's fixed-point IDCT, chroma
    up-sampling and YCbCr conversion). The texels enter the image through Texture::getColor, so the bar is byte for byte:
    wallpaper.jpg is progressive 4:4:4 with an Adobe marker, chair.jpg baseline 4:2:0. With it the whole ingest -- text scene,
    OBJ meshes, BVH, textures -- is native: the description below is built without any pixels from outside."""
    m = svgf()
    blob = m.SceneBlob(m.scene_path(name))
    sc = m.SceneFile(os.path.join(REF_SCENES, name + ".txt"))
    assert sc.load_textures(os.path.join(REF_SCENES, "Textures")) == len(blob.textures)
    d = sc.desc(64, 48)
    assert d.n_textures == len(blob.textures)
    for i, (w, h, c, px) in enumerate(blob.textures):
        got = m.jpeg_decode(open(os.path.join(REF_SCENES, "Textures", sc.texture_files[i]), "rb").read())
        assert got.shape == (h, w, c)
        assert np.array_equal(got.reshape(-1), px), "%s: %d texels differ from stb_image's" % (sc.texture_files[i], int((got.reshape(-1) != px).sum()))


def test_jpeg_decoder_rejects_what_it_does_not_read():
    m = svgf()
    for junk in (b"", b"not a jpeg", b"\xff\xd8\xff\xd9", b"\xff\xd8" + b"\xff\xc0\x00\x0b\x10\x00\x01\x00\x01\x01\x01\x11\x00"):      # empty, text, no frame, 16-bit samples
        with pytest.raises(m.SvgfError):
            m.jpeg_decode(junk)
    sc = m.SceneFile(os.path.join(OWN_SCENES, "textured.txt")) if os.path.exists(os.path.join(OWN_SCENES, "textured.txt")) else None
    if sc is not None and sc.texture_files:
        assert sc.load_textures("/nonexistent") == 0        # a missing file is reported, not fatal


JPEG_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpeg")


def _jpeg_cases():
    import json
    return sorted(json.load(open(os.path.join(JPEG_GOLDEN, "index.json"))).items())


@pytest.mark.parametrize("name,whc", _jpeg_cases(), ids=lambda v: v if isinstance(v, str) else "")
def test_jpeg_decoder_equals_stb_image_on_every_coding_mode(name, whc):
    """tests/golden/jpeg: small files in every mode the decoder reads -- baseline and progressive, 4:4:4 / 4:2:2 / 4:2:0, greyscale,
    sizes that are not a multiple of the MCU, optimised Huffman tables, restart intervals, a single pixel -- with the bytes the
    reference's stb_image returned for them (make_fixtures.py, run where /root/reference exists). Byte for byte."""
    m = svgf()
    w, h, c = whc
    want = np.fromfile(os.path.join(JPEG_GOLDEN, name + ".rgb"), dtype=np.uint8)
    got = m.jpeg_decode(open(os.path.join(JPEG_GOLDEN, name + ".jpg"), "rb").read())
    assert got.shape == (h, w, c)
    assert np.array_equal(got.reshape(-1), want), "%s: %d bytes differ from stb_image's" % (name, int((got.reshape(-1) != want).sum()))


def test_jpeg_decoder_survives_damaged_files():
    """A deterministic slice of tools/fuzz/fuzz_jpeg.cpp through the C ABI: truncated, bit-flipped and spliced copies of the
    fixtures either decode to an image of the announced size or are rejected with an error -- never anything else."""
    m = svgf()
    rng = np.random.default_rng(7)
    decoded = rejected = 0
    for name, _ in _jpeg_cases():
        base = np.frombuffer(open(os.path.join(JPEG_GOLDEN, name + ".jpg"), "rb").read(), dtype=np.uint8)
        for it in range(60):
            v = base.copy()
            mode = it % 4
            if mode == 0:
                v = v[:int(rng.integers(0, len(v)))]
            elif mode == 1:
                for _ in range(int(rng.integers(1, 6))):
                    v[int(rng.integers(0, len(v)))] ^= np.uint8(1 << int(rng.integers(0, 8)))
            elif mode == 2:
                for _ in range(int(rng.integers(1, 6))):
                    v[int(rng.integers(0, len(v)))] = np.uint8(rng.integers(0, 256))
            else:
                a, b = sorted(int(x) for x in rng.integers(0, len(v), 2))
                v = np.concatenate([v[:a], v[b:]])
            try:
                img = m.jpeg_decode(v.tobytes())
                assert img.ndim == 3 and img.shape[2] in (1, 3) and img.size > 0
                decoded += 1
            except m.SvgfError:
                rejected += 1
    assert decoded > 0 and rejected > 0


def test_scene_reader_survives_damaged_files(tmp_path):
    """A deterministic slice of tools/fuzz/fuzz_scene.cpp and fuzz_obj.cpp through the C ABI: damaged scene texts and OBJ files are
    either loaded (and described) or rejected with a message -- no crash, no endless BVH recursion on non-finite vertices."""
    import shutil
    m = svgf()
    rng = np.random.default_rng(11)
    words = [b"MATERIAL", b"OBJECT", b"CAMERA", b"mesh", b"cube", b"sphere", b"TRANS", b"SCALE", b"nan", b"inf", b"1e38", b"-1", b"\n", b" "]
    models = tmp_path / "Models"
    shutil.copytree(os.path.join(OWN_SCENES, "Models"), models)
    loaded = rejected = 0
    for name in ("two_meshes.txt", "two_lights.txt"):
        base = open(os.path.join(OWN_SCENES, name), "rb").read()
        for it in range(80):
            v = bytearray(base)
            for _ in range(int(rng.integers(1, 5))):
                pos = int(rng.integers(0, max(1, len(v))))
                mode = int(rng.integers(0, 4))
                if mode == 0: del v[pos:]
                elif mode == 1 and v: v[pos] = int(rng.integers(32, 127))
                elif mode == 2: v[pos:pos] = words[int(rng.integers(0, len(words)))]
                elif v: del v[pos:pos + int(rng.integers(1, 30))]
            f = tmp_path / "s.txt"
            f.write_bytes(bytes(v))
            try:
                sc = m.SceneFile(str(f), str(models))
                loaded += 1
            except m.SvgfError as e:
                assert str(e)
                rejected += 1
    # OBJ files with vertices that overflow or are not numbers at all
    for junk in (b"v nan 0 0\nv 0 1 0\nv 1 0 0\nf 1 2 3\n", b"v 1e39 0 0\nv 0 1 0\nv 1 0 0\nf 1 2 3\n" * 1,
                 b"v 0 0 0\nv 0 1 0\nv 1 0 0\n" + b"f 1 2 3\n" * 40, b"f 1 2 3\n", b"v 0 0 0\nf 1 1 1 1 1 1 1\n"):
        (models / "quad.obj").write_bytes(junk)
        try:
            m.SceneFile(os.path.join(OWN_SCENES, "two_meshes.txt"), str(models)); loaded += 1
        except m.SvgfError:
            rejected += 1
    assert loaded > 0 and rejected > 0
