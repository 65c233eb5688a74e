"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/svgf_b200.h
declares, keeps the reference's struct layouts, refuses to run without a GPU (no CPU fallback), and its host-side
camera logic reproduces the reference's main.cpp camera handling bit for bit (via the pinned oracle)."""
import ctypes
import os
import re

import numpy as np
import pytest

from util import ROOT, svgf, have_gpu
import orc


def test_library_exports_every_declared_symbol():
    m = svgf()
    hdr = open(m.HEADER_PATH).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(svgf_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(m.EXPORTS), "header and binding disagree: %r" % (declared ^ set(m.EXPORTS))
    L = ctypes.CDLL(m.LIB_PATH)
    for s in sorted(declared):
        assert hasattr(L, s), "libsvgf_b200.so does not export %s" % s
    assert L.svgf_abi_version() == 1


def test_struct_sizes_match_reference_abi():
    m = svgf()
    # Camera 84 B; the rest is static_assert'ed in csrc/api.cu against the same table
    assert ctypes.sizeof(m.Camera) == 84
    assert ctypes.sizeof(m.Params) == 80
    assert ctypes.sizeof(m.SceneDesc) == 88
    src = open(os.path.join(ROOT, m.PKG_DIR if hasattr(m, "PKG_DIR") else "cuda-path-tracer-denoising_b200", "csrc", "api.cu")).read()
    for t, n in (("svgf_geom", 248), ("svgf_material", 56), ("svgf_triangle", 136), ("svgf_bvh_node", 40),
                 ("svgf_gbuffer_texel", 52), ("svgf_path_segment", 48), ("svgf_intersection", 36)):
        assert "sizeof(%s) == %d" % (t, n) in src


def test_default_params_are_the_reference_defaults():
    m = svgf()
    p = m.default_params(); o = orc.default_params()
    for name, _ in m.Params._fields_:
        assert getattr(p, name) == pytest.approx(getattr(o, name)), name
    assert (p.tracedepth, p.atrous_nlevel, p.history_level) == (4, 5, 1)         # main.cpp:42,57,58
    assert p.sigmal == pytest.approx(0.45) and p.sigmax == pytest.approx(0.35) and p.sigman == pytest.approx(0.2)


@pytest.mark.skipif(have_gpu(), reason="checks the behaviour on a machine WITHOUT a GPU")
def test_no_cpu_fallback():
    m = svgf()
    blob = m.SceneBlob(m.scene_path("cornell"))
    with pytest.raises(m.SvgfError, match="no usable CUDA device"):
        m.Renderer(blob.desc(32, 32), 32, 32)


@pytest.mark.parametrize("scene", ["cornell", "room", "bunny", "diamond"])
@pytest.mark.parametrize("res", [(64, 64), (1920, 1080), (3840, 2160)])
@pytest.mark.parametrize("moving", [False, True])
def test_camera_logic_matches_oracle_bit_for_bit(scene, res, moving):
    """svgf_camera_init/step + view matrix vs the oracle's restatement, which tests/test_oracle_vs_reference.py pins
    against the reference's own resetCamera/runCuda/GetViewMatrix."""
    m = svgf()
    W, H = res
    blob = m.SceneBlob(m.scene_path(scene))
    drv = blob.camera_driver(W, H, automate=moving)
    odrv = orc.CameraDriver(orc.Scene(scene), W, H, automate=moving)
    for f in range(6):
        a = drv.step().as_array(); b = odrv.step().as_array()
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "camera differs at frame %d" % f


def test_scene_blob_reader_matches_oracle_reader():
    m = svgf()
    for scene in ("cornell", "room", "bunny", "diamond"):
        blob = m.SceneBlob(m.scene_path(scene))
        assert blob.counts == orc.Scene(scene).counts()


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/svgf_b200.h must compile as C99 (no C++/CUDA/torch types), and a C program must link
    against the library through it (host-only calls; nothing here touches a GPU)."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text('#include "svgf_b200.h"\n#include <stdio.h>\n'
                   "int main(void) {\n"
                   "  svgf_params p; svgf_camera cam; svgf_camera_rig rig; svgf_scene *sc = 0; svgf_scene_desc d;\n"
                   "  const float eye[3] = {0, 5, 10}, at[3] = {0, 5, 0}, up[3] = {0, 1, 0};\n"
                   "  svgf_params_default(&p);\n"
                   "  svgf_camera_init(&cam, &rig, eye, at, up, 45.0f, 64, 48);\n"
                   "  if (svgf_scene_load(&sc, \"/nonexistent.txt\", 0) == SVGF_OK) return 2;\n"
                   "  printf(\"%d %d %d %s\\n\", svgf_abi_version(), p.atrous_nlevel, cam.resolution[0], svgf_scene_error(sc));\n"
                   "  svgf_scene_free(sc); (void)d;\n"
                   "  return (sizeof(svgf_geom) == 248 && sizeof(svgf_gbuffer_texel) == 52) ? 0 : 1;\n}\n")
    libdir = os.path.join(ROOT, "cuda-path-tracer-denoising_b200")
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lsvgf_b200", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split(" ", 3)
    assert out[0] == "1" and out[1] == "5" and out[2] == "64" and "cannot open scene file" in out[3]
