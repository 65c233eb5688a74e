"""ctypes driver for the reference harness libraries (oracle/_ref/libref_*.so) -- TEST INFRASTRUCTURE.

The libraries contain the reference's own src/pathtrace.cu + src/denoise.cu (compiled from /root/reference by
oracle/Makefile) behind the headless harness oracle/ref/harness.cpp. Variants:
  cpu / cpu_jacobi : kernels executed on the host through oracle/ref/cuda_emu (runs without a GPU)
  gpu / gpu_jacobi : nvcc sm_100 build (runs on the B200 box)
  shim             : the *product's* drop-in shim behind the same harness (our kernels, reference entry points)
`*_jacobi` = ATrousFilter writes variance to a second buffer (race-free; see oracle/ref/stage.sh).
Only tests/, bench.py's reference/cpu_baseline legs and __graft_entry__.smoke() may import this module.
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
SCENE_DIR = os.path.join(os.path.dirname(_HERE), "tests", "golden", "scenes")

# name -> (numpy dtype, trailing shape) for refh_fetch
BUFFERS = {
    "image": (np.float32, (3,)), "denoised": (np.float32, (3,)), "gbuffer": (np.float32, (13,)),
    "intersections": (np.float32, (9,)), "paths": (np.float32, (12,)),
    "variance": (np.float32, ()), "color_acc": (np.float32, (3,)), "color_history": (np.float32, (3,)),
    "moment_acc": (np.float32, (2,)), "moment_history": (np.float32, (2,)),
    "history_length": (np.int32, ()), "history_length_update": (np.int32, ()),
    "gbuffer_prev": (np.float32, (13,)), "temp0": (np.float32, (3,)), "temp1": (np.float32, (3,)),
    "host_image": (np.float32, (3,)),
}

ALL_ON = dict(denoise_enable=1, temporal_enable=1, spatial_enable=1, sepcolor=1, addcolor=1)


def lib_path(variant):
    return os.path.join(REF_DIR, "libref_%s.so" % variant)


def available(variant):
    return os.path.exists(lib_path(variant))


def scene_blob(name):
    return os.path.join(SCENE_DIR, name + ".scene")


class RefHarness:
    """One reference instance per process per variant (the reference keeps its state in file statics)."""

    def __init__(self, variant, path=None):
        self.variant = variant
        self.lib = ctypes.CDLL(path or lib_path(variant), mode=ctypes.RTLD_LOCAL)
        L = self.lib
        L.refh_load_scene_blob.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
        L.refh_set_param.argtypes = [ctypes.c_char_p, ctypes.c_double]
        L.refh_fetch.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t]
        L.refh_time_frames.argtypes = [ctypes.c_int]
        L.refh_time_frames.restype = ctypes.c_double
        L.refh_scene_counts.argtypes = [ctypes.c_void_p]
        if hasattr(L, "refh_load_scene_txt"):
            L.refh_load_scene_txt.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
            L.refh_export_scene.argtypes = [ctypes.c_char_p]
        if hasattr(L, "refh_denoise_host"):
            L.refh_denoise_host.argtypes = [ctypes.c_void_p] * 3
            L.refh_set_camera.argtypes = [ctypes.c_void_p]
        if hasattr(L, "refh_host_intersect"):
            L.refh_host_intersect.argtypes = [ctypes.c_void_p] * 7
        self.W = self.H = 0

    def load_blob(self, name_or_path, W, H):
        p = name_or_path if os.path.exists(name_or_path) else scene_blob(name_or_path)
        rc = self.lib.refh_load_scene_blob(p.encode(), W, H)
        if rc:
            raise RuntimeError("refh_load_scene_blob(%s) -> %d" % (p, rc))
        self.W, self.H = W, H

    def load_txt(self, scene_txt, W, H, reference_root="/root/reference"):
        rc = self.lib.refh_load_scene_txt(reference_root.encode(), scene_txt.encode(), W, H)
        if rc:
            raise RuntimeError("refh_load_scene_txt(%s) -> %d" % (scene_txt, rc))
        self.W, self.H = W, H

    def export_scene(self, out_path):
        rc = self.lib.refh_export_scene(out_path.encode())
        if rc:
            raise RuntimeError("refh_export_scene -> %d" % rc)

    def set_params(self, **kw):
        for k, v in kw.items():
            if self.lib.refh_set_param(k.encode(), float(v)):
                raise KeyError(k)

    def frame(self):
        r = self.lib.refh_frame()
        if r < 0:
            raise RuntimeError("refh_frame -> %d" % r)
        return r

    def init(self):
        """pathtraceFree/Init + denoiseFree/Init without rendering (the reference's reset order, main.cpp:192-201)."""
        if self.lib.refh_init():
            raise RuntimeError("refh_init failed")

    def set_camera(self, cam21):
        a = np.ascontiguousarray(cam21, np.float32)
        assert a.nbytes == 84
        if self.lib.refh_set_camera(a.ctypes.data):
            raise RuntimeError("refh_set_camera failed")

    def denoise(self, color_in, gbuffer):
        """denoise(output, input, gbuffer) (src/denoise.h:8) of whatever this library links: the reference's or the shim's."""
        ci = np.ascontiguousarray(color_in, np.float32); g = np.ascontiguousarray(gbuffer, np.float32)
        out = np.empty_like(ci)
        if self.lib.refh_denoise_host(out.ctypes.data, ci.ctypes.data, g.ctypes.data):
            raise RuntimeError("refh_denoise_host failed")
        return out

    def time_frames(self, n):
        return self.lib.refh_time_frames(n)

    def fetch(self, name):
        if name == "pbo":
            a = np.empty((self.H, 2 * self.W, 4), np.uint8)
        elif name == "camera":
            a = np.empty(21, np.float32)
        else:
            dt, tail = BUFFERS[name]
            a = np.empty((self.H, self.W) + tail, dt)
        rc = self.lib.refh_fetch(name.encode(), a.ctypes.data, a.nbytes)
        if rc:
            raise RuntimeError("refh_fetch(%s) -> %d" % (name, rc))
        return a

    def scene_counts(self):
        a = np.zeros(6, np.int32)
        self.lib.refh_scene_counts(a.ctypes.data)
        return dict(zip(["geoms", "materials", "tris", "bvh", "boxes", "textures"], a.tolist()))

    def host_intersect(self, origin, direction):
        o = np.asarray(origin, np.float32); d = np.asarray(direction, np.float32)
        t = np.zeros(1, np.float32); n = np.zeros(3, np.float32); uv = np.zeros(2, np.float32)
        g = np.zeros(1, np.int32); m = np.zeros(1, np.int32)
        hit = self.lib.refh_host_intersect(o.ctypes.data, d.ctypes.data, t.ctypes.data, n.ctypes.data,
                                           uv.ctypes.data, g.ctypes.data, m.ctypes.data)
        return hit, float(t[0]), n, uv, int(g[0]), int(m[0])


def gbuffer_fields(g):
    """Split a (H, W, 13) float32 view of 52-byte GBufferTexel (sceneStructs.h:113-119)."""
    return dict(normal=g[..., 0:3], position=g[..., 3:6], albedo=g[..., 6:9], ialbedo=g[..., 9:12],
                geomId=g[..., 12].view(np.int32))
