"""ctypes driver for oracle/liboracle.so (the CPU restatement) -- TEST INFRASTRUCTURE.

Only tests/, bench.py's cpu_baseline/reference legs and __graft_entry__.smoke() may import this module.
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")
SCENE_DIR = os.path.join(os.path.dirname(_HERE), "tests", "golden", "scenes")

VAR_JACOBI, VAR_INPLACE_SEQ = 0, 1


class Camera(ctypes.Structure):     # svgf_camera, include/svgf_b200.h (== reference Camera, 84 B)
    _fields_ = [("resolution", ctypes.c_int32 * 2), ("position", ctypes.c_float * 3), ("lookAt", ctypes.c_float * 3),
                ("view", ctypes.c_float * 3), ("up", ctypes.c_float * 3), ("right", ctypes.c_float * 3),
                ("fov", ctypes.c_float * 2), ("pixelLength", ctypes.c_float * 2)]

    def as_array(self):
        return np.frombuffer(bytes(self), np.float32).copy()

    @classmethod
    def from_array(cls, a):
        return cls.from_buffer_copy(np.ascontiguousarray(a, np.float32).tobytes())


class CameraRig(ctypes.Structure):  # svgf_camera_rig
    _fields_ = [(n, ctypes.c_float) for n in ("zoom", "theta", "phi", "tx", "ty", "tz", "ttheta", "tphi", "fovy")]


class Params(ctypes.Structure):     # svgf_params
    _fields_ = [("tracedepth", ctypes.c_int32), ("shadowray", ctypes.c_int32), ("reducevar", ctypes.c_int32),
                ("sintensity", ctypes.c_float), ("lightradius", ctypes.c_float),
                ("denoise_enable", ctypes.c_int32), ("sepcolor", ctypes.c_int32), ("temporal_enable", ctypes.c_int32),
                ("color_alpha", ctypes.c_float), ("moment_alpha", ctypes.c_float), ("right_view_option", ctypes.c_int32),
                ("atrous_nlevel", ctypes.c_int32), ("spatial_enable", ctypes.c_int32), ("history_level", ctypes.c_int32),
                ("sigmal", ctypes.c_float), ("sigman", ctypes.c_float), ("sigmax", ctypes.c_float),
                ("blurvariance", ctypes.c_int32), ("addcolor", ctypes.c_int32), ("reserved_variance_mode", ctypes.c_int32)]


def default_params(**over):
    """src/main.cpp:39-62 defaults with the GUI 'All' button (preview.cpp:294-299)."""
    p = Params(tracedepth=4, shadowray=1, reducevar=1, sintensity=2.7, lightradius=1.4, denoise_enable=1, sepcolor=1,
               temporal_enable=1, color_alpha=0.2, moment_alpha=0.2, right_view_option=0, atrous_nlevel=5,
               spatial_enable=1, history_level=1, sigmal=0.45, sigman=0.2, sigmax=0.35, blurvariance=1, addcolor=1,
               reserved_variance_mode=0)
    for k, v in over.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


class TextureDesc(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32), ("components", ctypes.c_int32),
                ("pixels", ctypes.c_void_p)]


class SceneDesc(ctypes.Structure):
    _fields_ = [("geoms", ctypes.c_void_p), ("n_geoms", ctypes.c_int32),
                ("materials", ctypes.c_void_p), ("n_materials", ctypes.c_int32),
                ("triangles", ctypes.c_void_p), ("n_triangles", ctypes.c_int32),
                ("bvh_nodes", ctypes.c_void_p), ("n_bvh_nodes", ctypes.c_int32),
                ("textures", ctypes.c_void_p), ("n_textures", ctypes.c_int32),
                ("width", ctypes.c_int32), ("height", ctypes.c_int32)]


BUFFERS = {
    "image": (np.float32, (3,)), "denoised": (np.float32, (3,)), "gbuffer": (np.float32, (13,)),
    "intersections": (np.float32, (9,)), "variance": (np.float32, ()), "color_acc": (np.float32, (3,)),
    "color_history": (np.float32, (3,)), "moment_acc": (np.float32, (2,)), "moment_history": (np.float32, (2,)),
    "history_length": (np.int32, ()), "history_length_update": (np.int32, ()), "gbuffer_prev": (np.float32, (13,)),
    "temp0": (np.float32, (3,)), "temp1": (np.float32, (3,)), "host_image": (np.float32, (3,)),
}

_lib = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE, "oracle"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = ctypes.CDLL(LIB, mode=ctypes.RTLD_LOCAL)
        L.orc_scene_load.restype = ctypes.c_void_p
        L.orc_scene_load.argtypes = [ctypes.c_char_p]
        L.orc_scene_free.argtypes = [ctypes.c_void_p]
        L.orc_scene_counts.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_scene_camera.argtypes = [ctypes.c_void_p, ctypes.POINTER(Camera), ctypes.POINTER(ctypes.c_float)]
        L.orc_scene_desc.argtypes = [ctypes.c_void_p, ctypes.POINTER(SceneDesc), ctypes.c_void_p, ctypes.c_int]
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        L.orc_reset.argtypes = [ctypes.c_void_p]
        L.orc_frame.argtypes = [ctypes.c_void_p, ctypes.POINTER(Camera), ctypes.POINTER(Params), ctypes.c_int,
                                ctypes.c_int, ctypes.c_int]
        L.orc_fetch.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t]
        L.orc_denoise.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(Camera),
                                  ctypes.POINTER(Params), ctypes.c_int, ctypes.c_int]
        L.orc_host_intersect.argtypes = [ctypes.c_void_p] * 8
        L.orc_atrous_level.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 4 + [ctypes.c_float] * 3 + [ctypes.c_int] * 4
        L.orc_camera_init.argtypes = [ctypes.POINTER(Camera), ctypes.POINTER(CameraRig), ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_int]
        L.orc_camera_step.argtypes = [ctypes.POINTER(Camera), ctypes.POINTER(CameraRig), ctypes.c_int, ctypes.c_void_p]
        L.orc_view_matrix.argtypes = [ctypes.POINTER(Camera), ctypes.c_void_p]
        _lib = L
    return _lib


def scene_blob(name):
    return name if os.path.exists(name) else os.path.join(SCENE_DIR, name + ".scene")


class Scene:
    def __init__(self, name):
        self.h = lib().orc_scene_load(scene_blob(name).encode())
        if not self.h:
            raise RuntimeError("cannot load scene blob %s" % name)
        self.name = name

    def counts(self):
        a = np.zeros(6, np.int32)
        lib().orc_scene_counts(self.h, a.ctypes.data)
        return dict(zip(["geoms", "materials", "tris", "bvh", "boxes", "textures"], a.tolist()))

    def loader_camera(self):
        cam = Camera(); fovy = ctypes.c_float()
        lib().orc_scene_camera(self.h, ctypes.byref(cam), ctypes.byref(fovy))
        return cam, fovy.value

    def desc(self, W, H):
        """svgf_scene_desc over the oracle-held arrays (kept alive by this Scene)."""
        d = SceneDesc(); tex = (TextureDesc * 16)()
        if lib().orc_scene_desc(self.h, ctypes.byref(d), tex, 16):
            raise RuntimeError("too many textures")
        d.width, d.height = W, H
        self._tex = tex
        return d

    def intersect(self, origin, direction):
        o = np.asarray(origin, np.float32); d = np.asarray(direction, np.float32)
        t = np.zeros(1, np.float32); n = np.zeros(3, np.float32); uv = np.zeros(2, np.float32)
        g = np.zeros(1, np.int32); m = np.zeros(1, np.int32)
        hit = lib().orc_host_intersect(self.h, o.ctypes.data, d.ctypes.data, t.ctypes.data, n.ctypes.data,
                                       uv.ctypes.data, g.ctypes.data, m.ctypes.data)
        return hit, float(t[0]), n, uv, int(g[0]), int(m[0])


class CameraDriver:
    """Host camera logic either side of the path (resetCamera + runCuda's camera block), oracle restatement."""
    SPEEDS_C5 = (0.05, 0.02, 0.02, 0.02, 0.05)

    def __init__(self, scene, W, H, automate=False, speeds=SPEEDS_C5):
        lc, fovy = scene.loader_camera()
        self.cam, self.rig = Camera(), CameraRig()
        eye = np.array(lc.position[:], np.float32); la = np.array(lc.lookAt[:], np.float32); up = np.array(lc.up[:], np.float32)
        lib().orc_camera_init(ctypes.byref(self.cam), ctypes.byref(self.rig), eye.ctypes.data, la.ctypes.data,
                              up.ctypes.data, fovy, W, H)
        self.automate = automate
        self.speeds = np.array(speeds, np.float32)
        self.first = True

    def step(self):
        """Camera for the next frame (runCuda order: automation, then the camchanged block)."""
        if self.automate or self.first:
            lib().orc_camera_step(ctypes.byref(self.cam), ctypes.byref(self.rig), int(self.automate), self.speeds.ctypes.data)
            self.first = False
        return self.cam


class Oracle:
    def __init__(self, scene, W, H):
        self.scene, self.W, self.H = scene, W, H
        self.h = lib().orc_create(scene.h, W, H)

    def reset(self):
        lib().orc_reset(self.h)

    def frame(self, cam, params, frame, variance_mode=VAR_JACOBI, threads=0):
        rc = lib().orc_frame(self.h, ctypes.byref(cam), ctypes.byref(params), frame, variance_mode, threads)
        if rc:
            raise RuntimeError("orc_frame -> %d" % rc)

    def denoise(self, color_in, gbuffer, cam, params, variance_mode=VAR_JACOBI, threads=0):
        """denoise(output, input, gbuffer) (src/denoise.h:8) on numpy arrays in the reference's AoS layouts."""
        ci = np.ascontiguousarray(color_in, np.float32); g = np.ascontiguousarray(gbuffer, np.float32)
        out = np.empty_like(ci)
        rc = lib().orc_denoise(self.h, out.ctypes.data, ci.ctypes.data, g.ctypes.data, ctypes.byref(cam), ctypes.byref(params),
                               variance_mode, threads)
        if rc:
            raise RuntimeError("orc_denoise -> %d" % rc)
        return out

    def fetch(self, name):
        if name == "pbo":
            a = np.empty((self.H, 2 * self.W, 4), np.uint8)
        elif name == "view_matrix_prev":
            a = np.empty(16, np.float32)
        else:
            dt, tail = BUFFERS[name]
            a = np.empty((self.H, self.W) + tail, dt)
        rc = lib().orc_fetch(self.h, name.encode(), a.ctypes.data, a.nbytes)
        if rc:
            raise RuntimeError("orc_fetch(%s) -> %d" % (name, rc))
        return a

    def __del__(self):
        try:
            lib().orc_destroy(self.h)
        except Exception:
            pass


def atrous_level(color_in, variance_in, gbuffer, level, is_last, params, variance_mode=VAR_JACOBI, threads=0):
    """ATrousFilter (denoise.cu:77-170) on numpy planes: color (H,W,3) f32, variance (H,W) f32, gbuffer (H,W,13) f32."""
    H, W = variance_in.shape
    ci = np.ascontiguousarray(color_in, np.float32); vi = np.ascontiguousarray(variance_in, np.float32)
    g = np.ascontiguousarray(gbuffer, np.float32)
    co = np.empty_like(ci); vo = np.empty_like(vi)
    lib().orc_atrous_level(co.ctypes.data, vo.ctypes.data, ci.ctypes.data, vi.ctypes.data, g.ctypes.data, W, H, level,
                           int(is_last), params.sigmal, params.sigman, params.sigmax, params.blurvariance,
                           int(params.sepcolor and params.addcolor), variance_mode, threads)
    return co, vo
