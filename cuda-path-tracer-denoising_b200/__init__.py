"""cuda-path-tracer-denoising_b200 -- host-side binding of the B200-native SVGF + 1-spp path-trace hot path.

The product is the C-ABI library `libsvgf_b200.so` (csrc/, built in-tree for sm_100a; header include/svgf_b200.h).
This module is the thin Python mirror of that ABI used by tests/ and bench.py. It mirrors the reference's
operator interface for the path (src/pathtrace.h:6-8, src/denoise.h:6-8):

    reference                               here
    ---------                               ----
    pathtraceInit(scene); denoiseInit(scene)    Renderer(scene_desc, W, H)        -> svgf_create
    pathtraceFree(); denoiseFree()              Renderer.close()                  -> svgf_destroy
    "Clear" / frame 0 (main.cpp:192-201)        Renderer.reset()                  -> svgf_reset
    pathtrace(pbo, frame)                       Renderer.pathtrace(cam, params, frame, host_image=...) -> svgf_render
    denoise(out, in, gbuffer)                   Renderer.denoise(in, gbuffer, cam, params)             -> svgf_denoise_host
    ui_* globals (main.h:39-69)                 Params (svgf_params)

There is NO CPU fallback: importing works without a GPU (so the ABI can be inspected), but creating a Renderer
without a CUDA device, or without the built library, raises.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SVGF_LIB_PATH") or os.path.join(_HERE, "libsvgf_b200.so")      # SVGF_LIB_PATH: A/B builds (tools/)
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "svgf_b200.h")

SVGF_OK = 0


class SvgfError(RuntimeError):
    pass


class Camera(ctypes.Structure):     # svgf_camera == reference Camera (sceneStructs.h:74-83), 84 bytes
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


class Params(ctypes.Structure):     # svgf_params: the 19 ui_* globals the path reads
    _fields_ = [("tracedepth", ctypes.c_int32), ("shadowray", ctypes.c_int32), ("reducevar", ctypes.c_int32),
                ("sintensity", ctypes.c_float), ("lightradius", ctypes.c_float),
                ("denoise_enable", ctypes.c_int32), ("sepcolor", ctypes.c_int32), ("temporal_enable", ctypes.c_int32),
                ("color_alpha", ctypes.c_float), ("moment_alpha", ctypes.c_float), ("right_view_option", ctypes.c_int32),
                ("atrous_nlevel", ctypes.c_int32), ("spatial_enable", ctypes.c_int32), ("history_level", ctypes.c_int32),
                ("sigmal", ctypes.c_float), ("sigman", ctypes.c_float), ("sigmax", ctypes.c_float),
                ("blurvariance", ctypes.c_int32), ("addcolor", ctypes.c_int32), ("reserved_variance_mode", ctypes.c_int32)]


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


class Shard(ctypes.Structure):
    _fields_ = [("rank", ctypes.c_int32), ("world", ctypes.c_int32), ("row_begin", ctypes.c_int32), ("row_end", ctypes.c_int32)]


# every symbol include/svgf_b200.h declares (tests/test_abi.py checks the header against this list and the .so)
EXPORTS = ["svgf_params_default", "svgf_create", "svgf_destroy", "svgf_reset", "svgf_render", "svgf_denoise", "svgf_sync",
           "svgf_denoise_host", "svgf_atrous_host", "svgf_fetch", "svgf_last_error", "svgf_abi_version", "svgf_stage_times",
           "svgf_set_profiling", "svgf_stream", "svgf_set_shard", "svgf_ipc_handles_size", "svgf_ipc_export",
           "svgf_ipc_connect", "svgf_peer_connect_local", "svgf_peer_error", "svgf_camera_init", "svgf_camera_step",
           "svgf_render_async", "svgf_wait_image", "svgf_register_host", "svgf_unregister_host", "svgf_scene_load", "svgf_scene_free", "svgf_scene_error", "svgf_scene_describe",
           "svgf_set_option", "svgf_rebuild_bvh", "svgf_refit_bvh", "svgf_scene_camera", "svgf_scene_num_textures", "svgf_scene_texture_file", "svgf_scene_set_texture", "svgf_scene_mesh_boxes", "svgf_scene_load_textures", "svgf_jpeg_decode_memory"]

_lib = None


def build():
    """Compile csrc/ into libsvgf_b200.so (nvcc, sm_100a). Works without a GPU."""
    subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "csrc")], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SvgfError("libsvgf_b200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_LOCAL)
        vp, cp, ci, cf = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_float
        L.svgf_params_default.argtypes = [ctypes.POINTER(Params)]
        L.svgf_params_default.restype = None
        L.svgf_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(SceneDesc), ci]
        L.svgf_destroy.argtypes = [vp]
        L.svgf_reset.argtypes = [vp]
        L.svgf_render.argtypes = [vp, ctypes.POINTER(Camera), ctypes.POINTER(Params), ci, vp, vp]
        L.svgf_set_option.argtypes = [vp, cp, ci]
        L.svgf_rebuild_bvh.argtypes = [vp]
        L.svgf_refit_bvh.argtypes = [vp, vp, ci]
        L.svgf_scene_load.argtypes = [ctypes.POINTER(vp), cp, cp]
        L.svgf_scene_free.argtypes = [vp]; L.svgf_scene_free.restype = None
        L.svgf_scene_error.argtypes = [vp]; L.svgf_scene_error.restype = ctypes.c_char_p
        L.svgf_scene_describe.argtypes = [vp, ci, ci, ctypes.POINTER(SceneDesc)]
        L.svgf_scene_camera.argtypes = [vp, vp, vp, vp, vp, vp]
        L.svgf_scene_num_textures.argtypes = [vp]
        L.svgf_scene_texture_file.argtypes = [vp, ci]; L.svgf_scene_texture_file.restype = ctypes.c_char_p
        L.svgf_scene_set_texture.argtypes = [vp, ci, ci, ci, ci, vp]
        L.svgf_scene_mesh_boxes.argtypes = [vp, vp, ci]
        L.svgf_render_async.argtypes = [vp, ctypes.POINTER(Camera), ctypes.POINTER(Params), ci, vp, vp]
        L.svgf_wait_image.argtypes = [vp, vp]
        L.svgf_register_host.argtypes = [vp, vp, ctypes.c_size_t]
        L.svgf_unregister_host.argtypes = [vp, vp]
        L.svgf_denoise.argtypes = [vp, vp, vp, vp, ctypes.POINTER(Camera), ctypes.POINTER(Params)]
        L.svgf_denoise_host.argtypes = [vp, vp, vp, vp, ctypes.POINTER(Camera), ctypes.POINTER(Params)]
        L.svgf_atrous_host.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci, ctypes.POINTER(Params)]
        L.svgf_sync.argtypes = [vp]
        L.svgf_fetch.argtypes = [vp, cp, vp, ctypes.c_size_t]
        L.svgf_last_error.argtypes = [vp]
        L.svgf_last_error.restype = cp
        L.svgf_stage_times.argtypes = [vp, vp]
        L.svgf_set_profiling.argtypes = [vp, ci]
        L.svgf_stream.argtypes = [vp]
        L.svgf_stream.restype = vp
        L.svgf_set_shard.argtypes = [vp, ctypes.POINTER(Shard)]
        L.svgf_ipc_export.argtypes = [vp, vp]
        L.svgf_ipc_connect.argtypes = [vp, ci, ci, vp, vp]
        L.svgf_peer_connect_local.argtypes = [vp, ci, vp]
        L.svgf_peer_error.argtypes = [vp]
        L.svgf_camera_init.argtypes = [ctypes.POINTER(Camera), ctypes.POINTER(CameraRig), vp, vp, vp, cf, ci, ci]
        L.svgf_camera_init.restype = None
        L.svgf_camera_step.argtypes = [ctypes.POINTER(Camera), ctypes.POINTER(CameraRig), ci, vp]
        L.svgf_camera_step.restype = None
        _lib = L
    return _lib


def default_params(**over):
    """main.cpp:39-62 defaults with the GUI "All" button (preview.cpp:294-299)."""
    p = Params()
    lib().svgf_params_default(ctypes.byref(p))
    for k, v in over.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


BUFFERS = {
    "image": (np.float32, (3,)), "denoised": (np.float32, (3,)), "host_image": (np.float32, (3,)),
    "gbuffer": (np.float32, (13,)), "variance": (np.float32, ()), "color_history": (np.float32, (3,)),
    "variance_history": (np.float32, ()), "moment_acc": (np.float32, (2,)), "moment_history": (np.float32, (2,)),
    "history_length": (np.int32, ()), "history_length_update": (np.int32, ()),
}


class CameraDriver:
    """resetCamera + the camera block of runCuda (main.cpp:77-101, 154-190) through the C ABI."""
    SPEEDS_C5 = (0.05, 0.02, 0.02, 0.02, 0.05)

    def __init__(self, eye, lookat, up, fovy, W, H, automate=False, speeds=SPEEDS_C5):
        self.cam, self.rig = Camera(), CameraRig()
        e, l, u = (np.asarray(v, np.float32) for v in (eye, lookat, up))
        lib().svgf_camera_init(ctypes.byref(self.cam), ctypes.byref(self.rig), e.ctypes.data, l.ctypes.data, u.ctypes.data,
                               fovy, W, H)
        self.automate, self.speeds, self.first = automate, np.asarray(speeds, np.float32), True

    def step(self):
        if self.automate or self.first:
            lib().svgf_camera_step(ctypes.byref(self.cam), ctypes.byref(self.rig), int(self.automate), self.speeds.ctypes.data)
            self.first = False
        return self.cam


class Renderer:
    """One svgf_ctx: pathtraceInit+denoiseInit ... pathtrace(pbo, frame) ... pathtraceFree+denoiseFree."""

    def __init__(self, scene_desc, W, H, device=0):
        self.W, self.H = W, H
        self._desc = scene_desc     # keeps host arrays alive during create
        scene_desc.width, scene_desc.height = W, H
        h = ctypes.c_void_p()
        rc = lib().svgf_create(ctypes.byref(h), ctypes.byref(scene_desc), device)
        if rc != SVGF_OK:
            raise SvgfError("svgf_create -> %d: %s" % (rc, lib().svgf_last_error(None).decode()))
        self.h = h

    def _ck(self, rc, what):
        if rc != SVGF_OK:
            raise SvgfError("%s -> %d: %s" % (what, rc, lib().svgf_last_error(self.h).decode()))

    def close(self):
        if getattr(self, "h", None):
            lib().svgf_destroy(self.h)
            self.h = None

    __del__ = close

    def reset(self):
        self._ck(lib().svgf_reset(self.h), "svgf_reset")

    def set_shard(self, rank, world, row_begin, row_end):
        s = Shard(rank, world, row_begin, row_end)
        self._ck(lib().svgf_set_shard(self.h, ctypes.byref(s)), "svgf_set_shard")

    def ipc_export(self):
        """This rank's IPC handles (bytes) for svgf_ipc_connect on the other ranks."""
        buf = np.zeros(lib().svgf_ipc_handles_size(), np.uint8)
        self._ck(lib().svgf_ipc_export(self.h, buf.ctypes.data), "svgf_ipc_export")
        return buf

    def ipc_connect(self, rank, world, all_handles, row_starts):
        a = np.ascontiguousarray(all_handles, np.uint8); rs = np.ascontiguousarray(row_starts, np.int32)
        assert a.size == world * lib().svgf_ipc_handles_size() and rs.size == world + 1
        self._ck(lib().svgf_ipc_connect(self.h, rank, world, a.ctypes.data, rs.ctypes.data), "svgf_ipc_connect")

    def peer_error(self):
        return lib().svgf_peer_error(self.h)

    def set_profiling(self, on):
        self._ck(lib().svgf_set_profiling(self.h, int(on)), "svgf_set_profiling")

    def stage_times(self):
        a = np.zeros(11, np.float32)
        self._ck(lib().svgf_stage_times(self.h, a.ctypes.data), "svgf_stage_times")
        return a

    def stream(self):
        """cudaStream_t (int) the context issues its work on."""
        return lib().svgf_stream(self.h)

    def pathtrace(self, cam, params, frame, pbo_dev=None, host_image=None):
        """== pathtrace(pbo, frame). host_image: (H, W, 3) float32 numpy array to fill (scene->state.image), or None."""
        hp = host_image.ctypes.data if host_image is not None else None
        self._ck(lib().svgf_render(self.h, ctypes.byref(cam), ctypes.byref(params), frame, pbo_dev, hp), "svgf_render")

    def pathtrace_async(self, cam, params, frame, host_image, pbo_dev=None):
        """Pipelined pathtrace(): queues the frame and the copy of its image into `host_image` and returns; the image is
        complete after wait_image(host_image). Alternate between two host arrays to keep one frame in flight."""
        self._ck(lib().svgf_render_async(self.h, ctypes.byref(cam), ctypes.byref(params), frame, pbo_dev, host_image.ctypes.data),
                 "svgf_render_async")

    def register_host(self, array):
        """Page-lock a numpy array for direct DMA (svgf_register_host); keep it alive until unregister_host/close."""
        self._ck(lib().svgf_register_host(self.h, array.ctypes.data, array.nbytes), "svgf_register_host")

    def unregister_host(self, array):
        self._ck(lib().svgf_unregister_host(self.h, array.ctypes.data), "svgf_unregister_host")

    def wait_image(self, host_image=None):
        self._ck(lib().svgf_wait_image(self.h, host_image.ctypes.data if host_image is not None else None), "svgf_wait_image")

    def rebuild_bvh(self):
        """Linear BVH built on the device over the context's triangles, swapped in for the uploaded tree (svgf_rebuild_bvh)."""
        self._ck(lib().svgf_rebuild_bvh(self.h), "svgf_rebuild_bvh")

    def refit_bvh(self, triangles):
        """New vertex data (n x 136-byte svgf_triangle records, create order) into the tree in place (svgf_refit_bvh)."""
        t = np.ascontiguousarray(triangles).view(np.uint8).reshape(-1)
        self._ck(lib().svgf_refit_bvh(self.h, t.ctypes.data, t.size // 136), "svgf_refit_bvh")

    def fetch_raw(self, name, count, dtype):
        a = np.empty(count, dtype)
        self._ck(lib().svgf_fetch(self.h, name.encode(), a.ctypes.data, a.nbytes), "svgf_fetch(%s)" % name)
        return a

    def set_option(self, name, value):
        """Quality switches beyond the reference (all off by default): 'reprojection_fov_aspect', 'history_cap', 'spatial_variance_estimate', 'light_sampling_all'."""
        self._ck(lib().svgf_set_option(self.h, name.encode(), int(value)), "svgf_set_option(%s)" % name)

    def sync(self):
        self._ck(lib().svgf_sync(self.h), "svgf_sync")

    def denoise(self, color_in, gbuffer, cam, params):
        """== denoise(output, input, gbuffer) on host arrays in the reference's AoS layouts."""
        ci = np.ascontiguousarray(color_in, np.float32); g = np.ascontiguousarray(gbuffer, np.float32)
        out = np.empty_like(ci)
        self._ck(lib().svgf_denoise_host(self.h, out.ctypes.data, ci.ctypes.data, g.ctypes.data, ctypes.byref(cam),
                                         ctypes.byref(params)), "svgf_denoise_host")
        return out

    def atrous_level(self, color_in, variance_in, gbuffer, level, is_last, params):
        ci = np.ascontiguousarray(color_in, np.float32); vi = np.ascontiguousarray(variance_in, np.float32)
        g = np.ascontiguousarray(gbuffer, np.float32)
        co = np.empty_like(ci); vo = np.empty_like(vi)
        self._ck(lib().svgf_atrous_host(self.h, co.ctypes.data, vo.ctypes.data, ci.ctypes.data, vi.ctypes.data, g.ctypes.data,
                                        level, int(is_last), ctypes.byref(params)), "svgf_atrous_host")
        return co, vo

    def fetch(self, name):
        if name == "pbo":
            a = np.empty((self.H, 2 * self.W, 4), np.uint8)
        elif name == "view_matrix_prev":
            a = np.empty(16, np.float32)
        else:
            dt, tail = BUFFERS[name]
            a = np.empty((self.H, self.W) + tail, dt)
        self._ck(lib().svgf_fetch(self.h, name.encode(), a.ctypes.data, a.nbytes), "svgf_fetch(%s)" % name)
        return a


# ---- scene blobs (tests/golden/scenes/*.scene) -------------------------------------------------------
# The arrays the reference's Scene loader produces (src/scene.cpp), in the reference's own struct layouts,
# exported once by oracle/ref/harness.cpp:refh_export_scene. Layout: 40-byte header {"SVGFSCN1", n_geoms,
# n_materials, n_tris, n_bvh, n_boxes, n_textures, fovy, 0}, Camera (84 B), Geom[248 B], Material[56 B],
# Triangle[136 B], BVH_ArrNode[40 B], BoundingBox[24 B], then per texture {w, h, comp} + RGB8 bytes.
class SceneBlob:
    def __init__(self, path):
        raw = np.fromfile(path, np.uint8)
        if raw[:8].tobytes() != b"SVGFSCN1":
            raise SvgfError("%s is not a scene blob" % path)
        hdr = raw[8:32].view(np.int32)
        ng, nm, nt, nb, nx, ntex = (int(v) for v in hdr)
        self.fovy = float(raw[32:36].view(np.float32)[0])
        off = 40
        self.loader_camera = Camera.from_buffer_copy(raw[off:off + 84].tobytes()); off += 84
        def take(n, sz):
            nonlocal off
            a = np.ascontiguousarray(raw[off:off + n * sz]).copy(); off += n * sz
            return a
        self.geoms, self.materials = take(ng, 248), take(nm, 56)
        self.triangles, self.bvh = take(nt, 136), take(nb, 40)
        take(nx, 24)
        self.textures = []
        for _ in range(ntex):
            w, h, c = (int(v) for v in raw[off:off + 12].view(np.int32)); off += 12
            self.textures.append((w, h, c, take(w * h * c, 1)))
        self.counts = dict(geoms=ng, materials=nm, tris=nt, bvh=nb, boxes=nx, textures=ntex)

    def desc(self, W, H):
        d = SceneDesc()
        d.geoms, d.n_geoms = self.geoms.ctypes.data, self.counts["geoms"]
        d.materials, d.n_materials = self.materials.ctypes.data, self.counts["materials"]
        d.triangles, d.n_triangles = self.triangles.ctypes.data, self.counts["tris"]
        d.bvh_nodes, d.n_bvh_nodes = self.bvh.ctypes.data, self.counts["bvh"]
        self._tex = (TextureDesc * max(1, len(self.textures)))()
        for i, (w, h, c, px) in enumerate(self.textures):
            self._tex[i] = TextureDesc(w, h, c, px.ctypes.data)
        d.textures, d.n_textures = ctypes.cast(self._tex, ctypes.c_void_p), len(self.textures)
        d.width, d.height = W, H
        d._keepalive = self
        return d

    def camera_driver(self, W, H, automate=False, speeds=CameraDriver.SPEEDS_C5):
        lc = self.loader_camera
        return CameraDriver(lc.position[:], lc.lookAt[:], lc.up[:], self.fovy, W, H, automate, speeds)


def row_partition(H, world):
    """Row strips [start[r], start[r+1]) of a frame of H rows over `world` ranks."""
    return [r * H // world for r in range(world)] + [H]


def connect_local(renderers, row_starts=None):
    """Wire several Renderers living in this process as the ranks of one sharded frame (svgf_peer_connect_local)."""
    world = len(renderers)
    rs = np.ascontiguousarray(row_starts if row_starts is not None else row_partition(renderers[0].H, world), np.int32)
    arr = (ctypes.c_void_p * world)(*[r.h for r in renderers])
    rc = lib().svgf_peer_connect_local(arr, world, rs.ctypes.data)
    if rc != SVGF_OK:
        raise SvgfError("svgf_peer_connect_local -> %d" % rc)


class SceneFile:
    """svgf_scene_*: the reference's text scene + OBJ meshes parsed, transformed and BVH-built natively (csrc/scene_ingest.cpp).
    Same duck type as SceneBlob for Renderer/camera_driver: .desc(W, H), .camera_driver(W, H, automate)."""

    def __init__(self, scene_file, models_dir=None):
        h = ctypes.c_void_p()
        rc = lib().svgf_scene_load(ctypes.byref(h), scene_file.encode(), models_dir.encode() if models_dir else None)
        self.h = h
        if rc != 0:
            msg = lib().svgf_scene_error(h).decode() if h else "svgf_scene_load failed"
            self.close()
            raise SvgfError(msg)
        eye = np.zeros(3, np.float32); look = np.zeros(3, np.float32); up = np.zeros(3, np.float32)
        fovy = ctypes.c_float(); res = np.zeros(2, np.int32)
        lib().svgf_scene_camera(h, eye.ctypes.data, look.ctypes.data, up.ctypes.data, ctypes.addressof(fovy), res.ctypes.data)
        self.eye, self.lookat, self.up, self.fovy, self.res = eye, look, up, float(fovy.value), (int(res[0]), int(res[1]))
        self.texture_files = [lib().svgf_scene_texture_file(h, i).decode() for i in range(lib().svgf_scene_num_textures(h))]

    def load_textures(self, textures_dir):
        """Decodes every texture still without pixels from <textures_dir>/<file> (svgf_scene_load_textures: the library's own JPEG
        decoder, byte-identical to the reference's stb_image). Returns how many textures have pixels now."""
        lib().svgf_scene_load_textures.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        n = lib().svgf_scene_load_textures(self.h, textures_dir.encode())
        if n < 0:
            raise SvgfError("svgf_scene_load_textures failed")
        return n

    def set_texture(self, index, pixels):
        """pixels: (H, W, C) uint8, row-major (what stb_image hands the reference, src/sceneStructs.h:193-199)."""
        a = np.ascontiguousarray(pixels, np.uint8)
        rc = lib().svgf_scene_set_texture(self.h, index, a.shape[1], a.shape[0], a.shape[2], a.ctypes.data)
        if rc != 0:
            raise SvgfError("svgf_scene_set_texture failed")

    def desc(self, W, H):
        d = SceneDesc()
        if lib().svgf_scene_describe(self.h, W, H, ctypes.byref(d)) != 0:
            raise SvgfError(lib().svgf_scene_error(self.h).decode())
        d._keepalive = self     # the arrays live inside the svgf_scene: keep it alive as long as the description
        return d

    def arrays(self, W=1, H=1):
        """Copies of the ingest result as raw byte arrays in the reference's struct layouts (tests)."""
        d = self.desc(W, H)
        def grab(ptr, n, sz):
            return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)), (max(n * sz, 0),)).copy() if n else np.zeros(0, np.uint8)
        boxes = np.zeros((max(lib().svgf_scene_mesh_boxes(self.h, None, 0), 1), 6), np.float32)
        nb = lib().svgf_scene_mesh_boxes(self.h, boxes.ctypes.data, boxes.shape[0])
        return dict(geoms=grab(d.geoms, d.n_geoms, 248), materials=grab(d.materials, d.n_materials, 56),
                    triangles=grab(d.triangles, d.n_triangles, 136), bvh=grab(d.bvh_nodes, d.n_bvh_nodes, 40), boxes=boxes[:nb])

    def camera_driver(self, W, H, automate=False, speeds=CameraDriver.SPEEDS_C5):
        return CameraDriver(self.eye, self.lookat, self.up, self.fovy, W, H, automate, speeds)

    def close(self):
        if getattr(self, "h", None):
            lib().svgf_scene_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def scene_path(name):
    if os.path.exists(name):
        return name
    return os.path.join(os.path.dirname(_HERE), "tests", "golden", "scenes", name + ".scene")


def jpeg_decode(data):
    """svgf_jpeg_decode_memory: bytes of a JPEG file -> (H, W, C) uint8, the pixels stb_image's stbi_load(..., 0) returns."""
    L = lib()
    L.svgf_jpeg_decode_memory.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                          ctypes.POINTER(ctypes.c_int), ctypes.c_void_p, ctypes.c_size_t]
    b = np.frombuffer(bytes(data), np.uint8).copy()
    w, h, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    if L.svgf_jpeg_decode_memory(b.ctypes.data, b.size, ctypes.byref(w), ctypes.byref(h), ctypes.byref(c), None, 0) != 0:
        raise SvgfError("svgf_jpeg_decode_memory: not a JPEG this decoder reads")
    out = np.zeros((h.value, w.value, c.value), np.uint8)
    if L.svgf_jpeg_decode_memory(b.ctypes.data, b.size, ctypes.byref(w), ctypes.byref(h), ctypes.byref(c), out.ctypes.data, out.size) != 0:
        raise SvgfError("svgf_jpeg_decode_memory failed")
    return out


def open_scene(name, W, H, device=0):
    """Convenience: blob -> (SceneBlob, Renderer)."""
    blob = SceneBlob(scene_path(name))
    return blob, Renderer(blob.desc(W, H), W, H, device)


def gather_handles(local_handles, dist, world):
    """All-gather every rank's IPC handle bytes in rank order over torch.distributed (any backend): the plumbing step
    between svgf_ipc_export and svgf_ipc_connect. Returns a (world * n,) uint8 numpy array."""
    import torch
    mine = torch.from_numpy(np.ascontiguousarray(local_handles, np.uint8).copy())
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = mine.to(dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return torch.cat(out).cpu().numpy()


def connect_ranks(renderer, dist, rank, world, row_starts=None):
    """Shard `renderer`'s frame over the ranks of an initialised torch.distributed group (one process per GPU)."""
    rs = list(row_starts) if row_starts is not None else row_partition(renderer.H, world)
    allh = gather_handles(renderer.ipc_export(), dist, world)
    renderer.ipc_connect(rank, world, allh, rs)
    dist.barrier()
    return rs


def balanced_partition(row_starts, cost_ms, min_rows=8):
    """New strip boundaries that equalise the measured per-rank cost, assuming each strip's cost is spread evenly over
    its rows (piecewise-constant cost density). Path-tracing cost per row depends on what the rows see, so equal-height
    strips are not equal-time strips; a static, measured partition fixes that without touching the data path."""
    world = len(cost_ms)
    H = row_starts[-1]
    dens = []
    for r in range(world):
        rows = max(1, row_starts[r + 1] - row_starts[r])
        dens += [max(float(cost_ms[r]), 1e-6) / rows] * (row_starts[r + 1] - row_starts[r])
    cum = np.concatenate([[0.0], np.cumsum(dens)])
    target = cum[-1] / world
    out = [0]
    for r in range(1, world):
        y = int(np.argmin(np.abs(cum - r * target)))
        y = min(max(y, out[-1] + min_rows), H - (world - r) * min_rows)
        out.append(y)
    out.append(H)
    return out
