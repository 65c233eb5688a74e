/* svgf_b200.h -- C ABI of the B200-native SVGF + 1-spp path-trace hot path.
 *
 * This is the drop-in boundary for the reference's per-frame path
 *     pathtrace(uchar4* pbo, int frame)            /root/reference/src/pathtrace.h:8  (pathtrace.cu:404-452)
 *       -> denoise(vec3* out, vec3* in, GBufferTexel*)   src/denoise.h:8            (denoise.cu:349-402)
 * and its lifecycle calls pathtraceInit/Free (pathtrace.h:6-7) and denoiseInit/Free (denoise.h:6-7).
 * Plain pointers and sizes only; no C++/torch/glm types. Every function returns 0 on success or a
 * negative svgf_status; svgf_last_error() gives the text. Nothing here ever falls back to a CPU path:
 * if the CUDA device or the kernels are unavailable the call fails.
 *
 * The reference's data contract is kept byte for byte: the structs below have the sizes and field offsets
 * of the reference's structs, so a Scene built by the reference's loader
 * can be handed over without conversion (see INTEGRATION.md for the C++ shim that does this).
 */
#ifndef SVGF_B200_H
#define SVGF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVGF_ABI_VERSION 1

typedef enum svgf_status {
    SVGF_OK = 0,
    SVGF_ERR_INVALID = -1,      /* bad argument / call order */
    SVGF_ERR_CUDA = -2,         /* CUDA runtime error (text in svgf_last_error) */
    SVGF_ERR_NO_DEVICE = -3,    /* no usable CUDA device: there is no CPU fallback */
    SVGF_ERR_COMM = -4,         /* multi-GPU exchange failed */
    SVGF_ERR_UNKNOWN_NAME = -5  /* svgf_fetch: no such buffer */
} svgf_status;

/* ---- reference-layout PODs (all alignof 4) ------------------------------------------------------------ */

/* Geom, sceneStructs.h:33-47 (248 B). Matrices are column-major (glm::mat4 memory order). */
typedef struct svgf_geom {
    int32_t type;               /* 0 SPHERE, 1 CUBE, 2 MESH (sceneStructs.h:18-22) */
    int32_t materialid;
    float translation[3], rotation[3], scale[3];
    float transform[16], inverseTransform[16], invTranspose[16];
    int32_t T_startidx, T_endidx, BoundIdx;
} svgf_geom;

/* Material, sceneStructs.h:49-72 (56 B) */
typedef struct svgf_material {
    float color[3];
    float specular_exponent;
    float specular_color[3];
    float hasReflective, hasRefractive, indexOfRefraction, emittance;
    int32_t matid, texid, norid;
} svgf_material;

/* Vertex / Triangle, sceneStructs.h:24-31, 121-181 (32 B / 136 B). World space, BVH order. */
typedef struct svgf_vertex { float pos[3], normal[3], uv[2]; } svgf_vertex;
typedef struct svgf_triangle {
    int32_t id;                 /* load-order id; geom ranges [T_startidx, T_endidx) refer to it */
    svgf_vertex verts[3];
    float normal[3];            /* unused by the hot path */
    float bb_max[3], bb_min[3]; /* unused by the hot path */
} svgf_triangle;

/* BVH_ArrNode, bvhtree.h:48-54 (40 B). Left child = index + 1. */
typedef struct svgf_bvh_node {
    float bounds_min[3], bounds_max[3];
    int32_t primitive_count;    /* > 0: leaf */
    int32_t axis;
    int32_t primitivesOffset;
    int32_t rightchildoffset;
} svgf_bvh_node;

/* Camera, sceneStructs.h:74-83 (84 B). `right` is unnormalised (main.cpp:179-184). */
typedef struct svgf_camera {
    int32_t resolution[2];
    float position[3], lookAt[3], view[3], up[3], right[3];
    float fov[2];
    float pixelLength[2];
} svgf_camera;

/* GBufferTexel, sceneStructs.h:113-119 (52 B) */
typedef struct svgf_gbuffer_texel {
    float normal[3], position[3], albedo[3], ialbedo[3];
    int32_t geomId;             /* -1 = miss */
} svgf_gbuffer_texel;

/* PathSegment, sceneStructs.h:92-99 (48 B) */
typedef struct svgf_path_segment {
    float origin[3], direction[3], color[3];
    int32_t pixelIndex, remainingBounces;
    uint8_t diffuse, specular, pad_[2];
} svgf_path_segment;

/* ShadeableIntersection, sceneStructs.h:104-111 (36 B); persistent per pixel across frames */
typedef struct svgf_intersection {
    float t, surfaceNormal[3];
    int32_t materialId, geomId;
    uint8_t outside, pad_[3];
    float uv[2];
} svgf_intersection;

/* Texture pixels, sceneStructs.h:184-222: RGB8, row-major, `components` must be 3 to be sampled. */
typedef struct svgf_texture_desc {
    int32_t width, height, components;
    const unsigned char *pixels;    /* host pointer, copied at svgf_create */
} svgf_texture_desc;

/* What pathtraceInit(Scene*) reads from the Scene (pathtrace.cu:103-158). Host pointers, copied. */
typedef struct svgf_scene_desc {
    const svgf_geom *geoms;             int32_t n_geoms;
    const svgf_material *materials;     int32_t n_materials;
    const svgf_triangle *triangles;     int32_t n_triangles;
    const svgf_bvh_node *bvh_nodes;     int32_t n_bvh_nodes;
    const svgf_texture_desc *textures;  int32_t n_textures;
    int32_t width, height;              /* cam.resolution at Init time */
} svgf_scene_desc;

/* The 19 `ui_*` globals the reference's hot path reads at link time (main.h:39-69), passed explicitly.
 * Field names follow the globals (ui_tracedepth -> tracedepth ...). Booleans are int32 0/1. */
typedef struct svgf_params {
    int32_t tracedepth;         /* pathtrace.cu:421,425 */
    int32_t shadowray, reducevar;
    float sintensity, lightradius;
    int32_t denoise_enable, sepcolor;   /* pathtrace.cu:429,436 */
    int32_t temporal_enable;    /* denoise.cu:360-362 */
    float color_alpha, moment_alpha;
    int32_t right_view_option;  /* denoise.cu:373-378: 0 filter, 1 history length /100, 2 variance /0.1 */
    int32_t atrous_nlevel, spatial_enable, history_level;   /* denoise.cu:380-391 */
    float sigmal, sigman, sigmax;
    int32_t blurvariance, addcolor;
    /* Not a reference global. The reference updates variance[] in place while neighbouring threads read
     * it (denoise.cu:111,153,161), so its a-trous output is not deterministic. 0 (default) = race-free
     * double-buffered variance: every read of a level sees the previous level's values. */
    int32_t reserved_variance_mode;
} svgf_params;

/* Defaults = src/main.cpp:39-62 with the GUI "All" button pressed (preview.cpp:294-299). */
void svgf_params_default(svgf_params *p);

/* Row strip owned by this process when a frame is sharded over several GPUs.
 * rows [row_begin, row_end) of the full W x H frame; world == 1 means the whole frame. */
typedef struct svgf_shard {
    int32_t rank, world;
    int32_t row_begin, row_end;
} svgf_shard;

typedef struct svgf_ctx svgf_ctx;

/* ---- lifecycle: pathtraceInit + denoiseInit / pathtraceFree + denoiseFree ------------------------------ */
int svgf_create(svgf_ctx **out, const svgf_scene_desc *scene, int device);
int svgf_destroy(svgf_ctx *ctx);                 /* NULL is a no-op, like Free before the first Init */
/* "Clear"/frame 0 (main.cpp:192-201): history_length = 0, moments = 0, variance = 0, image = 0,
 * persistent intersections = 0. view_matrix_prev survives (denoise.cu:15). */
int svgf_reset(svgf_ctx *ctx);

/* ---- the per-frame path -------------------------------------------------------------------------------- */
/* == pathtrace(pbo, frame), pathtrace.cu:404-452. `pbo_dev` (device, 2W x H uchar4) and `host_image`
 * (host, W*H*3 float, = scene->state.image) may be NULL. Synchronous like the reference when host_image is
 * given; otherwise work is left queued on the context's stream (svgf_sync to wait). */
int svgf_render(svgf_ctx *ctx, const svgf_camera *cam, const svgf_params *params, int frame,
                void *pbo_dev, float *host_image);
/* Pipelined form of svgf_render (the reference blocks on a pageable D2H every frame,
 * pathtrace.cu:450). Queues the frame AND the copy of its image into `host_image` (page-locked by the library on first
 * sight) and returns; the copy runs on a second stream while the next frame renders. `host_image` holds the frame once
 * svgf_wait_image(ctx, host_image) (or svgf_sync) has returned; alternate between two host buffers to keep one frame in
 * flight. Results are bit-identical to svgf_render. */
int svgf_render_async(svgf_ctx *ctx, const svgf_camera *cam, const svgf_params *params, int frame,
                      void *pbo_dev, float *host_image);
/* Blocks until the image most recently queued into `host_image` has arrived (NULL: every queued image). */
int svgf_wait_image(svgf_ctx *ctx, const float *host_image);
/* Page-locking of caller memory. svgf_render copies into `host_image` by direct DMA when the buffer is page-locked (by the
 * caller, or registered here) and through the context's own staging buffer otherwise -- it never registers memory behind
 * the caller's back. svgf_render_async needs page-locked memory and registers an unregistered buffer on first sight. A
 * registered buffer MUST stay allocated until svgf_unregister_host or svgf_destroy: freeing or resizing it earlier leaves
 * the registration on pages the process no longer owns. */
int svgf_register_host(svgf_ctx *ctx, void *host, size_t bytes);
int svgf_unregister_host(svgf_ctx *ctx, void *host);
/* == denoise(output, input, gbuffer), denoise.cu:349-402, on caller-owned DEVICE buffers in the reference's
 * AoS layouts (vec3 colour, 52-byte texels). Stream semantics are the reference's: the work is ordered after everything the
 * caller queued in the legacy default stream before the call (e.g. the cudaMemcpy that filled the inputs, which may return
 * before its DMA has landed), and has completed when the call returns (denoise.cu:401). */
int svgf_denoise(svgf_ctx *ctx, float *output_dev, const float *input_dev, const svgf_gbuffer_texel *gbuffer_dev,
                 const svgf_camera *cam, const svgf_params *params);
int svgf_sync(svgf_ctx *ctx);

/* Host-buffer convenience for callers without device memory of their own (tests, benches, FFI users):
 * copies `input`/`gbuffer` H2D, runs svgf_denoise, copies `output` back. */
int svgf_denoise_host(svgf_ctx *ctx, float *output, const float *input, const svgf_gbuffer_texel *gbuffer,
                      const svgf_camera *cam, const svgf_params *params);

/* One edge-avoiding a-trous level (ATrousFilter, denoise.cu:77-170) on HOST planes; `level` in [1,7],
 * step = 1 << level. colour: W*H*3, variance: W*H, gbuffer: 52-byte texels. Used by the parity tests and
 * the roofline bench. */
int svgf_atrous_host(svgf_ctx *ctx, float *color_out, float *variance_out, const float *color_in,
                     const float *variance_in, const svgf_gbuffer_texel *gbuffer, int level, int is_last,
                     const svgf_params *params);

/* ---- introspection (goldens, tests) -------------------------------------------------------------------- */
/* Copies a named intermediate to host. Names follow the reference's file-static buffers:
 * image, denoised, gbuffer (52 B AoS, repacked), intersections, variance, color_acc, color_history,
 * moment_acc, moment_history, history_length, pbo. `bytes` must match exactly. */
int svgf_fetch(svgf_ctx *ctx, const char *name, void *host, size_t bytes);
const char *svgf_last_error(const svgf_ctx *ctx);    /* ctx may be NULL: last create error */
int svgf_abi_version(void);

/* Per-stage event timing (off by default; it adds 12 event records per frame to the stream, no host syncs).
 * Enabling it (re)starts a measurement window of at most 512 frames. */
int svgf_set_profiling(svgf_ctx *ctx, int enabled);
/* Average device time (ms) per stage over the frames rendered since profiling was enabled, from CUDA events on the
 * context's stream: [0] path trace, [1] temporal, [2..8] a-trous level 1..7, [9] pbo pack, [10] whole frame.
 * Synchronises the stream. */
int svgf_stage_times(svgf_ctx *ctx, float *ms11);
/* The CUDA stream (cudaStream_t) all of the context's work is issued on, for callers that bracket it with their
 * own events. */
void *svgf_stream(svgf_ctx *ctx);

/* ---- quality switches beyond the reference; all 0 = the reference's behaviour, bit for bit ------- */
/*   "reprojection_fov_aspect" 0/1  the temporal back-projection divides by tan(fovy)*aspect and tan(fovy). The reference
 *                                  leaves that term out (denoise.cu:200-207), so its history only lines up for FOVY 45 on
 *                                  square frames: at 16:9 a static camera maps x to cx + 1.78 (x - cx) and most of the
 *                                  frame never accumulates.
 *   "history_cap"             n    history_length saturates at n > 0 (unbounded in the reference, denoise.cu:290-294).
 *   "spatial_variance_estimate" 0/1  pixels whose history is shorter than 4 frames estimate their variance from the luminance
 *                                  moments of their 7x7 neighbourhood on the same surface (SVGF paper, section 4.2) instead of
 *                                  the constants 100 / 10 the reference assigns (denoise.cu:315, 320-329: EstimateVariance is a
 *                                  stub). Single-GPU, whole-frame contexts.
 *   "cuda_graph"              0/1  (frame driver) the frame's kernels are launched as one CUDA graph that is
 *                                  re-captured and updated in place every frame; results are bit-identical.
 *   "frame_overlap"           0/1  (frame driver) single-GPU frames: the path tracer of the next frame runs on a second stream next
 *                                  to the a-trous stage of the current one (second set of the buffers both touch); bit-identical.
 *                                  Helps when many frames are queued without waiting (+2-6 %), hurts a caller that waits for
 *                                  every image (-4-19 %): off by default.
 *   "light_sampling_all"      0/1  every shadow ray samples one of the scene's emissive cubes/spheres (uniformly, contribution
 *                                  scaled by their number) instead of geoms[0] only (pathtrace.cu:359-361). */
int svgf_set_option(svgf_ctx *ctx, const char *name, int value);

/* ---- BVH build on the device -------------------------------------------------------------------- */
/* Replaces the tree uploaded by svgf_create (the reference's host-built SAH tree, src/bvhtree.cpp) with a linear BVH built on
 * the GPU over the context's triangles (Morton codes, radix sort, Karras' radix tree; csrc/lbvh_core.h), re-ordering the
 * triangle records to match. Same node semantics (pre-order, left child = index + 1, BVH_ArrNode, src/bvhtree.h:48-54), so the
 * traversal is unchanged and frames match the host-built tree up to ties between equal hit distances. For geometry that
 * changes on the device; synchronises the context's stream. svgf_fetch names: "bvh_packed", "triangle_ids". */
int svgf_rebuild_bvh(svgf_ctx *ctx);
/* Geometry that moves keeps its tree: new vertex data for ALL triangles, in the order they were given to svgf_create, is written
 * into the device records and the boxes of the tree in place (the uploaded one, or the one svgf_rebuild_bvh built) are recomputed
 * bottom-up; topology and leaf contents stay, so the traversal order does too. Stream-ordered with the frames, no
 * synchronisation. The reference has no counterpart (it rebuilds on the host and re-uploads through pathtraceInit). */
int svgf_refit_bvh(svgf_ctx *ctx, const svgf_triangle *triangles, int n_triangles);

/* ---- multi-GPU: a frame sharded by row strips, one process (context) per GPU -------------------------------- */
/* Every rank holds full-frame planes and renders rows [row_starts[rank], row_starts[rank+1]); rows owned by other
 * ranks are read in place from the owner's memory over NVLink (CUDA IPC), ordered by per-stage flags -- no halo copy,
 * no collective on the data path. All ranks must call svgf_reset/svgf_render in lock-step with identical arguments.
 *   1. every rank:  svgf_ipc_export(ctx, my_handles)          (svgf_ipc_handles_size() bytes)
 *   2. all-gather the handles in rank order (torch.distributed / MPI / any byte transport)
 *   3. every rank:  svgf_ipc_connect(ctx, rank, world, all_handles, row_starts)   (row_starts: world + 1 ints, 0 .. H)
 * The reference has no multi-GPU path; this is new. */
int svgf_ipc_handles_size(void);
int svgf_ipc_export(svgf_ctx *ctx, void *handles_out);
int svgf_ipc_connect(svgf_ctx *ctx, int rank, int world, const void *all_handles, const int *row_starts);
/* Ranks that live in one process (several contexts on one or more GPUs): wire them without IPC. */
int svgf_peer_connect_local(svgf_ctx **ctxs, int world, const int *row_starts);
/* 1 if a cross-rank wait gave up (a peer made no progress for ~2 s), else 0. The frame in flight continued on stale rows, so
 * from then on svgf_render / svgf_render_async / svgf_sync / svgf_wait_image return SVGF_ERR_COMM until svgf_reset. */
int svgf_peer_error(svgf_ctx *ctx);
/* Restrict a lone context to a row strip (no peers: taps outside the strip read this context's own planes). */
int svgf_set_shard(svgf_ctx *ctx, const svgf_shard *shard);

/* ---- camera control: the host logic either side of the path (main.cpp:77-101, 154-190) ------------------ */
typedef struct svgf_camera_rig {
    float zoom, theta, phi;
    float tx, ty, tz, ttheta, tphi;     /* automation phases (main.cpp:24-28) */
    float fovy;
} svgf_camera_rig;
/* Fill resolution-dependent fields (scene.cpp:159-166) and derive zoom/theta/phi (resetCamera, main.cpp:77-101). */
void svgf_camera_init(svgf_camera *cam, svgf_camera_rig *rig, const float eye[3], const float lookat[3],
                      const float up[3], float fovy, int width, int height);
/* Optional automation step (main.cpp:156-169) followed by the camchanged block (main.cpp:171-190). */
void svgf_camera_step(svgf_camera *cam, svgf_camera_rig *rig, int automate, const float speeds[5]);

/* ---- scene ingest: the reference's scene files -> the arrays svgf_create takes ---------------- */
/* Parses the reference's text scene format (MATERIAL / OBJECT / CAMERA blocks, src/scene.cpp:9-238) and its OBJ meshes,
 * transforms the triangles to world space (scene.cpp:240-311) and builds the SAH BVH (src/bvhtree.cpp), reproducing the
 * arrays of the reference's own loader bit for bit (csrc/scene_ingest.cpp). `models_dir` NULL = "<dir of scene_file>/Models"
 * (the reference hard-codes ../scenes/Models/, scene.cpp:236). A scene object is returned even on failure so that
 * svgf_scene_error() can say why; free it in both cases. Host only: no CUDA call is made. */
typedef struct svgf_scene svgf_scene;
int svgf_scene_load(svgf_scene **out, const char *scene_file, const char *models_dir);
void svgf_scene_free(svgf_scene *scene);
const char *svgf_scene_error(const svgf_scene *scene);
/* Fills `desc` with pointers into `scene` (valid until svgf_scene_free); every texture must have pixels by then. */
int svgf_scene_describe(svgf_scene *scene, int width, int height, svgf_scene_desc *desc);
/* The CAMERA block: EYE / LOOKAT / UP / FOVY / RES, the inputs of svgf_camera_init. Any pointer may be NULL. */
int svgf_scene_camera(const svgf_scene *scene, float eye[3], float lookat[3], float up[3], float *fovy, int res[2]);
/* Textures are referenced by file name (TEXTURE lines, scene.cpp:213-219; the reference decodes "../scenes/Textures/<name>" with
 * stb_image, sceneStructs.h:198-199). Either attach decoded pixels (RGB8, row-major) with svgf_scene_set_texture, or let
 * svgf_scene_load_textures decode every texture still without pixels from `<textures_dir>/<name>`: a JPEG decoder (baseline and
 * progressive, csrc/jpeg_decode.cpp) whose output is byte-identical to stb_image's for the reference's files. It returns how many
 * textures have pixels afterwards (svgf_scene_error says why one is missing). svgf_jpeg_decode_memory is the decoder on its
 * own: `out` may be NULL to ask for the size (width * height * components bytes). */
int svgf_scene_load_textures(svgf_scene *scene, const char *textures_dir);
int svgf_jpeg_decode_memory(const unsigned char *bytes, size_t n, int *width, int *height, int *components, unsigned char *out, size_t out_bytes);
int svgf_scene_num_textures(const svgf_scene *scene);
const char *svgf_scene_texture_file(const svgf_scene *scene, int index);
int svgf_scene_set_texture(svgf_scene *scene, int index, int width, int height, int components, const unsigned char *pixels);
/* Scene::BoudningBoxs (one {min, max} per MESH geom, scene.cpp:309; not read by the hot path). Returns their number. */
int svgf_scene_mesh_boxes(const svgf_scene *scene, float *out6, int max_boxes);

#ifdef __cplusplus
}
#endif
#endif /* SVGF_B200_H */
