// svgf_internal.h -- device data layout and context of the B200-native SVGF + path-trace hot path.
// (internal; the public boundary is include/svgf_b200.h)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/svgf_b200.h"

// ---------------------------------------------------------------------------------------------------
// HBM layout. Everything per-pixel is a 16-byte-element plane (float4), row-major, so that one thread
// moves one pixel with one LDG.128/STG.128 and TMA can tile it with 16 B elements.
//
//   cv[3]        {r, g, b, variance}      colour+variance; rotates between "accumulated", "a-trous ping/pong"
//                                         and "history" by pointer (no D2D copies, cf. denoise.cu:366,391,396-398)
//   nrm[2]       {nx, ny, nz, geomId}     current / previous frame (swap), geomId as int bits
//   pos          {px, py, pz, 0}
//   alb          {ar, ag, ab, 0}          first-hit albedo (re-modulated by the last a-trous level)
//   mom[2]       {m1, m2}                 luminance moments history / accumulated (swap)
//   hlen[2]      int32                    history length before / after back-projection (swap)
//   image        vec3 AoS (12 B)          1-spp radiance, the reference's dev_image layout (read by pbo pack)
//   denoised     vec3 AoS (12 B)          final colour, the reference's dev_denoised_image layout (D2H'd as is)
//   var_out      float                    final variance (what the reference leaves in dev_variance)
//   stale        {nx, ny, nz, matId}, {u, v}   the part of the reference's persistent per-pixel
//                                         ShadeableIntersection that survives a miss (pathtrace.cu:267-271,316-322)
// ---------------------------------------------------------------------------------------------------

struct GeomD {              // what the closest-hit loop needs of a Geom (sceneStructs.h:33-47), 256 B
    int type, materialid, tri_begin, tri_end;
    float translation[3], pad_;
    float inverseTransform[16], transform[16], invTranspose[16];
    // conservative world-space bounds of a cube/sphere (inflated by 1 % of the diagonal + 1e-3): a ray that misses them
    // cannot pass the reference's exact object-space test, so that test (2 mat4 x vec4, a normalize, 6 IEEE divides) is skipped
    float aabb_min[3], pad1_, aabb_max[3], pad2_;
};
static_assert(sizeof(GeomD) == 256, "GeomD");

struct TexD { int w, h, comp, pad; const unsigned char *px; };

struct DeviceScene {
    GeomD *geoms = nullptr;             int n_geoms = 0;
    svgf_material *materials = nullptr; int n_materials = 0;
    float4 *bvh = nullptr;              int n_nodes = 0;    // 2 x float4 per node: {min, count|axis<<16 as int}, {max, offset}
    float4 *tri_hot = nullptr;          int n_tris = 0;     // 3 x float4 per tri: {v0, id}, {e1, owner geom}, {e2, 0}
    float4 *tri_cold = nullptr;                             // 4 x float4 per tri: {n0,u0} {n1,v0} {n2,u1} {v1,u2,v2,0}
    TexD *textures = nullptr;           int n_textures = 0;
    std::vector<unsigned char *> tex_pixels;
};

enum { SVGF_MAX_LEVELS = 7, SVGF_MAX_RANKS = 8 };

// ---- multi-GPU: a frame is sharded by row strips, one process per GPU. Every rank allocates full-frame planes and
// owns the rows of its strip; rows of other strips are read IN PLACE from the owner's memory over NVLink (CUDA IPC
// mappings), so there is no halo copy and no collective on the data path: the a-trous tile loader and the temporal
// reprojection simply pick the owner's base pointer for each row they touch. Ordering between ranks is a per-stage
// sequence flag pushed into every peer's memory (signal kernel) and polled locally (wait kernel). ----
struct RowOwner {               // rows [start[r], start[r+1]) belong to rank r
    int world;
    int start[SVGF_MAX_RANKS + 1];
};
template <typename T> struct PeerPtr { T *p[SVGF_MAX_RANKS]; };
__host__ __device__ inline int owner_of(const RowOwner &ro, int y) {
    int r = 0;
    for (int i = 1; i < ro.world; i++) r += (y >= ro.start[i]);
    return r;
}
enum { SVGF_STAGE_RT = 0, SVGF_STAGE_TEMPORAL = 1, SVGF_STAGE_LEVEL0 = 1 /* + level */, SVGF_STAGE_FRAME = 9, SVGF_NUM_STAGES = 10 };
enum { SVGF_NCV = 4 };          // colour/variance buffers in rotation: accumulated, a-trous ping/pong, history of the PREVIOUS frame
                                // (never written while the frame that reads it is in flight, so only neighbours need to be waited for)
enum { SVGF_IPC_NBUF = 2 * SVGF_NCV + 9 };    // cv[4] lv[4] nrm[2] mom[2] hlen[2] gnp gzl flags
enum { SVGF_PAD_ROWS = 130, SVGF_PAD_PX = 2 << SVGF_MAX_LEVELS };   // zeroed padding behind the planes the TMA tile loads address: lattice
                                // extents round up to ceil(H/s)*s rows and ceil(W/2s)*2s columns (s <= 128), i.e. < (H + 128) * W + 256 pixels

// Sharded frames, push mode: the ranks whose strips lie within `reach` rows of mine. rows [lo[i], hi[i]) of MY strip are tapped
// by rank[i] at the next stage (its producer stores them into that rank's copy of the plane as well), and the same ranks are the
// ones whose rows I tap (the reach is symmetric), i.e. whose stage flag I have to see before reading my halo rows.
struct HaloPeers {
    int n;
    int rank[SVGF_MAX_RANKS - 1], lo[SVGF_MAX_RANKS - 1], hi[SVGF_MAX_RANKS - 1];
};
struct HaloOut {                // producer side of one stage; peers.n == 0: single GPU / nobody in reach
    HaloPeers peers;
    unsigned *flag[SVGF_MAX_RANKS - 1];     // where my flag of this stage lives in peers.rank[i]'s memory
    unsigned *counter;                      // blocks of this launch that have finished
    unsigned seq;
    int signal;                             // 1: the last block raises the flags (0: a separate signal kernel follows)
};

struct HaloIn {                 // consumer side: flags (in MY memory) that must have reached `seq` before halo rows are read
    int n;
    const unsigned *flag[SVGF_MAX_RANKS - 1];
    unsigned seq;
    unsigned *err;              // pinned, mapped error word (svgf_ctx::comm_err)
};


struct svgf_ctx {
    int device = 0;
    int W = 0, H = 0;
    size_t px = 0;
    cudaStream_t stream = nullptr;
    DeviceScene scene;

    // shard (rows [row_begin,row_end) of the frame); world==1: whole frame
    svgf_shard shard{0, 1, 0, 0};
    RowOwner rows{1, {0}};
    // peer views of every plane another rank may read (index [rank]; [shard.rank] is this context's own pointer)
    PeerPtr<float4> p_cv[SVGF_NCV]; PeerPtr<float2> p_lv[SVGF_NCV]; PeerPtr<float4> p_nrm[2]; PeerPtr<float2> p_mom[2]; PeerPtr<int> p_hlen[2];
    PeerPtr<float4> p_gnp; PeerPtr<float2> p_gzl; PeerPtr<unsigned> p_flags;
    unsigned *flags = nullptr;          // [SVGF_MAX_RANKS][SVGF_NUM_STAGES] sequence numbers written by the peers
    unsigned *done_count = nullptr;     // [SVGF_NUM_STAGES] blocks of a producer kernel that have finished (the last one raises the flags)
    unsigned *comm_err = nullptr;       // pinned, mapped: set by a wait that gave up (a peer stopped making progress); device alias below
    unsigned *comm_err_dev = nullptr;
    unsigned seq = 0;                   // frame sequence number (identical on all ranks)
    std::vector<void *> ipc_opened;

    float4 *cv[SVGF_NCV] = {nullptr, nullptr, nullptr, nullptr};
    // {luminance of cv[i].rgb in the reference's fp64 formula (denoise.cu:121), copy of cv[i].w}: what a tap needs besides
    // colour, and what the 3x3 variance blur reads (8 B/px, coalesced)
    float2 *lv[SVGF_NCV] = {nullptr, nullptr, nullptr, nullptr};
    // TMA descriptors (CUtensorMap, 128 B each) for the lattice tiles of cv[3], lv[3], gnp, gzl: [plane 8][level 1..7][shape 2]
    void *tmaps = nullptr; int tma_ok = 0;
    void *tmaps_slide = nullptr;        // ... and of the sliding kernel: [plane][level]
    // sharded frames: 1 = every stage pushes the rows its neighbours will tap into their copy of the plane, levels read local
    // memory only (TMA tiles everywhere); 0 = levels read neighbours' rows in place over NVLink (SVGF_HALO=pull, A/B)
    int halo_push = 1;
    int atrous_variant = 2;             // 1 = direct (one thread per pixel), 2 = lattice-tiled (TMA tile loads), 3 = lattice-tiled (cp.async),
                                        // 4 = lattice-tiled, symmetric two-phase pair arithmetic (atrous_pair_core.h),
                                        // 5 = sliding warps, symmetric pair arithmetic in registers (atrous_slide_core.h)
    bool atrous_attr_set = false, atrous_pair_attr_set = false, atrous_slide_attr_set = false;
    int atrous_slide_blocks_per_sm = 0, atrous_slide_sms = 148;
    int atrous_slide_bands = 0;         // SVGF_ATROUS_BANDS: row bands per (class, strip) of the sliding kernel (0 = fill the warp slots once)
    // tile shape of the lattice-tiled kernel (index into atrous.cu's table): -1 = chosen per level by the cost model;
    // SVGF_ATROUS_SHAPE=<id> forces one for every level, SVGF_ATROUS_SHAPES=<id>,<id>,... one per level (A/B runs)
    int atrous_pair_rows = 2;           // SVGF_ATROUS_PAIR_ROWS: centre rows per thread in phase 2 of the pair kernel (1 = 256-thread blocks)
    int atrous_probe = 0;               // SVGF_ATROUS_PROBE: timing probes of the tiled kernel (results are garbage)
    int atrous_shape = -1, atrous_shape_level[SVGF_MAX_LEVELS + 1] = {-1, -1, -1, -1, -1, -1, -1, -1};
    int rt_compact = 0;                 // A/B (SVGF_RT_COMPACT): 1 = compacting rt_kernel for scenes with next to no mesh, 2 = always (see rt_kernel, CP)
    int rt_variant = 0;                 // 0 = state machine, one pixel per thread (default), 1 = wavefront (stage kernels +
                                        // ballot-compacted queues), 2 = persistent state machine with work refill
    unsigned int *rt_counter = nullptr; int rt_blocks = 0;
    void *wf_mem = nullptr;             // wavefront ray/hit/queue buffers (allocated on first use)
    int hist_cv = -1;                   // which cv[] holds the colour history for the next frame (-1: none yet)
    float4 *nrm[2] = {nullptr, nullptr};
    int cur_nrm = 0, gbuf_nrm = 0;
    float4 *pos = nullptr, *alb = nullptr;
    // a-trous view of the G-buffer, pre-scaled by the edge-stopping constants kn = log2(e)/(sigma_n + 1e-6) and
    // kx = log2(e)/(sigma_x + 1e-6) and interleaved for the packed-fp32 distance code: {kn nx, kx px, kn ny, kx py}, {kn nz, kx pz}
    float4 *gnp = nullptr; float2 *gzl = nullptr;
    float2 *mom[2] = {nullptr, nullptr};
    int cur_mom = 0;                    // mom[cur_mom] = history, the other = accumulated
    int *hlen[2] = {nullptr, nullptr};
    int cur_hlen = 0;
    float *image = nullptr, *denoised = nullptr, *var_out = nullptr;
    float4 *stale_nm = nullptr; float2 *stale_uv = nullptr;
    float *kl = nullptr;                // per-level scratch: luminance-weight scale per pixel (atrous.cu)
    unsigned char *pbo_own = nullptr;   // used when the caller passes no PBO

    // scratch for the AoS entry point svgf_denoise() and the host conveniences
    float *aos_in = nullptr, *aos_out = nullptr; svgf_gbuffer_texel *aos_g = nullptr;
    float *pinned_image = nullptr;      // W*H*3 pinned staging for the per-frame D2H
    // svgf_render_async: the image of frame N is copied out on copy_stream while frame N+1 renders; `denoised` and
    // `denoised_alt` swap every async frame, copy_done[slot] gates the reuse of a buffer two frames later
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t frame_done = nullptr, copy_done[2] = {nullptr, nullptr};
    float *denoised_alt = nullptr;
    int copy_slot = 0, copy_pending[2] = {0, 0};
    const float *copy_host[2] = {nullptr, nullptr};

    // quality switches the reference leaves as TODOs; all off by default (parity), svgf_set_option
    int opt_reprojection_fov_aspect = 0;    // 1: the back-projection honours FOV and aspect ratio (denoise.cu:200-207 does not)
    int opt_history_cap = 0;                // > 0: history length saturates there (unbounded in the reference)
    // Can a normal or position of this frame's G-buffer be NaN? Decided from the scene at upload (a mesh whose vertex normals can
    // interpolate to zero shades with normalize(0), sceneStructs.h:168-172; a non-finite matrix), always true for G-buffers handed
    // in through svgf_denoise. false: the a-trous tile kernel drops the NaN guard of its distances (atrous_tile_core.h: dist_of).
    bool scene_nan_possible = true, gbuf_nan_possible = true;
    int atrous_fused = 0;                   // SVGF_ATROUS_FUSED=1 (A/B): the a-trous stage as one launch (atrous_stage_kernel); measured slower, see atrous.cu
    unsigned *stage_ctr = nullptr; int stage_blocks = 0; bool stage_attr_set = false;
    // Cross-frame overlap (single-GPU frames; option "frame_overlap" / SVGF_FRAME_OVERLAP=1, OFF by default): the path tracer of
    // frame N + 1 runs on its own stream next to the a-trous stage of frame N. Bit-identical frames. Measured on B200
    // (profiles/r2_ab_frame_overlap.txt): frames queued back to back gain 2-6 % (C2 745 -> 761 fps, C5 550 -> 583), but a caller
    // that waits for every image (the pipelined e2e loop of bench.py) LOSES 4-19 % (C2 740 -> 714, C3 297 -> 242): the path
    // tracer's blocks live ~58 us, so each of frame N's eleven denoise launches finds the SMs full of them and ramps up slowly,
    // and with a-trous blocks resident the SM's carve-out is all shared memory, which costs the path tracer its L1. What the path tracer writes and the rest of the previous frame still reads
    // exists twice (image, gnp, gzl, alb: `*_set`; c->image / gnp / gzl / alb are aliases of the current frame's set); everything
    // else it writes (normals of two frames ago, positions) is last read by the previous frame's temporal pass, which it waits for.
    int frame_overlap = 0, gset = 0;
    float *image_set[2] = {nullptr, nullptr}; float4 *gnp_set[2] = {nullptr, nullptr}, *alb_set[2] = {nullptr, nullptr}; float2 *gzl_set[2] = {nullptr, nullptr};
    cudaStream_t rt_stream = nullptr, rt_launch_stream = nullptr;      // rt_launch_stream: where launch_pathtrace puts its kernels (null: c->stream)
    cudaEvent_t ev_rt_done = nullptr, ev_temporal_done = nullptr; bool temporal_done_valid = false;
    int *slot_of_input = nullptr, *bvh_parent = nullptr; void *refit_stage = nullptr;      // svgf_refit_bvh (lbvh.cu): where input triangle i lives, parents + flags of the tree in place, upload staging
    int halo_copy_from_rows = 32;           // sharded frames: levels whose next-level halo is at least this many rows push it with the copy kernel
    int opt_cuda_graph = 0;                 // 1: the frame's launches run as one CUDA graph, updated in place every frame (N1)
    cudaGraphExec_t graph_exec = nullptr; int frames_rendered = 0;
    cudaEvent_t legacy_fence = nullptr;     // svgf_denoise: orders the library's stream after the caller's legacy default stream
    int opt_light_sampling_all = 0;         // 1: shadow rays sample every emissive cube/sphere (the reference: geoms[0] only)
    int opt_spatial_variance = 0;           // 1: pixels with a history shorter than 4 frames estimate their variance spatially
    int n_lights = 0, lights[8] = {0};      // emissive cubes/spheres of the uploaded scene

    float view_matrix_prev[16];         // denoise.cu:15; identity until the first denoise (glm::mat4())
    int last_variance_valid = 0;        // var_out holds the final variance of the last frame

    // profiling: a pool of per-frame event sets so that a whole timed region can be measured without a
    // host synchronisation per frame; svgf_stage_times() averages over the frames recorded since it was enabled
    int profiling = 0;
    struct ProfFrame { cudaEvent_t ev[12]; int nlevel; int denoise; };
    std::vector<ProfFrame> prof_pool;
    int prof_count = 0;
    float stage_ms[11] = {};

    // host buffers the caller passed as host_image, page-locked once so the per-frame D2H is a direct DMA
    std::vector<std::pair<void *, size_t>> registered_hosts;

    std::string err;
};

// ---- helpers ---------------------------------------------------------------------------------------
#define SVGF_CUDA(ctx, call)                                                                          \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                          \
            return SVGF_ERR_CUDA;                                                                     \
        }                                                                                             \
    } while (0)

// kernels / stage launchers (each returns a cudaError_t from the launch)
struct RtParams {
    int W, H, row_begin, row_end;       // rows this launch covers
    int frame, max_depth;
    int trace_shadowray, reduce_var, denoise, sepcolor;
    float sintensity, lightradius;
    float kn, kx;                       // a-trous edge-stopping scales for the pre-scaled G-buffer planes
    int n_lights, lights[8];            // "light_sampling_all": the emissive cubes/spheres (n_lights <= 1: the reference's geoms[0])
    svgf_camera cam;
};
void atrous_scales(float sigma_n, float sigma_x, float *kn, float *kx);
cudaError_t launch_pathtrace(svgf_ctx *c, const RtParams &p, float4 *nrm_out, int gbuf_reach, bool *pushed);
cudaError_t launch_temporal(svgf_ctx *c, const float *image, const float4 *nrm_cur, const PeerPtr<float4> &nrm_prev,
                            const float4 *pos, const PeerPtr<float4> &hist_cv, const PeerPtr<float2> &mom_hist,
                            const PeerPtr<int> &hlen_in, float4 *acc_cv, float2 *acc_lv, float2 *mom_acc, int *hlen_out,
                            const float *prev_viewmat, float color_alpha, float moment_alpha, float clip_rx, float clip_ry,
                            const HaloOut &ho, const PeerPtr<float4> &acc_cv_peers, const PeerPtr<float2> &acc_lv_peers);
HaloPeers halo_peers(const svgf_ctx *c, int reach);
HaloOut halo_out(svgf_ctx *c, int stage, int reach, bool fused_signal);
HaloIn halo_in(svgf_ctx *c, int stage, int reach, unsigned seq);
struct HaloPlane { const void *local; void *peer[SVGF_MAX_RANKS]; int esz; };     // esz = bytes per pixel (multiple of 8)
cudaError_t launch_halo_push(svgf_ctx *c, int halo_rows, const HaloPlane *planes, int nplanes, const HaloOut *sig = nullptr);
cudaError_t launch_signal(svgf_ctx *c, int stage, int reach);                 // reach < 0: every connected rank
cudaError_t launch_wait(svgf_ctx *c, int stage, unsigned seq, int reach);
cudaError_t launch_no_temporal(svgf_ctx *c, const float *image, float4 *acc_cv, float2 *acc_lv);
cudaError_t launch_spatial_variance(svgf_ctx *c, const int *hlen, const float2 *mom, const float4 *nrm, float4 *cv, float2 *lv);
struct AtrousArgs {
    int src_slot;                                   // cv/lum input = c->p_cv[src_slot] / c->p_lum[src_slot] (peer-readable)
    const float4 *cv_in; float4 *cv_out;            // cv_out may be null on the last level
    const float2 *lv_in; float2 *lv_out;            // {luminance, variance} planes; slot numbers select the TMA descriptors
    int dst_slot;
    const float4 *nrm, *pos, *alb;
    const float4 *gnp; const float2 *gzl;
    float *denoised_out; float *var_out;            // last level only (AoS vec3 + float plane)
    int level, is_last, blur_variance, addcolor;
    float sigma_c, sigma_n, sigma_x;
    int gset = 0;                                   // which G-buffer set gnp/gzl belong to (selects their TMA descriptors)
    HaloIn wait{};                                  // sharded frames: flags to see before halo rows are read (n == 0: none)
    HaloOut ho{};                                   // ... and who gets this level's edge rows and its flag
};
cudaError_t launch_atrous(svgf_ctx *c, const AtrousArgs &a);
bool atrous_stage_possible(const svgf_ctx *c, const AtrousArgs *a, int n);         // all levels of the stage in one launch?
cudaError_t launch_atrous_stage(svgf_ctx *c, const AtrousArgs *a, int n);
cudaError_t launch_cv_to_outputs(svgf_ctx *c, const float4 *cv, float *denoised, float *var_out);
cudaError_t launch_debug_view(svgf_ctx *c, int option, const int *hlen, const float4 *cv, float *denoised);
cudaError_t launch_pack_pbo(svgf_ctx *c, unsigned char *pbo, const float *left, const float *right);
cudaError_t launch_aos_to_soa(svgf_ctx *c, const svgf_gbuffer_texel *g, float4 *nrm, float4 *pos, float4 *alb, float kn, float kx);
cudaError_t launch_soa_to_aos(svgf_ctx *c, const float4 *nrm, const float4 *pos, const float4 *alb, svgf_gbuffer_texel *g);
cudaError_t launch_copy_f3(svgf_ctx *c, float *dst, const float *src);

int atrous_build_tensor_maps(svgf_ctx *c);
void preload_pathtrace_kernels(); void preload_denoise_kernels(); void preload_atrous_kernels();
void svgf_view_matrix(const svgf_camera *cam, float *out16);
