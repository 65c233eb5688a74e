// api.cu -- the C ABI (include/svgf_b200.h): context lifecycle, scene upload, the per-frame driver.
//
// Frame driver = pathtrace() (src/pathtrace.cu:404-452) + denoise() (src/denoise.cu:349-402) of the reference,
// re-plumbed for one CUDA stream with no host synchronisation inside the frame:
//   * the reference's 6-7 device-to-device copies per frame (176 B/pixel) are replaced by rotating which buffer
//     plays "history" / "accumulated" / "ping" / "pong";
//   * its 4 cudaDeviceSynchronize + pageable D2H are replaced by one async copy into pinned memory and a single
//     stream synchronise, only when the caller asked for the host image.
// There is no CPU fallback anywhere in this file: every entry point needs a CUDA device.
#include "svgf_internal.h"

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>

static thread_local std::string g_create_err;

static_assert(sizeof(svgf_geom) == 248 && sizeof(svgf_material) == 56 && sizeof(svgf_triangle) == 136 &&
              sizeof(svgf_bvh_node) == 40 && sizeof(svgf_camera) == 84 && sizeof(svgf_gbuffer_texel) == 52 &&
              sizeof(svgf_path_segment) == 48 && sizeof(svgf_intersection) == 36,
              "reference ABI sizes (SURVEY.md 8(a) T1-T10)");
static_assert(offsetof(svgf_geom, transform) == 44 && offsetof(svgf_geom, inverseTransform) == 108 &&
              offsetof(svgf_geom, invTranspose) == 172 && offsetof(svgf_geom, T_startidx) == 236, "Geom offsets");
static_assert(offsetof(svgf_material, specular_color) == 16 && offsetof(svgf_material, hasReflective) == 28 &&
              offsetof(svgf_material, emittance) == 40 && offsetof(svgf_material, texid) == 48, "Material offsets");
static_assert(offsetof(svgf_triangle, verts) == 4 && offsetof(svgf_triangle, normal) == 100, "Triangle offsets");
static_assert(offsetof(svgf_bvh_node, primitive_count) == 24 && offsetof(svgf_bvh_node, rightchildoffset) == 36, "BVH_ArrNode offsets");
static_assert(offsetof(svgf_camera, position) == 8 && offsetof(svgf_camera, right) == 56 && offsetof(svgf_camera, pixelLength) == 76, "Camera offsets");
static_assert(offsetof(svgf_gbuffer_texel, position) == 12 && offsetof(svgf_gbuffer_texel, geomId) == 48, "GBufferTexel offsets");
static_assert(offsetof(svgf_intersection, materialId) == 16 && offsetof(svgf_intersection, uv) == 28, "ShadeableIntersection offsets");
static_assert(offsetof(svgf_path_segment, pixelIndex) == 36 && offsetof(svgf_path_segment, diffuse) == 44, "PathSegment offsets");

#define CK(call)                                                                    \
    do {                                                                            \
        cudaError_t e_ = (call);                                                    \
        if (e_ != cudaSuccess) {                                                    \
            c->err = std::string(#call) + ": " + cudaGetErrorString(e_);            \
            return SVGF_ERR_CUDA;                                                   \
        }                                                                           \
    } while (0)

template <typename T> static cudaError_t dalloc(T **p, size_t n) { return cudaMalloc((void **)p, n ? n * sizeof(T) : sizeof(T)); }

static int upload_scene(svgf_ctx *c, const svgf_scene_desc *d) {
    DeviceScene &s = c->scene;
    // geoms -> GeomD
    std::vector<GeomD> gd(d->n_geoms);
    for (int i = 0; i < d->n_geoms; i++) {
        const svgf_geom &g = d->geoms[i];
        GeomD &o = gd[i];
        memset(&o, 0, sizeof(o));
        o.type = g.type; o.materialid = g.materialid;
        o.tri_begin = g.type == 2 ? g.T_startidx : 0; o.tri_end = g.type == 2 ? g.T_endidx : 0;
        memcpy(o.translation, g.translation, 12);
        memcpy(o.inverseTransform, g.inverseTransform, 64); memcpy(o.transform, g.transform, 64); memcpy(o.invTranspose, g.invTranspose, 64);
        if (g.materialid < 0 || g.materialid >= d->n_materials) { c->err = "geom references a material out of range"; return SVGF_ERR_INVALID; }
        {   // conservative world bounds (fp64): cube = hull of the 8 transformed corners; sphere (r = 0.5) = centre +- 0.5 * |row_i of M3x3|
            double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
            const float *M = g.transform;       // column-major
            if (g.type == 0) {
                for (int i = 0; i < 3; i++) {
                    const double e = 0.5 * sqrt((double)M[0 + i] * M[0 + i] + (double)M[4 + i] * M[4 + i] + (double)M[8 + i] * M[8 + i]);
                    lo[i] = M[12 + i] - e; hi[i] = M[12 + i] + e;
                }
            } else {
                for (int k = 0; k < 8; k++) {
                    const double x = (k & 1) ? 0.5 : -0.5, y = (k & 2) ? 0.5 : -0.5, z = (k & 4) ? 0.5 : -0.5;
                    for (int i = 0; i < 3; i++) {
                        const double v = M[0 + i] * x + M[4 + i] * y + M[8 + i] * z + M[12 + i];
                        if (v < lo[i]) lo[i] = v;
                        if (v > hi[i]) hi[i] = v;
                    }
                }
            }
            const double diag = sqrt((hi[0] - lo[0]) * (hi[0] - lo[0]) + (hi[1] - lo[1]) * (hi[1] - lo[1]) + (hi[2] - lo[2]) * (hi[2] - lo[2]));
            // Padding: the exact tests work in fp32 object space, whose rounding moves a silhouette by a few ulps of the world
            // coordinates (~1e-6 at |x| ~ 10), so ~50 ulps are ample. It must stay well below the 1e-4 by which bounce and
            // shadow rays are lifted off the surface they leave (pathtrace.cu:366, interactions.h:103): then a ray leaving a
            // wall starts OUTSIDE the wall's bounds and is rejected by the slab test instead of costing an exact test.
            double maxabs = 0.0;
            for (int i = 0; i < 3; i++) { maxabs = std::max(maxabs, fabs(lo[i])); maxabs = std::max(maxabs, fabs(hi[i])); }
            const double pad = 2e-6 * (1.0 + maxabs + diag);
            for (int i = 0; i < 3; i++) { o.aabb_min[i] = (float)(lo[i] - pad); o.aabb_max[i] = (float)(hi[i] + pad); }
            if (!(diag == diag) || g.type == 2) for (int i = 0; i < 3; i++) { o.aabb_min[i] = -3e38f; o.aabb_max[i] = 3e38f; }
        }
    }
    {   // May a shadow query stop at the first occluder (pathtrace.cu: computeIntersection, light_query)? The answer "is geoms[0] the
        // closest hit" is unchanged by that iff geoms[0] is a cube or a sphere and every triangle a traversal can return belongs to
        // some MESH geom's id range (the closest-hit search ignores a closest triangle nobody owns). SVGF_RT_ANYHIT=0: A/B switch.
        bool ok = d->geoms[0].type != 2 && !(getenv("SVGF_RT_ANYHIT") && atoi(getenv("SVGF_RT_ANYHIT")) == 0);
        for (int i = 0; ok && i < d->n_triangles; i++) {
            bool owned = false;
            for (int g = 0; g < d->n_geoms && !owned; g++)
                owned = d->geoms[g].type == 2 && d->triangles[i].id >= d->geoms[g].T_startidx && d->triangles[i].id < d->geoms[g].T_endidx;
            ok = owned;
        }
        gd[0].pad_ = ok ? 1.0f : 0.0f;
        c->n_lights = 0;
        for (int g = 0; g < d->n_geoms && g < 32 && c->n_lights < 8; g++)
            if (d->geoms[g].type != 2 && d->materials[d->geoms[g].materialid].emittance > 0.0f) c->lights[c->n_lights++] = g;
    }
    s.n_geoms = d->n_geoms; s.n_materials = d->n_materials; s.n_nodes = d->n_bvh_nodes; s.n_tris = d->n_triangles; s.n_textures = d->n_textures;
    CK(dalloc(&s.geoms, gd.size()));
    CK(cudaMemcpy(s.geoms, gd.data(), gd.size() * sizeof(GeomD), cudaMemcpyHostToDevice));
    for (int i = 0; i < d->n_materials; i++)
        if (d->materials[i].texid != -1 && (d->materials[i].texid < 0 || d->materials[i].texid >= d->n_textures)) {
            c->err = "material references a texture out of range"; return SVGF_ERR_INVALID;
        }
    CK(dalloc(&s.materials, (size_t)d->n_materials));
    CK(cudaMemcpy(s.materials, d->materials, (size_t)d->n_materials * sizeof(svgf_material), cudaMemcpyHostToDevice));
    // BVH nodes 40 B -> 2 x float4
    std::vector<float4> bv(2 * (size_t)d->n_bvh_nodes);
    for (int i = 0; i < d->n_bvh_nodes; i++) {
        const svgf_bvh_node &n = d->bvh_nodes[i];
        const bool leaf = n.primitive_count > 0;
        if (leaf && (n.primitive_count > 0xffff || n.primitivesOffset < 0 || n.primitivesOffset + n.primitive_count > d->n_triangles)) {
            c->err = "BVH leaf out of range"; return SVGF_ERR_INVALID;
        }
        if (!leaf && (n.axis < 0 || n.axis > 2 || n.rightchildoffset <= i || n.rightchildoffset >= d->n_bvh_nodes || i + 1 >= d->n_bvh_nodes)) {
            c->err = "BVH interior node malformed"; return SVGF_ERR_INVALID;
        }
        int meta = leaf ? n.primitive_count : (n.axis << 16);
        int off = leaf ? n.primitivesOffset : n.rightchildoffset;
        float fm, fo; memcpy(&fm, &meta, 4); memcpy(&fo, &off, 4);
        bv[2 * i] = make_float4(n.bounds_min[0], n.bounds_min[1], n.bounds_min[2], fm);
        bv[2 * i + 1] = make_float4(n.bounds_max[0], n.bounds_max[1], n.bounds_max[2], fo);
    }
    CK(dalloc(&s.bvh, bv.size()));
    if (!bv.empty()) CK(cudaMemcpy(s.bvh, bv.data(), bv.size() * sizeof(float4), cudaMemcpyHostToDevice));
    // triangles 136 B -> hot {v0,id} {e1} {e2} + cold {n0,u0} {n1,v0} {n2,u1} {v1,u2,v2}
    std::vector<float4> hot(3 * (size_t)d->n_triangles), cold(4 * (size_t)d->n_triangles);
    for (int i = 0; i < d->n_triangles; i++) {
        const svgf_triangle &t = d->triangles[i];
        const float *v0 = t.verts[0].pos, *v1 = t.verts[1].pos, *v2 = t.verts[2].pos;
        float fid; memcpy(&fid, &t.id, 4);
        // e1 = v1 - v0, e2 = v2 - v0 exactly as glm::intersectRayTriangle forms them (gtx/intersect.inl:44-45);
        // an fp32 subtraction gives the same bits on the host as on the device
        hot[3 * i] = make_float4(v0[0], v0[1], v0[2], fid);
        hot[3 * i + 1] = make_float4(v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2], 0.f);
        hot[3 * i + 2] = make_float4(v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2], 0.f);
        const float *n0 = t.verts[0].normal, *n1 = t.verts[1].normal, *n2 = t.verts[2].normal;
        cold[4 * i] = make_float4(n0[0], n0[1], n0[2], t.verts[0].uv[0]);
        cold[4 * i + 1] = make_float4(n1[0], n1[1], n1[2], t.verts[0].uv[1]);
        cold[4 * i + 2] = make_float4(n2[0], n2[1], n2[2], t.verts[1].uv[0]);
        cold[4 * i + 3] = make_float4(t.verts[1].uv[1], t.verts[2].uv[0], t.verts[2].uv[1], 0.f);
    }
    {   // NaN normals: normalize(b0 n0 + b1 n1 + b2 n2) of a triangle is NaN when the combination can vanish or is not finite. With
        // all pairwise dot products of the three vertex normals positive no convex combination vanishes; anything else is "possible".
        bool nan_possible = getenv("SVGF_ATROUS_NAN_GUARD") && atoi(getenv("SVGF_ATROUS_NAN_GUARD")) != 0;      // A/B: force the guard
        for (int i = 0; i < d->n_triangles && !nan_possible; i++) {
            const float *n[3] = {d->triangles[i].verts[0].normal, d->triangles[i].verts[1].normal, d->triangles[i].verts[2].normal};
            for (int a = 0; a < 3; a++) {
                const int b = (a + 1) % 3;
                const float dot = n[a][0] * n[b][0] + n[a][1] * n[b][1] + n[a][2] * n[b][2];
                if (!(dot > 0.0f) || !std::isfinite(dot)) nan_possible = true;
            }
        }
        for (int g = 0; g < d->n_geoms && !nan_possible; g++)
            for (int e = 0; e < 16; e++)
                if (!std::isfinite(d->geoms[g].transform[e]) || !std::isfinite(d->geoms[g].inverseTransform[e]) || !std::isfinite(d->geoms[g].invTranspose[e])) nan_possible = true;
        c->scene_nan_possible = nan_possible;
    }
    CK(dalloc(&s.tri_hot, hot.size())); CK(dalloc(&s.tri_cold, cold.size()));
    if (!hot.empty()) {
        CK(cudaMemcpy(s.tri_hot, hot.data(), hot.size() * sizeof(float4), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(s.tri_cold, cold.data(), cold.size() * sizeof(float4), cudaMemcpyHostToDevice));
    }
    // textures
    std::vector<TexD> td(d->n_textures);
    for (int i = 0; i < d->n_textures; i++) {
        const svgf_texture_desc &t = d->textures[i];
        if (t.width <= 0 || t.height <= 0 || t.components <= 0 || !t.pixels) {
            c->err = "texture without pixels"; return SVGF_ERR_INVALID;
        }
        const size_t n = (size_t)t.width * t.height * t.components;
        unsigned char *dp = nullptr;
        CK(cudaMalloc((void **)&dp, n ? n : 1));
        s.tex_pixels.push_back(dp);
        if (n) CK(cudaMemcpy(dp, t.pixels, n, cudaMemcpyHostToDevice));
        td[i].w = t.width; td[i].h = t.height; td[i].comp = t.components; td[i].pad = 0; td[i].px = dp;
    }
    CK(dalloc(&s.textures, td.size()));
    if (!td.empty()) CK(cudaMemcpy(s.textures, td.data(), td.size() * sizeof(TexD), cudaMemcpyHostToDevice));
    return SVGF_OK;
}

// The planes another rank may read, in the fixed order used for IPC export (SVGF_IPC_NBUF entries).
static void shared_bufs(svgf_ctx *c, void **o) {
    int n = 0;
    for (int i = 0; i < SVGF_NCV; i++) o[n++] = c->cv[i];
    for (int i = 0; i < SVGF_NCV; i++) o[n++] = c->lv[i];
    for (int i = 0; i < 2; i++) o[n++] = c->nrm[i];
    for (int i = 0; i < 2; i++) o[n++] = c->mom[i];
    for (int i = 0; i < 2; i++) o[n++] = c->hlen[i];
    o[n++] = c->gnp; o[n++] = c->gzl; o[n++] = c->flags;
}
static void set_peer(svgf_ctx *c, int r, void *const *b) {
    int n = 0;
    for (int i = 0; i < SVGF_NCV; i++) c->p_cv[i].p[r] = (float4 *)b[n++];
    for (int i = 0; i < SVGF_NCV; i++) c->p_lv[i].p[r] = (float2 *)b[n++];
    for (int i = 0; i < 2; i++) c->p_nrm[i].p[r] = (float4 *)b[n++];
    for (int i = 0; i < 2; i++) c->p_mom[i].p[r] = (float2 *)b[n++];
    for (int i = 0; i < 2; i++) c->p_hlen[i].p[r] = (int *)b[n++];
    c->p_gnp.p[r] = (float4 *)b[n++]; c->p_gzl.p[r] = (float2 *)b[n++]; c->p_flags.p[r] = (unsigned *)b[n++];
}

static int alloc_frame_buffers(svgf_ctx *c) {
    const size_t px = c->px;
    // planes the TMA tile loader addresses get SVGF_PAD_ROWS rows of (zeroed, never written) padding: lattice extents round up
    const size_t ppx = px + (size_t)SVGF_PAD_ROWS * c->W + SVGF_PAD_PX;
    for (int i = 0; i < SVGF_NCV; i++) {
        CK(dalloc(&c->cv[i], ppx)); CK(dalloc(&c->lv[i], ppx));
        CK(cudaMemset(c->cv[i], 0, ppx * sizeof(float4))); CK(cudaMemset(c->lv[i], 0, ppx * sizeof(float2)));
    }
    for (int i = 0; i < 2; i++) { CK(dalloc(&c->nrm[i], px)); CK(dalloc(&c->mom[i], px)); CK(dalloc(&c->hlen[i], px)); }
    CK(dalloc(&c->pos, px));
    for (int g = 0; g < 2; g++) {       // two sets of what the path tracer writes and the rest of the previous frame still reads (svgf_internal.h)
        CK(dalloc(&c->alb_set[g], px)); CK(dalloc(&c->gnp_set[g], ppx)); CK(dalloc(&c->gzl_set[g], ppx)); CK(dalloc(&c->image_set[g], 3 * px));
        CK(cudaMemset(c->gnp_set[g], 0, ppx * sizeof(float4))); CK(cudaMemset(c->gzl_set[g], 0, ppx * sizeof(float2)));
    }
    c->gset = 0; c->alb = c->alb_set[0]; c->gnp = c->gnp_set[0]; c->gzl = c->gzl_set[0]; c->image = c->image_set[0];
    CK(dalloc(&c->denoised, 3 * px)); CK(dalloc(&c->var_out, px));
    CK(dalloc(&c->stale_nm, px)); CK(dalloc(&c->stale_uv, px)); CK(dalloc(&c->kl, px));
    CK(cudaMalloc((void **)&c->pbo_own, px * 8));
    CK(cudaMalloc((void **)&c->flags, SVGF_MAX_RANKS * SVGF_NUM_STAGES * sizeof(unsigned)));
    CK(cudaMemset(c->flags, 0, SVGF_MAX_RANKS * SVGF_NUM_STAGES * sizeof(unsigned)));
    CK(cudaMalloc((void **)&c->done_count, SVGF_NUM_STAGES * sizeof(unsigned)));
    CK(cudaMemset(c->done_count, 0, SVGF_NUM_STAGES * sizeof(unsigned)));
    // the "a cross-rank wait gave up" word lives in mapped host memory: kernels set it, the host reads it without a copy
    CK(cudaHostAlloc((void **)&c->comm_err, sizeof(unsigned), cudaHostAllocMapped));
    *c->comm_err = 0u;
    CK(cudaHostGetDevicePointer((void **)&c->comm_err_dev, c->comm_err, 0));
    {   // until peers are connected every table entry is this context's own plane
        void *own[SVGF_IPC_NBUF]; shared_bufs(c, own);
        for (int r = 0; r < SVGF_MAX_RANKS; r++) set_peer(c, r, own);
        c->rows.world = 1; c->rows.start[0] = 0; for (int r = 1; r <= SVGF_MAX_RANKS; r++) c->rows.start[r] = c->H;
    }
    CK(cudaMallocHost((void **)&c->pinned_image, px * 12));
    atrous_build_tensor_maps(c);        // c->tma_ok = 0 (cp.async loader) when the driver entry point or the row pitch does not allow it
    return SVGF_OK;
}

extern "C" {

int svgf_abi_version(void) { return SVGF_ABI_VERSION; }

void svgf_params_default(svgf_params *p) {
    memset(p, 0, sizeof(*p));
    p->tracedepth = 4; p->shadowray = 1; p->reducevar = 1; p->sintensity = 2.7f; p->lightradius = 1.4f;
    p->denoise_enable = 1; p->sepcolor = 1; p->temporal_enable = 1; p->color_alpha = 0.2f; p->moment_alpha = 0.2f;
    p->right_view_option = 0; p->atrous_nlevel = 5; p->spatial_enable = 1; p->history_level = 1;
    p->sigmal = 0.45f; p->sigman = 0.2f; p->sigmax = 0.35f; p->blurvariance = 1; p->addcolor = 1;
}

const char *svgf_last_error(const svgf_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int svgf_create(svgf_ctx **out, const svgf_scene_desc *scene, int device) {
    if (!out || !scene || scene->width <= 0 || scene->height <= 0 || scene->n_geoms <= 0 || scene->n_materials <= 0 ||
        !scene->geoms || !scene->materials || scene->n_triangles < 0 || scene->n_bvh_nodes < 0 || scene->n_textures < 0 ||
        (scene->n_triangles > 0 && !scene->triangles) || (scene->n_bvh_nodes > 0 && !scene->bvh_nodes) ||
        (scene->n_textures > 0 && !scene->textures)) {
        g_create_err = "svgf_create: invalid scene description";
        return SVGF_ERR_INVALID;
    }
    if (out) *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        g_create_err = "svgf_create: no usable CUDA device (this library has no CPU path)";
        (void)cudaGetLastError();
        return SVGF_ERR_NO_DEVICE;
    }
    svgf_ctx *c = new (std::nothrow) svgf_ctx();
    if (!c) return SVGF_ERR_INVALID;
    c->device = device; c->W = scene->width; c->H = scene->height; c->px = (size_t)c->W * c->H;
    c->shard = svgf_shard{0, 1, 0, c->H};
    if (const char *v = getenv("SVGF_RT_VARIANT")) c->rt_variant = (!strcmp(v, "wavefront") || !strcmp(v, "1")) ? 1 : ((!strcmp(v, "persistent") || !strcmp(v, "2")) ? 2 : 0);
    if (const char *v = getenv("SVGF_RT_COMPACT")) c->rt_compact = atoi(v);       // A/B testing
    if (const char *v = getenv("SVGF_HALO")) c->halo_push = strcmp(v, "pull") != 0;      // A/B testing
    if (const char *v = getenv("SVGF_ATROUS_VARIANT")) c->atrous_variant = (atoi(v) == 1 || atoi(v) == 3 || atoi(v) == 4 || atoi(v) == 5) ? atoi(v) : 2;    // A/B testing
    if (const char *v = getenv("SVGF_ATROUS_BANDS")) c->atrous_slide_bands = atoi(v);
    if (const char *v = getenv("SVGF_CUDA_GRAPH")) c->opt_cuda_graph = atoi(v) != 0;
    if (const char *v = getenv("SVGF_FRAME_OVERLAP")) c->frame_overlap = atoi(v) != 0;      // A/B testing
    if (const char *v = getenv("SVGF_HALO_COPY_FROM")) c->halo_copy_from_rows = atoi(v);        // A/B testing
    if (const char *v = getenv("SVGF_ATROUS_FUSED")) c->atrous_fused = atoi(v) != 0;        // A/B testing
    if (const char *v = getenv("SVGF_ATROUS_SHAPE")) c->atrous_shape = atoi(v);
    if (const char *v = getenv("SVGF_ATROUS_PROBE")) c->atrous_probe = atoi(v);
    if (const char *v = getenv("SVGF_ATROUS_PAIR_ROWS")) c->atrous_pair_rows = atoi(v) == 1 ? 1 : 2;
    if (const char *v = getenv("SVGF_ATROUS_SHAPES")) {      // "a,b,c,...": shape of level 1, 2, 3, ...
        int level = 1;
        for (const char *q = v; *q && level <= SVGF_MAX_LEVELS; level++) {
            c->atrous_shape_level[level] = atoi(q);
            while (*q && *q != ',') q++;
            if (*q == ',') q++;
        }
    }
    memset(c->view_matrix_prev, 0, sizeof(c->view_matrix_prev));
    c->view_matrix_prev[0] = c->view_matrix_prev[5] = c->view_matrix_prev[10] = c->view_matrix_prev[15] = 1.0f;   // glm::mat4()
    int rc = SVGF_OK;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { g_create_err = std::string("svgf_create: ") + cudaGetErrorString(e); delete c; return SVGF_ERR_CUDA; }
    // CUDA loads a kernel's code at its first launch, which may need the device to go idle. A sharded frame launches kernels
    // WHILE kernels of other ranks spin on flags that only later launches raise (the frame's first signal kernel, say), so a
    // first launch in that situation deadlocks until the spin gives up. Load everything now, once per process and device.
    {
        static bool loaded[64] = {false};
        if (device < 64 && !loaded[device]) { preload_pathtrace_kernels(); preload_denoise_kernels(); preload_atrous_kernels(); loaded[device] = true; }
    }
    rc = upload_scene(c, scene);
    if (rc == SVGF_OK) rc = alloc_frame_buffers(c);
    if (rc == SVGF_OK) rc = svgf_reset(c);
    if (rc != SVGF_OK) { g_create_err = c->err; svgf_destroy(c); return rc; }
    *out = c;
    return SVGF_OK;
}

int svgf_destroy(svgf_ctx *c) {
    if (!c) return SVGF_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    DeviceScene &s = c->scene;
    cudaFree(s.geoms); cudaFree(s.materials); cudaFree(s.bvh); cudaFree(s.tri_hot); cudaFree(s.tri_cold); cudaFree(s.textures);
    for (unsigned char *p : s.tex_pixels) cudaFree(p);      // the reference leaks these (pathtrace.cu:136 vs 160-183)
    for (int i = 0; i < SVGF_NCV; i++) { cudaFree(c->cv[i]); cudaFree(c->lv[i]); }
    free(c->tmaps); free(c->tmaps_slide);
    cudaFree(c->done_count);
    if (c->comm_err) cudaFreeHost(c->comm_err);
    for (int i = 0; i < 2; i++) { cudaFree(c->nrm[i]); cudaFree(c->mom[i]); cudaFree(c->hlen[i]); }
    cudaFree(c->pos); cudaFree(c->denoised); cudaFree(c->var_out);
    for (int g = 0; g < 2; g++) { cudaFree(c->alb_set[g]); cudaFree(c->gnp_set[g]); cudaFree(c->gzl_set[g]); cudaFree(c->image_set[g]); }
    if (c->rt_stream) { cudaStreamSynchronize(c->rt_stream); cudaStreamDestroy(c->rt_stream); cudaEventDestroy(c->ev_rt_done); cudaEventDestroy(c->ev_temporal_done); }
    cudaFree(c->stale_nm); cudaFree(c->stale_uv); cudaFree(c->pbo_own); cudaFree(c->kl); cudaFree(c->flags); cudaFree(c->wf_mem); cudaFree(c->rt_counter);
    for (void *p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    cudaFree(c->aos_in); cudaFree(c->aos_out); cudaFree(c->aos_g);
    if (c->pinned_image) cudaFreeHost(c->pinned_image);
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); cudaEventDestroy(c->frame_done);
        cudaEventDestroy(c->copy_done[0]); cudaEventDestroy(c->copy_done[1]); cudaFree(c->denoised_alt);
    }
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->legacy_fence) cudaEventDestroy(c->legacy_fence);
    cudaFree(c->slot_of_input); cudaFree(c->bvh_parent); cudaFree(c->refit_stage);
    cudaFree(c->stage_ctr);
    for (auto &pf : c->prof_pool) for (int i = 0; i < 12; i++) cudaEventDestroy(pf.ev[i]);
    for (auto &r : c->registered_hosts) cudaHostUnregister(r.first);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return SVGF_OK;
}

// pathtraceInit (pathtrace.cu:108-128) + denoiseInit (denoise.cu:41-60): zero what the reference zeroes. Buffers it
// leaves uninitialised (colour history, previous G-buffer, ping-pong) are zeroed too; they are never read before
// being written because history_length == 0 gates every history read (denoise.cu:198).
int svgf_reset(svgf_ctx *c) {
    if (!c) return SVGF_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    const size_t px = c->px;
    cudaStream_t st = c->stream;
    for (int i = 0; i < SVGF_NCV; i++) {
        CK(cudaMemsetAsync(c->cv[i], 0, px * sizeof(float4), st)); CK(cudaMemsetAsync(c->lv[i], 0, px * sizeof(float2), st));
    }
    CK(cudaMemsetAsync(c->done_count, 0, SVGF_NUM_STAGES * sizeof(unsigned), st));
    for (int i = 0; i < 2; i++) {
        CK(cudaMemsetAsync(c->nrm[i], 0, px * sizeof(float4), st));
        CK(cudaMemsetAsync(c->mom[i], 0, px * sizeof(float2), st));
        CK(cudaMemsetAsync(c->hlen[i], 0, px * sizeof(int), st));
    }
    if (c->rt_stream) CK(cudaStreamSynchronize(c->rt_stream));
    c->temporal_done_valid = false;
    CK(cudaMemsetAsync(c->pos, 0, px * sizeof(float4), st));
    for (int g = 0; g < 2; g++) {
        CK(cudaMemsetAsync(c->alb_set[g], 0, px * sizeof(float4), st));
        CK(cudaMemsetAsync(c->gnp_set[g], 0, px * sizeof(float4), st)); CK(cudaMemsetAsync(c->gzl_set[g], 0, px * sizeof(float2), st));
        CK(cudaMemsetAsync(c->image_set[g], 0, px * 12, st));
    }
    c->gset = 0; c->alb = c->alb_set[0]; c->gnp = c->gnp_set[0]; c->gzl = c->gzl_set[0]; c->image = c->image_set[0];
    CK(cudaMemsetAsync(c->denoised, 0, px * 12, st));
    CK(cudaMemsetAsync(c->var_out, 0, px * 4, st));
    if (c->copy_stream) {       // images still in flight belong to the history being discarded
        CK(cudaStreamSynchronize(c->copy_stream));
        c->copy_pending[0] = c->copy_pending[1] = 0;
        CK(cudaMemsetAsync(c->denoised_alt, 0, px * 12, st));
    }
    CK(cudaMemsetAsync(c->stale_nm, 0, px * sizeof(float4), st)); CK(cudaMemsetAsync(c->stale_uv, 0, px * sizeof(float2), st));
    CK(cudaMemsetAsync(c->pbo_own, 0, px * 8, st));
    c->hist_cv = 0; c->cur_nrm = 0; c->cur_mom = 0; c->cur_hlen = 0;
    CK(cudaStreamSynchronize(st));
    if (c->comm_err) *c->comm_err = 0u;      // a reset is the way out of a communication error
    return SVGF_OK;
}

int svgf_set_shard(svgf_ctx *c, const svgf_shard *s) {
    if (!c || !s || s->world < 1 || s->rank < 0 || s->rank >= s->world || s->row_begin < 0 || s->row_end > c->H || s->row_begin > s->row_end) {
        if (c) c->err = "svgf_set_shard: invalid shard";
        return SVGF_ERR_INVALID;
    }
    c->shard = *s;
    return SVGF_OK;
}

// ---- multi-GPU wiring: who owns which rows, and where every rank's planes are mapped ----
static int set_rows(svgf_ctx *c, int rank, int world, const int *row_starts) {
    if (world < 1 || world > SVGF_MAX_RANKS || rank < 0 || rank >= world || !row_starts || row_starts[0] != 0 || row_starts[world] != c->H) {
        c->err = "invalid rank/world/row partition"; return SVGF_ERR_INVALID;
    }
    for (int r = 0; r < world; r++) if (row_starts[r] > row_starts[r + 1]) { c->err = "row partition not monotonic"; return SVGF_ERR_INVALID; }
    c->rows.world = world;
    for (int r = 0; r <= SVGF_MAX_RANKS; r++) c->rows.start[r] = r <= world ? row_starts[r] : c->H;
    c->shard = svgf_shard{rank, world, row_starts[rank], row_starts[rank + 1]};
    return SVGF_OK;
}

int svgf_ipc_handles_size(void) { return SVGF_IPC_NBUF * (int)sizeof(cudaIpcMemHandle_t); }

int svgf_ipc_export(svgf_ctx *c, void *out) {
    if (!c || !out) return SVGF_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    void *b[SVGF_IPC_NBUF]; shared_bufs(c, b);
    cudaIpcMemHandle_t *h = static_cast<cudaIpcMemHandle_t *>(out);
    for (int i = 0; i < SVGF_IPC_NBUF; i++) CK(cudaIpcGetMemHandle(&h[i], b[i]));
    return SVGF_OK;
}

// all_handles: world x SVGF_IPC_NBUF handles in rank order (an all-gather of svgf_ipc_export); row_starts: world + 1 entries.
int svgf_ipc_connect(svgf_ctx *c, int rank, int world, const void *all_handles, const int *row_starts) {
    if (!c || !all_handles) return SVGF_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    int rc = set_rows(c, rank, world, row_starts);
    if (rc) return rc;
    const cudaIpcMemHandle_t *h = static_cast<const cudaIpcMemHandle_t *>(all_handles);
    for (int r = 0; r < world; r++) {
        void *b[SVGF_IPC_NBUF];
        if (r == rank) shared_bufs(c, b);
        else for (int i = 0; i < SVGF_IPC_NBUF; i++) {
            cudaError_t e = cudaIpcOpenMemHandle(&b[i], h[r * SVGF_IPC_NBUF + i], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { c->err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e); return SVGF_ERR_COMM; }
            c->ipc_opened.push_back(b[i]);
        }
        set_peer(c, r, b);
    }
    return SVGF_OK;
}

// Ranks living in ONE process (tests on a single GPU, or a single-process multi-GPU host): wire them directly.
int svgf_peer_connect_local(svgf_ctx **ctxs, int world, const int *row_starts) {
    if (!ctxs || world < 1 || world > SVGF_MAX_RANKS) return SVGF_ERR_INVALID;
    for (int r = 0; r < world; r++) {
        svgf_ctx *c = ctxs[r];
        if (!c || c->W != ctxs[0]->W || c->H != ctxs[0]->H) return SVGF_ERR_INVALID;
        int rc = set_rows(c, r, world, row_starts);
        if (rc) return rc;
        for (int q = 0; q < world; q++) {
            // Two ranks on ONE device: the persistent blocks of a rank's stage kernel fill the GPU and poll the other rank's flags,
            // whose kernel could then never be scheduled. Such ranks (tests) run the stage level by level.
            if (q != r && ctxs[q]->device == c->device) c->atrous_fused = 0;
            if (ctxs[q]->device != c->device) {
                cudaSetDevice(c->device);
                cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { c->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); return SVGF_ERR_COMM; }
                (void)cudaGetLastError();
            }
            void *b[SVGF_IPC_NBUF]; shared_bufs(ctxs[q], b);
            set_peer(c, q, b);
        }
    }
    return SVGF_OK;
}

// 1 if a cross-rank wait timed out since the last svgf_reset (a peer stopped making progress).
int svgf_peer_error(svgf_ctx *c) {
    if (!c) return SVGF_ERR_INVALID;
    return c->comm_err ? (int)*static_cast<volatile unsigned *>(c->comm_err) : 0;
}

enum { SVGF_PROF_MAX_FRAMES = 512 };

int svgf_set_profiling(svgf_ctx *c, int enabled) {
    if (!c) return SVGF_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    c->profiling = enabled != 0;
    c->prof_count = 0;
    return SVGF_OK;
}

// Average device time per stage over the frames rendered since profiling was enabled (at most 512):
// [0] path trace, [1] temporal, [2..8] a-trous level 1..7, [9] pbo pack, [10] whole frame (incl. D2H when requested).
int svgf_stage_times(svgf_ctx *c, float *ms11) {
    if (!c || !ms11) return SVGF_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    double acc[11] = {0}; int cnt[11] = {0};
    for (int f = 0; f < c->prof_count; f++) {
        svgf_ctx::ProfFrame &pf = c->prof_pool[f];
        float ms;
        auto add = [&](int slot, cudaEvent_t a, cudaEvent_t b) { if (cudaEventElapsedTime(&ms, a, b) == cudaSuccess) { acc[slot] += ms; cnt[slot]++; } };
        add(0, pf.ev[0], pf.ev[1]);
        cudaEvent_t prev = pf.ev[1];
        if (pf.denoise) {
            add(1, pf.ev[1], pf.ev[2]); prev = pf.ev[2];
            for (int l = 1; l <= pf.nlevel; l++) { add(1 + l, prev, pf.ev[2 + l]); prev = pf.ev[2 + l]; }
        }
        add(9, prev, pf.ev[10]);
        add(10, pf.ev[0], pf.ev[11]);
    }
    (void)cudaGetLastError();
    for (int i = 0; i < 11; i++) c->stage_ms[i] = cnt[i] ? (float)(acc[i] / cnt[i]) : 0.f;
    memcpy(ms11, c->stage_ms, sizeof(c->stage_ms));
    return SVGF_OK;
}

void *svgf_stream(svgf_ctx *c) { return c ? (void *)c->stream : nullptr; }

int svgf_set_option(svgf_ctx *c, const char *name, int value) {
    if (!c || !name) return SVGF_ERR_INVALID;
    if (!strcmp(name, "reprojection_fov_aspect")) { c->opt_reprojection_fov_aspect = value != 0; return SVGF_OK; }
    if (!strcmp(name, "history_cap")) { if (value < 0) return SVGF_ERR_INVALID; c->opt_history_cap = value; return SVGF_OK; }
    if (!strcmp(name, "light_sampling_all")) { c->opt_light_sampling_all = value != 0; return SVGF_OK; }
    if (!strcmp(name, "cuda_graph")) { c->opt_cuda_graph = value != 0; return SVGF_OK; }
    if (!strcmp(name, "frame_overlap")) { c->frame_overlap = value != 0; return SVGF_OK; }
    if (!strcmp(name, "spatial_variance_estimate")) { c->opt_spatial_variance = value != 0; return SVGF_OK; }
    c->err = std::string("svgf_set_option: unknown option '") + name + "'";
    return SVGF_ERR_UNKNOWN_NAME;
}

// A wait that gave up let its frame continue on stale rows of a peer: everything rendered since is suspect.
static int comm_check(svgf_ctx *c) {
    if (c->rows.world > 1 && c->comm_err && *static_cast<volatile unsigned *>(c->comm_err)) {
        c->err = "a cross-rank wait timed out (a peer stopped making progress); frames since then are invalid -- svgf_reset on all ranks to recover";
        return SVGF_ERR_COMM;
    }
    return SVGF_OK;
}

int svgf_sync(svgf_ctx *c) {
    if (!c) return SVGF_ERR_INVALID;
    CK(cudaStreamSynchronize(c->stream));
    if (c->copy_stream) CK(cudaStreamSynchronize(c->copy_stream));
    return comm_check(c);
}

// Page-lock a caller buffer for svgf_render_async (and for direct DMA from svgf_render). The buffer must stay allocated
// until svgf_unregister_host or svgf_destroy.
int svgf_register_host(svgf_ctx *c, void *host, size_t bytes) {
    if (!c || !host || !bytes) return SVGF_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    for (auto &r : c->registered_hosts) if (r.first == host) return r.second >= bytes ? SVGF_OK : SVGF_ERR_INVALID;
    CK(cudaHostRegister(host, bytes, cudaHostRegisterDefault));
    c->registered_hosts.push_back({host, bytes});
    return SVGF_OK;
}

int svgf_unregister_host(svgf_ctx *c, void *host) {
    if (!c || !host) return SVGF_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    for (size_t i = 0; i < c->registered_hosts.size(); i++)
        if (c->registered_hosts[i].first == host) {
            CK(cudaStreamSynchronize(c->stream));
            if (c->copy_stream) CK(cudaStreamSynchronize(c->copy_stream));
            for (int k = 0; k < 2; k++) if (c->copy_host[k] == host) { c->copy_pending[k] = 0; c->copy_host[k] = nullptr; }
            CK(cudaHostUnregister(host));
            c->registered_hosts.erase(c->registered_hosts.begin() + i);
            return SVGF_OK;
        }
    c->err = "svgf_unregister_host: not registered by this context";
    return SVGF_ERR_INVALID;
}

}  // extern "C"

// ---- the denoise driver on SoA planes (denoise.cu:349-402) -------------------------------------------
// Inputs: c->image (1-spp colour), c->nrm[cur_nrm], c->pos, c->alb. Output: c->denoised, c->var_out.
template <class T> static HaloPlane halo_plane(const T *local, const PeerPtr<T> &peers) {
    HaloPlane h; h.local = local; h.esz = (int)sizeof(T);
    for (int r = 0; r < SVGF_MAX_RANKS; r++) h.peer[r] = peers.p[r];
    return h;
}

static int denoise_soa(svgf_ctx *c, const float *image, const svgf_camera *cam, const svgf_params *P, cudaEvent_t *ev, bool gbuf_pushed) {
    // Four colour buffers rotate: H = the history the previous frame left (read in place by every rank's temporal pass and
    // therefore never written during this frame), the accumulated colour, and two for the a-trous ping-pong.
    const int H_old = c->hist_cv;
    const int acc_slot = (H_old + 1) % SVGF_NCV;
    float4 *acc = c->cv[acc_slot];
    // Sharded frames, push mode: a level taps rows up to 2*step beyond the strip. Instead of reading them from their owners
    // in place (32-byte gathers over NVLink, every coarse tile), each stage's producer stores the rows its neighbours will
    // tap into THEIR copy of the plane as it produces them (dual stores), and the levels read local memory only. The
    // G-buffer view travels once per frame, for the coarsest level's reach.
    const bool filter = P->right_view_option == 0 && P->atrous_nlevel > 0 && P->spatial_enable;
    const bool sharded = c->rows.world > 1 && filter;
    const bool push = c->halo_push && sharded;
    if (push && !gbuf_pushed) {
        const HaloPlane g[2] = {halo_plane(c->gnp, c->p_gnp), halo_plane(c->gzl, c->p_gzl)};
        CK(launch_halo_push(c, 2 << P->atrous_nlevel, g, 2));
    }
    const float color_alpha = P->temporal_enable ? P->color_alpha : 1.0f;
    const float moment_alpha = P->temporal_enable ? P->moment_alpha : 1.0f;
    // generateRayFromCamera spreads the frame over +-pixelLength * res / 2 = +-tan(fovy) * aspect and +-tan(fovy) at unit depth
    // (scene.cpp:159-166, pathtrace.cu:197-200); the reference's back-projection assumes both are 1 (denoise.cu:201-207)
    float clip_rx = 1.0f, clip_ry = 1.0f;
    if (c->opt_reprojection_fov_aspect) {
        clip_rx = 1.0f / (cam->pixelLength[0] * (float)c->W * 0.5f);
        clip_ry = 1.0f / (cam->pixelLength[1] * (float)c->H * 0.5f);
    }
    if (P->temporal_enable) {
        // level 1 (step 2) taps +-4 rows: those rows of the accumulated planes go to the neighbours from inside the kernel,
        // whose last block raises the stage flag
        HaloOut ho = halo_out(c, SVGF_STAGE_TEMPORAL, sharded ? 4 : 0, true);
        if (!push) memset(&ho.peers.lo, 0, sizeof(ho.peers.lo)), memset(&ho.peers.hi, 0, sizeof(ho.peers.hi));   // pull mode: flags only
        CK(launch_temporal(c, image, c->nrm[c->cur_nrm], c->p_nrm[c->cur_nrm ^ 1], c->pos, c->p_cv[H_old], c->p_mom[c->cur_mom],
                           c->p_hlen[c->cur_hlen], acc, c->lv[acc_slot], c->mom[c->cur_mom ^ 1], c->hlen[c->cur_hlen ^ 1],
                           c->view_matrix_prev, color_alpha, moment_alpha, clip_rx, clip_ry, ho, c->p_cv[acc_slot], c->p_lv[acc_slot]));
        if (c->opt_spatial_variance) {
            // whole-frame contexts only: the estimate reads a 7x7 neighbourhood of moments that other ranks would still be writing
            if (c->rows.world > 1 || c->shard.row_begin != 0 || c->shard.row_end != c->H) { c->err = "spatial_variance_estimate: single-GPU, whole-frame contexts only"; return SVGF_ERR_INVALID; }
            CK(launch_spatial_variance(c, c->hlen[c->cur_hlen ^ 1], c->mom[c->cur_mom ^ 1], c->nrm[c->cur_nrm], acc, c->lv[acc_slot]));
        }
    } else {
        CK(launch_no_temporal(c, image, acc, c->lv[acc_slot]));
        if (push) {
            const HaloPlane h[2] = {halo_plane(c->cv[acc_slot], c->p_cv[acc_slot]), halo_plane(c->lv[acc_slot], c->p_lv[acc_slot])};
            CK(launch_halo_push(c, 4, h, 2));
        }
        if (sharded) CK(launch_signal(c, SVGF_STAGE_TEMPORAL, 4));
    }
    if (ev) CK(cudaEventRecord(ev[2], c->stream));
    // the next frame's path tracer may start from here (cross-frame overlap, frame_body): nothing below reads what it writes
    if (c->ev_temporal_done) { CK(cudaEventRecord(c->ev_temporal_done, c->stream)); c->temporal_done_valid = true; }
    int new_hist = acc_slot;        // denoise.cu:366/370: colour history := accumulated (or input) colour
    if (P->right_view_option == 1) {
        // DebugView shows dev_history_length, i.e. the length BEFORE this frame's update (denoise.cu:374)
        CK(launch_cv_to_outputs(c, acc, c->denoised, c->var_out));
        CK(launch_debug_view(c, 1, c->hlen[c->cur_hlen], acc, c->denoised));
    } else if (P->right_view_option == 2) {
        CK(launch_cv_to_outputs(c, acc, c->denoised, c->var_out));
        CK(launch_debug_view(c, 2, c->hlen[c->cur_hlen], acc, c->denoised));
    } else if (P->atrous_nlevel == 0 || !P->spatial_enable) {
        CK(launch_cv_to_outputs(c, acc, c->denoised, c->var_out));
    } else {
        int src = acc_slot;
        AtrousArgs largs[SVGF_MAX_LEVELS];
        bool post_push[SVGF_MAX_LEVELS] = {false}; HaloOut post_ho[SVGF_MAX_LEVELS];
        for (int level = 1; level <= P->atrous_nlevel; level++) {
            const bool last = level == P->atrous_nlevel;
            const bool is_hist = level == P->history_level;
            // destination: a slot that is neither the source, nor the (new) history, nor the history other ranks still read
            int dst = -1;
            for (int s = 0; s < SVGF_NCV; s++) if (s != src && s != new_hist && s != H_old) { dst = s; break; }
            AtrousArgs a;
            a.src_slot = src;
            a.cv_in = c->cv[src];
            a.cv_out = (!last || is_hist) ? c->cv[dst] : nullptr;
            a.lv_in = c->lv[src]; a.lv_out = c->lv[dst]; a.dst_slot = dst;
            a.nrm = c->nrm[c->cur_nrm]; a.pos = c->pos; a.alb = c->alb; a.gnp = c->gnp; a.gzl = c->gzl; a.gset = c->gset;
            a.denoised_out = last ? c->denoised : nullptr; a.var_out = last ? c->var_out : nullptr;
            a.level = level; a.is_last = last; a.blur_variance = P->blurvariance; a.addcolor = (P->sepcolor && P->addcolor);
            a.sigma_c = P->sigmal; a.sigma_n = P->sigman; a.sigma_x = P->sigmax;
            // A level reads rows of the neighbours in reach (2 * step): they must have produced them -- and be done reading what
            // this level's stores overwrite in their planes, which the same flag says. The wait sits in the edge blocks of the
            // level's pre-pass; the edge rows and the flag for the NEXT level (reach 4 * step) leave from the tile kernel.
            const int prev_stage = level == 1 ? SVGF_STAGE_TEMPORAL : SVGF_STAGE_LEVEL0 + level - 1;
            a.wait = halo_in(c, prev_stage, sharded ? (2 << level) : 0, c->seq);
            a.ho = halo_out(c, SVGF_STAGE_LEVEL0 + level, (sharded && !last) ? (4 << level) : 0, true);
            if (!push) memset(&a.ho.peers.lo, 0, sizeof(a.ho.peers.lo)), memset(&a.ho.peers.hi, 0, sizeof(a.ho.peers.hi));
            // From level 3 on the rows a neighbour taps next (4 * step either side: a quarter, then half of a 270-row strip) leave
            // through the dense copy kernel instead of the tile kernel's dual stores: those are 16-byte stores `step` pixels
            // apart, which crawl over NVLink (8 x B200, 4K: level 4 took 137 us against 61 us for the strip's own work).
            post_push[level - 1] = push && !last && (4 << level) >= c->halo_copy_from_rows;
            if (post_push[level - 1]) { post_ho[level - 1] = a.ho; memset(&a.ho, 0, sizeof(a.ho)); }
            largs[level - 1] = a;
            if (is_hist) new_hist = dst;     // denoise.cu:391: colour history := this level's output
            src = dst;
        }
        // One launch for the whole stage where possible (atrous.cu: atrous_stage_kernel), else level by level. (Per-level event
        // records: with the single launch they all follow it, so the first level's interval is the stage's.)
        if (atrous_stage_possible(c, largs, P->atrous_nlevel)) {
            for (int l = 0; l < P->atrous_nlevel; l++) if (post_push[l]) largs[l].ho = post_ho[l];        // the stage kernel pushes and signals by itself
            CK(launch_atrous_stage(c, largs, P->atrous_nlevel));
            if (ev) for (int level = 1; level <= P->atrous_nlevel; level++) CK(cudaEventRecord(ev[2 + level], c->stream));
        } else {
            for (int level = 1; level <= P->atrous_nlevel; level++) {
                const AtrousArgs &a = largs[level - 1];
                CK(launch_atrous(c, a));
                if (post_push[level - 1]) {
                    const HaloPlane h[2] = {halo_plane(c->cv[a.dst_slot], c->p_cv[a.dst_slot]), halo_plane(c->lv[a.dst_slot], c->p_lv[a.dst_slot])};
                    CK(launch_halo_push(c, 4 << level, h, 2, &post_ho[level - 1]));
                }
                if (ev) CK(cudaEventRecord(ev[2 + level], c->stream));
            }
        }
    }
    // denoise.cu:396-399 by rotation
    c->hist_cv = new_hist;
    c->cur_nrm ^= 1;            // this frame's normals/geomIds become "previous"
    if (P->temporal_enable) { c->cur_mom ^= 1; c->cur_hlen ^= 1; }
    svgf_view_matrix(cam, c->view_matrix_prev);
    return SVGF_OK;
}
// NOTE on the temporal-off path: the reference still copies moment_acc/history_length_update (never written in that
// frame, denoise.cu:397-398) over the histories, i.e. leaves them undefined; here they simply stay as they were.

// ev[0] start, ev[1] after rt, ev[2] after temporal, ev[3..9] after a-trous level 1..7, ev[10] after pack, ev[11] end
static cudaEvent_t *prof_begin(svgf_ctx *c, const svgf_params *P) {
    if (!c->profiling || c->prof_count >= SVGF_PROF_MAX_FRAMES) return nullptr;
    if ((int)c->prof_pool.size() <= c->prof_count) {
        svgf_ctx::ProfFrame pf;
        for (int i = 0; i < 12; i++) if (cudaEventCreate(&pf.ev[i]) != cudaSuccess) return nullptr;
        c->prof_pool.push_back(pf);
    }
    svgf_ctx::ProfFrame &pf = c->prof_pool[c->prof_count++];
    const bool filt = P->denoise_enable && P->right_view_option == 0 && P->spatial_enable && P->atrous_nlevel > 0;
    pf.denoise = P->denoise_enable ? 1 : 0;
    pf.nlevel = filt ? P->atrous_nlevel : 0;
    return pf.ev;
}

// Is `p` page-locked right now (by the caller, or through svgf_register_host)? Asked of the driver on every call: a cached
// answer would outlive the memory it was given for.
static bool host_is_pinned(const void *p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost) return true;
    (void)cudaGetLastError();
    return false;
}

// Pipelined readback: the frame's image leaves through a second stream while the next frame renders, so the final colour
// needs two buffers (the copy of frame N reads one while frame N+1's last level writes the other).
static int async_setup(svgf_ctx *c) {
    if (c->copy_stream) return SVGF_OK;
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->frame_done, cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) CK(cudaEventCreateWithFlags(&c->copy_done[i], cudaEventDisableTiming));
    CK(dalloc(&c->denoised_alt, 3 * c->px));
    CK(cudaMemsetAsync(c->denoised_alt, 0, c->px * 12, c->stream));
    return SVGF_OK;
}

// Work that writes `denoised` outside the async rotation must not overtake an image copy still reading it.
static int order_after_copies(svgf_ctx *c) {
    for (int i = 0; i < 2; i++) if (c->copy_pending[i]) CK(cudaStreamWaitEvent(c->stream, c->copy_done[i], 0));
    return SVGF_OK;
}

// The launches of one frame, from the frame-boundary wait to the PBO pack (everything that can live in a CUDA graph).
static int frame_body(svgf_ctx *c, const svgf_camera *cam, const svgf_params *P, int frame, void *pbo_dev, cudaEvent_t *ev) {
    c->seq++;
    // before this frame overwrites planes that peers read in place (any rank: the reprojection may land in any strip), they
    // must have finished the previous frame
    if (c->seq > 1) CK(launch_wait(c, SVGF_STAGE_FRAME, c->seq - 1, -1));
    if (ev) CK(cudaEventRecord(ev[0], c->stream));      // after the wait: stage times measure this rank's own work
    RtParams rp;
    rp.W = c->W; rp.H = c->H; rp.row_begin = c->shard.row_begin; rp.row_end = c->shard.row_end;
    rp.frame = frame; rp.max_depth = P->tracedepth; rp.trace_shadowray = P->shadowray; rp.reduce_var = P->reducevar;
    rp.denoise = P->denoise_enable; rp.sepcolor = P->sepcolor; rp.sintensity = P->sintensity; rp.lightradius = P->lightradius;
    atrous_scales(P->sigman, P->sigmax, &rp.kn, &rp.kx);
    rp.n_lights = (c->opt_light_sampling_all && c->scene.geoms && c->n_lights > 1) ? c->n_lights : 0;
    for (int i = 0; i < 8; i++) rp.lights[i] = c->lights[i];
    rp.cam = *cam;
    c->gbuf_nrm = c->cur_nrm;            // where this frame's normals/geomIds live (svgf_fetch("gbuffer"))
    const bool filter = P->denoise_enable && P->right_view_option == 0 && P->atrous_nlevel > 0 && P->spatial_enable;
    bool gbuf_pushed = false;
    // The G-buffer rows the neighbours tap leave through a copy kernel after the path tracer. Dual stores from inside rt_kernel
    // (SVGF_RT_PUSH=1, A/B) save that launch but cost the kernel 15 % (1.80 vs 1.57 ms for half a 4K frame on 2 x B200): the
    // extra live state does not fit its 64 registers.
    static const bool rt_push = getenv("SVGF_RT_PUSH") && atoi(getenv("SVGF_RT_PUSH")) != 0;
    c->gbuf_nan_possible = c->scene_nan_possible;      // this frame's G-buffer comes from our own path tracer
    // Cross-frame overlap: this frame's path tracer goes to its own stream and starts as soon as the PREVIOUS frame's temporal
    // pass -- the last reader of the buffers it writes that exist only once (normals of two frames ago, positions) -- has finished,
    // next to that frame's a-trous stage (which is short of issue slots at every launch boundary and tail, while the path tracer
    // never runs out of blocks). Frames of a single GPU in the denoising configuration only; not under the stage timers (whose
    // intervals assume one stream) and not inside a captured graph.
    const bool overlap = c->frame_overlap && c->rows.world == 1 && c->shard.row_begin == 0 && c->shard.row_end == c->H && !ev && !c->opt_cuda_graph &&
                         P->denoise_enable && c->atrous_variant == 2 && c->rt_variant == 0;
    if (overlap) {
        if (!c->rt_stream) {
            CK(cudaStreamCreateWithFlags(&c->rt_stream, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&c->ev_rt_done, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->ev_temporal_done, cudaEventDisableTiming));
        }
        c->gset ^= 1;
        c->alb = c->alb_set[c->gset]; c->gnp = c->gnp_set[c->gset]; c->gzl = c->gzl_set[c->gset]; c->image = c->image_set[c->gset];
        // after the previous frame's temporal pass (if that frame ran without overlap, after all of it)
        if (!c->temporal_done_valid) CK(cudaEventRecord(c->ev_temporal_done, c->stream));
        CK(cudaStreamWaitEvent(c->rt_stream, c->ev_temporal_done, 0));
        c->rt_launch_stream = c->rt_stream;
    } else if (c->rt_stream) {
        CK(cudaStreamWaitEvent(c->stream, c->ev_rt_done, 0));      // (defensive: every overlapped path tracer was already waited for)
    }
    c->temporal_done_valid = false;
    {
        const cudaError_t e = launch_pathtrace(c, rp, c->nrm[c->cur_nrm], (rt_push && filter && c->halo_push && c->rows.world > 1) ? (2 << P->atrous_nlevel) : 0, &gbuf_pushed);
        c->rt_launch_stream = nullptr;
        CK(e);
    }
    if (overlap) {
        CK(cudaEventRecord(c->ev_rt_done, c->rt_stream));
        CK(cudaStreamWaitEvent(c->stream, c->ev_rt_done, 0));
    }
    if (ev) CK(cudaEventRecord(ev[1], c->stream));
    if (P->denoise_enable) {
        int rc = denoise_soa(c, c->image, cam, P, ev, gbuf_pushed);
        if (rc != SVGF_OK) return rc;
    } else {
        CK(launch_copy_f3(c, c->denoised, c->image));        // pathtrace.cu:440
    }
    CK(launch_signal(c, SVGF_STAGE_FRAME, -1));
    unsigned char *pbo = pbo_dev ? static_cast<unsigned char *>(pbo_dev) : c->pbo_own;
    CK(launch_pack_pbo(c, pbo, c->image, c->denoised));
    if (ev) CK(cudaEventRecord(ev[10], c->stream));
    return SVGF_OK;
}

static int render_impl(svgf_ctx *c, const svgf_camera *cam, const svgf_params *P, int frame, void *pbo_dev, float *host_image, bool async) {
    if (!c || !cam || !P) return SVGF_ERR_INVALID;
    if (cam->resolution[0] != c->W || cam->resolution[1] != c->H) { c->err = "svgf_render: camera resolution differs from the context's"; return SVGF_ERR_INVALID; }
    if (P->atrous_nlevel < 0 || P->atrous_nlevel > SVGF_MAX_LEVELS || P->tracedepth < 0) { c->err = "svgf_render: parameter out of range"; return SVGF_ERR_INVALID; }
    CK(cudaSetDevice(c->device));
    if (async && host_image) {
        int rc = async_setup(c);
        if (rc != SVGF_OK) return rc;
        // the copy outlives this call, so the buffer must be page-locked; done here on first sight (documented in the header:
        // the caller keeps it allocated until svgf_unregister_host / svgf_destroy)
        if (!host_is_pinned(host_image) && svgf_register_host(c, host_image, c->px * 12) != SVGF_OK) {
            c->err = "svgf_render_async: host_image cannot be page-locked"; return SVGF_ERR_INVALID;
        }
        // this frame writes the buffer whose copy was queued two frames ago: that copy must have drained
        std::swap(c->denoised, c->denoised_alt);
        c->copy_slot ^= 1;
        if (c->copy_pending[c->copy_slot]) CK(cudaStreamWaitEvent(c->stream, c->copy_done[c->copy_slot], 0));
    } else {
        int rc = order_after_copies(c);
        if (rc != SVGF_OK) return rc;
    }
    { int rc = comm_check(c); if (rc != SVGF_OK) return rc; }
    cudaEvent_t *ev = prof_begin(c, P);
    // the frame as a CUDA graph ("cuda_graph" option). The frame's launches are stream-captured every
    // frame -- the host side of a capture is the cheap part, and the kernel arguments change every frame (frame number,
    // camera, the rotating buffer roles, the sequence number of a sharded frame) -- and the captured graph UPDATES the
    // instantiated one in place (same topology: cudaGraphExecUpdate only patches node parameters), which is then launched
    // as one unit: the device runs the 13 kernels of a frame back to back without per-launch front-end gaps. Frames whose
    // kernel sequence differs (other switches) re-instantiate. Not while profiling (event records between the stages) and
    // not for the first two frames of a context (one-time function attributes are set outside any capture).
    const bool graph = c->opt_cuda_graph && !ev && c->frames_rendered >= 2;
    if (graph) CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
    int body_rc = frame_body(c, cam, P, frame, pbo_dev, ev);
    if (graph) {
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamEndCapture(c->stream, &g);
        if (body_rc == SVGF_OK && e != cudaSuccess) { c->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e); body_rc = SVGF_ERR_CUDA; }
        if (body_rc == SVGF_OK && c->graph_exec) {
            cudaGraphExecUpdateResultInfo info;
            if (cudaGraphExecUpdate(c->graph_exec, g, &info) != cudaSuccess) {      // another kernel sequence: start over
                (void)cudaGetLastError();
                cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr;
            }
        }
        if (body_rc == SVGF_OK && !c->graph_exec) {
            e = cudaGraphInstantiate(&c->graph_exec, g, 0);
            if (e != cudaSuccess) { c->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e); body_rc = SVGF_ERR_CUDA; }
        }
        if (g) cudaGraphDestroy(g);
        if (body_rc == SVGF_OK) {
            e = cudaGraphLaunch(c->graph_exec, c->stream);
            if (e != cudaSuccess) { c->err = std::string("cudaGraphLaunch: ") + cudaGetErrorString(e); body_rc = SVGF_ERR_CUDA; }
        }
    }
    if (body_rc != SVGF_OK) return body_rc;
    c->frames_rendered++;
    const size_t off = (size_t)c->shard.row_begin * c->W * 3, n = (size_t)(c->shard.row_end - c->shard.row_begin) * c->W * 3;
    if (host_image && async) {
        if (ev) CK(cudaEventRecord(ev[11], c->stream));
        CK(cudaEventRecord(c->frame_done, c->stream));
        CK(cudaStreamWaitEvent(c->copy_stream, c->frame_done, 0));
        CK(cudaMemcpyAsync(host_image + off, c->denoised + off, n * sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream));
        CK(cudaEventRecord(c->copy_done[c->copy_slot], c->copy_stream));
        c->copy_pending[c->copy_slot] = 1; c->copy_host[c->copy_slot] = host_image;
    } else if (host_image) {   // pathtrace.cu:450 (scene->state.image): synchronous like the reference
        // page-locked memory (the caller's own, or svgf_register_host) takes the DMA directly; anything else goes through the
        // context's staging buffer -- caller memory is never registered behind the caller's back here
        if (host_is_pinned(host_image)) {
            CK(cudaMemcpyAsync(host_image + off, c->denoised + off, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
            if (ev) CK(cudaEventRecord(ev[11], c->stream));
            CK(cudaStreamSynchronize(c->stream));
        } else {
            CK(cudaMemcpyAsync(c->pinned_image + off, c->denoised + off, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
            if (ev) CK(cudaEventRecord(ev[11], c->stream));
            CK(cudaStreamSynchronize(c->stream));
            memcpy(host_image + off, c->pinned_image + off, n * sizeof(float));
        }
    } else if (ev) {
        CK(cudaEventRecord(ev[11], c->stream));
    }
    return SVGF_OK;
}

extern "C" int svgf_render(svgf_ctx *c, const svgf_camera *cam, const svgf_params *P, int frame, void *pbo_dev, float *host_image) {
    return render_impl(c, cam, P, frame, pbo_dev, host_image, false);
}

extern "C" int svgf_render_async(svgf_ctx *c, const svgf_camera *cam, const svgf_params *P, int frame, void *pbo_dev, float *host_image) {
    return render_impl(c, cam, P, frame, pbo_dev, host_image, true);
}

// Blocks until the image most recently queued into `host_image` by svgf_render_async has arrived (NULL: all of them).
extern "C" int svgf_wait_image(svgf_ctx *c, const float *host_image) {
    if (!c) return SVGF_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    for (int i = 0; i < 2; i++)
        if (c->copy_pending[i] && (!host_image || c->copy_host[i] == host_image)) {
            CK(cudaEventSynchronize(c->copy_done[i]));
            c->copy_pending[i] = 0;
        }
    return comm_check(c);
}

extern "C" int svgf_denoise(svgf_ctx *c, float *output_dev, const float *input_dev, const svgf_gbuffer_texel *gbuffer_dev,
                            const svgf_camera *cam, const svgf_params *P) {
    if (!c || !output_dev || !input_dev || !gbuffer_dev || !cam || !P) return SVGF_ERR_INVALID;
    if (cam->resolution[0] != c->W || cam->resolution[1] != c->H) { c->err = "svgf_denoise: camera resolution differs from the context's"; return SVGF_ERR_INVALID; }
    if (P->atrous_nlevel < 0 || P->atrous_nlevel > SVGF_MAX_LEVELS) { c->err = "svgf_denoise: atrous_nlevel out of range"; return SVGF_ERR_INVALID; }
    if (c->rows.world > 1 || c->shard.row_begin != 0 || c->shard.row_end != c->H) { c->err = "svgf_denoise: the AoS entry point is single-GPU; sharded frames go through svgf_render"; return SVGF_ERR_INVALID; }
    CK(cudaSetDevice(c->device));
    { int rc = order_after_copies(c); if (rc != SVGF_OK) return rc; }
    // The reference's denoise() launches into the legacy default stream, so whatever the caller queued there before the call --
    // typically the cudaMemcpy that fills `input_dev` / `gbuffer_dev`, which for pageable host memory RETURNS BEFORE its DMA has
    // landed -- is ordered before its kernels. This library's stream is non-blocking: take the same ordering explicitly.
    // (Found by the drop-in test: the shim's denoise() behind the reference's harness read half-copied inputs.)
    if (!c->legacy_fence) CK(cudaEventCreateWithFlags(&c->legacy_fence, cudaEventDisableTiming));
    CK(cudaEventRecord(c->legacy_fence, cudaStreamLegacy));
    CK(cudaStreamWaitEvent(c->stream, c->legacy_fence, 0));
    cudaEvent_t *ev = prof_begin(c, P);
    if (ev) { CK(cudaEventRecord(ev[0], c->stream)); CK(cudaEventRecord(ev[1], c->stream)); }
    float kn, kx;
    atrous_scales(P->sigman, P->sigmax, &kn, &kx);
    c->gbuf_nrm = c->cur_nrm;
    c->gbuf_nan_possible = true;            // a caller's G-buffer: anything may be in it
    CK(launch_aos_to_soa(c, gbuffer_dev, c->nrm[c->cur_nrm], c->pos, c->alb, kn, kx));
    int rc = denoise_soa(c, input_dev, cam, P, ev, false);
    if (rc != SVGF_OK) return rc;
    CK(cudaMemcpyAsync(output_dev, c->denoised, c->px * 12, cudaMemcpyDeviceToDevice, c->stream));
    if (ev) { CK(cudaEventRecord(ev[10], c->stream)); CK(cudaEventRecord(ev[11], c->stream)); }
    CK(cudaStreamSynchronize(c->stream));       // denoise.cu:401
    return SVGF_OK;
}

static int ensure_aos(svgf_ctx *c) {
    if (!c->aos_in) { CK(dalloc(&c->aos_in, 3 * c->px)); CK(dalloc(&c->aos_out, 3 * c->px)); CK(dalloc(&c->aos_g, c->px)); }
    return SVGF_OK;
}

extern "C" int svgf_denoise_host(svgf_ctx *c, float *output, const float *input, const svgf_gbuffer_texel *gbuffer,
                                 const svgf_camera *cam, const svgf_params *P) {
    if (!c || !output || !input || !gbuffer) return SVGF_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    int rc = ensure_aos(c);
    if (rc) return rc;
    CK(cudaMemcpyAsync(c->aos_in, input, c->px * 12, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->aos_g, gbuffer, c->px * sizeof(svgf_gbuffer_texel), cudaMemcpyHostToDevice, c->stream));
    rc = svgf_denoise(c, c->aos_out, c->aos_in, c->aos_g, cam, P);
    if (rc) return rc;
    CK(cudaMemcpy(output, c->aos_out, c->px * 12, cudaMemcpyDeviceToHost));
    return SVGF_OK;
}

extern "C" int svgf_atrous_host(svgf_ctx *c, float *color_out, float *variance_out, const float *color_in,
                                const float *variance_in, const svgf_gbuffer_texel *gbuffer, int level, int is_last,
                                const svgf_params *P) {
    if (!c || !color_out || !variance_out || !color_in || !variance_in || !gbuffer || !P || level < 1 || level > SVGF_MAX_LEVELS) return SVGF_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    int rc = ensure_aos(c);
    if (rc) return rc;
    const size_t px = c->px;
    std::vector<float4> cv(px);
    for (size_t i = 0; i < px; i++) cv[i] = make_float4(color_in[3 * i], color_in[3 * i + 1], color_in[3 * i + 2], variance_in[i]);
    CK(cudaMemcpyAsync(c->cv[0], cv.data(), px * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->aos_g, gbuffer, px * sizeof(svgf_gbuffer_texel), cudaMemcpyHostToDevice, c->stream));
    float kn, kx;
    atrous_scales(P->sigman, P->sigmax, &kn, &kx);
    c->gbuf_nan_possible = true;            // a caller's G-buffer
    CK(launch_aos_to_soa(c, c->aos_g, c->nrm[0], c->pos, c->alb, kn, kx));
    {
        std::vector<float2> lm(px);
        for (size_t i = 0; i < px; i++)
            lm[i] = make_float2((float)(0.2126 * color_in[3 * i] + 0.7152 * color_in[3 * i + 1] + 0.0722 * color_in[3 * i + 2]), variance_in[i]);
        CK(cudaMemcpyAsync(c->lv[0], lm.data(), px * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    AtrousArgs a;
    a.src_slot = 0;
    a.cv_in = c->cv[0]; a.cv_out = c->cv[1]; a.lv_in = c->lv[0]; a.lv_out = c->lv[1]; a.dst_slot = 1; a.nrm = c->nrm[0]; a.pos = c->pos; a.alb = c->alb; a.gnp = c->gnp; a.gzl = c->gzl; a.gset = c->gset;
    a.denoised_out = is_last ? c->denoised : nullptr; a.var_out = is_last ? c->var_out : nullptr;
    a.level = level; a.is_last = is_last != 0; a.blur_variance = P->blurvariance; a.addcolor = (P->sepcolor && P->addcolor);
    a.sigma_c = P->sigmal; a.sigma_n = P->sigman; a.sigma_x = P->sigmax;
    CK(launch_atrous(c, a));
    CK(cudaMemcpyAsync(cv.data(), c->cv[1], px * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < px; i++) { color_out[3 * i] = cv[i].x; color_out[3 * i + 1] = cv[i].y; color_out[3 * i + 2] = cv[i].z; variance_out[i] = cv[i].w; }
    c->hist_cv = 0; c->cur_nrm = 0;     // scratch use of the frame buffers: caller should svgf_reset before rendering again
    return SVGF_OK;
}

namespace {
__global__ void split_cv_kernel(size_t n, const float4 *__restrict__ cv, float *__restrict__ rgb, float *__restrict__ w) {
    const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float4 c = cv[p];
    if (rgb) { rgb[3 * p] = c.x; rgb[3 * p + 1] = c.y; rgb[3 * p + 2] = c.z; }
    if (w) w[p] = c.w;
}
}  // namespace

extern "C" int svgf_fetch(svgf_ctx *c, const char *name, void *host, size_t bytes) {
    if (!c || !name || !host) return SVGF_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    const size_t px = c->px;
    auto d2h = [&](const void *src, size_t need) -> int {
        if (bytes != need) { c->err = std::string("svgf_fetch(") + name + "): size mismatch"; return SVGF_ERR_INVALID; }
        CK(cudaMemcpy(host, src, need, cudaMemcpyDeviceToHost));
        return SVGF_OK;
    };
    if (!strcmp(name, "image")) return d2h(c->image, px * 12);
    if (!strcmp(name, "denoised") || !strcmp(name, "host_image")) return d2h(c->denoised, px * 12);
    if (!strcmp(name, "variance")) return d2h(c->var_out, px * 4);
    if (!strcmp(name, "pbo")) return d2h(c->pbo_own, px * 8);
    // after a frame the rotation has already happened: "accumulated"/"update" buffers are the new histories
    if (!strcmp(name, "moment_acc") || !strcmp(name, "moment_history")) return d2h(c->mom[c->cur_mom], px * 8);
    if (!strcmp(name, "history_length") || !strcmp(name, "history_length_update")) return d2h(c->hlen[c->cur_hlen], px * 4);
    int rc = ensure_aos(c);
    if (rc) return rc;
    if (!strcmp(name, "color_history") || !strcmp(name, "variance_history")) {
        split_cv_kernel<<<(unsigned)((px + 255) / 256), 256, 0, c->stream>>>(px, c->cv[c->hist_cv], c->aos_out, (float *)c->aos_g);
        CK(cudaGetLastError()); CK(cudaStreamSynchronize(c->stream));
        return !strcmp(name, "color_history") ? d2h(c->aos_out, px * 12) : d2h(c->aos_g, px * 4);
    }
    if (!strcmp(name, "gbuffer") || !strcmp(name, "gbuffer_prev")) {
        CK(launch_soa_to_aos(c, c->nrm[c->gbuf_nrm], c->pos, c->alb, c->aos_g));
        CK(cudaStreamSynchronize(c->stream));
        return d2h(c->aos_g, px * sizeof(svgf_gbuffer_texel));
    }
    if (!strcmp(name, "stage_timers")) {        // diagnostic builds (SVGF_STAGE_TIMERS): eight 64-bit clock sums behind the stage kernel's counters
        if (!c->stage_ctr) { memset(host, 0, bytes); return SVGF_OK; }
        return d2h(c->stage_ctr + 8 + SVGF_MAX_LEVELS * (2 * 192 + 8), 64);
    }
    if (!strcmp(name, "bvh_packed")) return d2h(c->scene.bvh, (size_t)c->scene.n_nodes * 32);        // 2 x float4 per node
    if (!strcmp(name, "triangle_ids")) {        // load-order id of the triangle in every slot, in the current (BVH) order
        std::vector<float4> hot(3 * (size_t)c->scene.n_tris);
        if (bytes != (size_t)c->scene.n_tris * 4) { c->err = "svgf_fetch(triangle_ids): size mismatch"; return SVGF_ERR_INVALID; }
        if (!hot.empty()) CK(cudaMemcpy(hot.data(), c->scene.tri_hot, hot.size() * sizeof(float4), cudaMemcpyDeviceToHost));
        for (int i = 0; i < c->scene.n_tris; i++) memcpy(static_cast<char *>(host) + 4 * (size_t)i, &hot[3 * (size_t)i].w, 4);
        return SVGF_OK;
    }
    if (!strcmp(name, "view_matrix_prev")) {
        if (bytes != 64) return SVGF_ERR_INVALID;
        memcpy(host, c->view_matrix_prev, 64);
        return SVGF_OK;
    }
    c->err = std::string("svgf_fetch: unknown buffer '") + name + "'";
    return SVGF_ERR_UNKNOWN_NAME;
}
