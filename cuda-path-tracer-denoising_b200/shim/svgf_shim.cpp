// svgf_shim.cpp -- the reference's six entry points, with the reference's exact C++ signatures, on top of the C ABI.
//
//   void pathtraceInit(Scene*), pathtraceFree(), pathtrace(uchar4* pbo, int frame)      src/pathtrace.h:6-8
//   void denoiseInit(Scene*),   denoiseFree(),   denoise(vec3* out, vec3* in, GBufferTexel*)   src/denoise.h:6-8
//
// Compile this ONE file against the reference's own headers in place of src/pathtrace.cu + src/denoise.cu and link
// libsvgf_b200.so: src/main.cpp, src/preview.cpp and the Scene loader stay untouched. It
//   * hands the Scene's arrays to svgf_create without conversion (the ABI structs are layout-identical: static_asserts),
//   * snapshots the 19 `ui_*` globals the path reads (src/main.h:39-69) into svgf_params on every call,
//   * re-reads scene->state.camera on every call, fills scene->state.image (pathtrace.cu:405,450),
//   * keeps the reference's error behaviour: print and exit(EXIT_FAILURE) (pathtrace.cu:25-43).
// It contains no kernels and no copies of reference code: only the adaptation between the two interfaces.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "main.h"           // reference: Scene, Camera, GBufferTexel, ui_* externs, pathtrace.h, denoise.h
#include "svgf_b200.h"

static_assert(sizeof(Geom) == sizeof(svgf_geom) && sizeof(Material) == sizeof(svgf_material) &&
              sizeof(Triangle) == sizeof(svgf_triangle) && sizeof(BVH_ArrNode) == sizeof(svgf_bvh_node) &&
              sizeof(Camera) == sizeof(svgf_camera) && sizeof(GBufferTexel) == sizeof(svgf_gbuffer_texel) &&
              sizeof(PathSegment) == sizeof(svgf_path_segment) && sizeof(ShadeableIntersection) == sizeof(svgf_intersection),
              "ABI structs must stay layout-identical to the reference's");
static_assert(offsetof(Geom, inverseTransform) == offsetof(svgf_geom, inverseTransform) && offsetof(Geom, T_startidx) == offsetof(svgf_geom, T_startidx) &&
              offsetof(Material, emittance) == offsetof(svgf_material, emittance) && offsetof(Material, texid) == offsetof(svgf_material, texid) &&
              offsetof(BVH_ArrNode, primitive_count) == offsetof(svgf_bvh_node, primitive_count) &&
              offsetof(Camera, pixelLength) == offsetof(svgf_camera, pixelLength) && offsetof(GBufferTexel, geomId) == offsetof(svgf_gbuffer_texel, geomId),
              "field offsets");

static Scene *g_scene = NULL;
static svgf_ctx *g_ctx = NULL;
static void *g_pinned = NULL;       // scene->state.image.data() as page-locked by pathtrace() (pathtrace.cu:450 copies into it)

static void die(const char *what, int rc) {     // == checkCUDAErrorFn, pathtrace.cu:25-43
    fprintf(stderr, "CUDA error (svgf_b200): %s: %d: %s\n", what, rc, svgf_last_error(g_ctx));
    exit(EXIT_FAILURE);
}

static svgf_params snapshot_ui() {
    svgf_params p;
    svgf_params_default(&p);
    p.tracedepth = ui_tracedepth; p.shadowray = ui_shadowray; p.reducevar = ui_reducevar;
    p.sintensity = ui_sintensity; p.lightradius = ui_lightradius;
    p.denoise_enable = ui_denoise_enable; p.sepcolor = ui_sepcolor;
    p.temporal_enable = ui_temporal_enable; p.color_alpha = ui_color_alpha; p.moment_alpha = ui_moment_alpha;
    p.right_view_option = ui_right_view_option; p.atrous_nlevel = ui_atrous_nlevel; p.spatial_enable = ui_spatial_enable;
    p.history_level = ui_history_level; p.sigmal = ui_sigmal; p.sigman = ui_sigman; p.sigmax = ui_sigmax;
    p.blurvariance = ui_blurvariance; p.addcolor = ui_addcolor;
    return p;
}

static svgf_camera snapshot_camera() {
    svgf_camera c;
    memcpy(&c, &g_scene->state.camera, sizeof(c));
    return c;
}

void pathtraceInit(Scene *scene) {
    g_scene = scene;
    std::vector<svgf_texture_desc> tex(scene->textures.size());
    for (size_t i = 0; i < tex.size(); i++) {
        tex[i].width = scene->textures[i].width; tex[i].height = scene->textures[i].height;
        tex[i].components = scene->textures[i].components; tex[i].pixels = scene->textures[i].image;
    }
    svgf_scene_desc d;
    memset(&d, 0, sizeof(d));
    d.geoms = reinterpret_cast<const svgf_geom *>(scene->geoms.data()); d.n_geoms = (int)scene->geoms.size();
    d.materials = reinterpret_cast<const svgf_material *>(scene->materials.data()); d.n_materials = (int)scene->materials.size();
    d.triangles = reinterpret_cast<const svgf_triangle *>(scene->triangles.data()); d.n_triangles = (int)scene->triangles.size();
    d.bvh_nodes = reinterpret_cast<const svgf_bvh_node *>(scene->bvh_nodes); d.n_bvh_nodes = scene->Node_count > 0 ? scene->Node_count : 0;
    d.textures = tex.data(); d.n_textures = (int)tex.size();
    d.width = scene->state.camera.resolution.x; d.height = scene->state.camera.resolution.y;
    int rc = svgf_create(&g_ctx, &d, 0);        // the reference renders on device 0 (preview.cpp:125)
    if (rc) die("pathtraceInit", rc);
}

void pathtraceFree() {          // may precede the first Init (main.cpp:195): no-op on NULL
    svgf_destroy(g_ctx);        // also drops the page-lock on the image buffer
    g_ctx = NULL; g_pinned = NULL;
}

void denoiseInit(Scene *scene) {    // history restarts (denoise.cu:41-60); the buffers live in the same context
    g_scene = scene;
    if (g_ctx) { int rc = svgf_reset(g_ctx); if (rc) die("denoiseInit", rc); }
}

void denoiseFree() {}

void pathtrace(uchar4 *pbo, int frame) {
    const svgf_params p = snapshot_ui();
    const svgf_camera c = snapshot_camera();
    // scene->state.image is a std::vector the Scene owns for its lifetime (sized at load, scene.cpp:170-172): page-lock it once so
    // that the per-frame copy is a direct DMA, and follow it should the vector ever be reallocated
    void *img = g_scene->state.image.data();
    if (img != g_pinned) {
        if (g_pinned) svgf_unregister_host(g_ctx, g_pinned);
        g_pinned = svgf_register_host(g_ctx, img, g_scene->state.image.size() * sizeof(glm::vec3)) == SVGF_OK ? img : NULL;
    }
    int rc = svgf_render(g_ctx, &c, &p, frame, pbo, reinterpret_cast<float *>(img));
    if (rc) die("pathtrace", rc);
}

void denoise(glm::vec3 *output, glm::vec3 *input, GBufferTexel *gbuffer) {
    const svgf_params p = snapshot_ui();
    const svgf_camera c = snapshot_camera();
    int rc = svgf_denoise(g_ctx, reinterpret_cast<float *>(output), reinterpret_cast<const float *>(input),
                          reinterpret_cast<const svgf_gbuffer_texel *>(gbuffer), &c, &p);
    if (rc) die("denoise", rc);
}

// The context behind the six entry points, for callers that want the library's extras (svgf_set_option, svgf_stage_times,
// svgf_fetch ...) next to the reference's interface. NULL before pathtraceInit / after pathtraceFree.
extern "C" svgf_ctx *svgf_shim_context(void) { return g_ctx; }
