// oracle/ref/harness.cpp -- TEST INFRASTRUCTURE (oracle/), not product code.
//
// Headless driver for the reference's hot path. It plays the role of the reference's src/main.cpp
// (which cannot be linked here: GLFW/GLEW/ImGui) and only that role:
//   * defines the `ui_*` globals the hot path reads at link time (src/main.cpp:37-75, src/main.h:39-69),
//   * restates resetCamera() (src/main.cpp:77-101) and the camera part of runCuda() (src/main.cpp:154-201),
//   * owns a fake PBO (2W x H uchar4) and calls pathtraceFree/Init, denoiseFree/Init, pathtrace(pbo, frame++)
//     in the reference's order.
// It is compiled three ways (oracle/Makefile):
//   libref_gpu*.so   against the reference's own .cu files with nvcc            (runs on the B200 box)
//   libref_cpu*.so   against the same files through cuda_emu/ with g++          (runs here, no GPU)
//   libshim_harness.so against the product's drop-in shim (same six entry points, our kernels)
// and exports one small C API (refh_*) that tests/bench load with ctypes.
//
// Scenes travel as "scene blobs" (tests/golden/scenes/*.scene): the arrays the reference's Scene loader
// produces, in the reference's own struct layouts, because
// /root/reference/scenes does not exist on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <fstream>
#include <sstream>
#include <iostream>
#include <chrono>
#include <unistd.h>

#include "main.h"   // reference header: Scene, Camera, ui_* externs, pathtrace.h, denoise.h

// ---- link-time globals of the reference (values = src/main.cpp:37-75 defaults) -------------------
Scene *scene = NULL;
int frame = 0;
int width = 0, height = 0;
float zoom, theta, phi;
bool camchanged = true;

bool ui_run = true;
bool ui_step = false;
bool ui_reset_denoiser = false;
int ui_tracedepth = 4;
bool ui_shadowray = true;
bool ui_reducevar = true;
float ui_sintensity = 2.7f;
float ui_lightradius = 1.4f;
bool ui_usekdtree = true;
bool ui_denoise_enable = false;
bool ui_temporal_enable = false;
bool ui_spatial_enable = false;
float ui_color_alpha = 0.2;
float ui_moment_alpha = 0.2;
bool ui_blurvariance = true;
float ui_sigmal = 0.45f;
float ui_sigmax = 0.35f;
float ui_sigman = 0.2f;
int ui_atrous_nlevel = 5;
int ui_history_level = 1;
bool ui_sepcolor = false;
bool ui_addcolor = false;
bool ui_automate_camera = false;
float ui_camera_speed_x = 0.0;
float ui_camera_speed_y = 0.0;
float ui_camera_speed_z = 0.0;
float ui_camera_speed_theta = 0.0;
float ui_camera_speed_phi = 0.0;
int ui_left_view_option = 0;
int ui_right_view_option = 0;

#ifdef REFH_CPU_EMU
// storage for the emulated CUDA built-ins declared in cuda_emu/cuda_runtime.h
uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;
#endif

static glm::vec3 cameraPosition;
static glm::vec3 ogLookAt;
static float camera_tx = 0.0f, camera_ty = 0.0f, camera_tz = 0.0f, camera_ttheta = 0.0f, camera_tphi = 0.0f;
static float g_fovy = 45.0f;
static uchar4 *g_pbo = NULL;
static size_t g_pbo_bytes = 0;

extern "C" int refh_fetch_denoise(const char *name, void *host, size_t bytes);
extern "C" int refh_fetch_pathtrace(const char *name, void *host, size_t bytes);

// ---- scene blob ---------------------------------------------------------------------------------
struct BlobHeader {
    char magic[8];  // "SVGFSCN1"
    int n_geoms, n_materials, n_tris, n_bvh, n_boxes, n_textures;
    float fovy;
    int reserved;
};
static_assert(sizeof(BlobHeader) == 40, "blob header");
static_assert(sizeof(Geom) == 248 && sizeof(Material) == 56 && sizeof(Triangle) == 136 &&
              sizeof(BVH_ArrNode) == 40 && sizeof(BoundingBox) == 24 && sizeof(Camera) == 84 &&
              sizeof(GBufferTexel) == 52 && sizeof(PathSegment) == 48 && sizeof(ShadeableIntersection) == 36,
              "reference ABI sizes (SURVEY.md 8(a))");

static Scene *make_empty_scene() {
    char tmpl[] = "/tmp/refh_empty_XXXXXX";
    int fd = mkstemp(tmpl);
    if (fd >= 0) close(fd);
    std::streambuf *old = std::cout.rdbuf();
    std::ostringstream sink; std::cout.rdbuf(sink.rdbuf());
    Scene *s = new Scene(std::string(tmpl));
    std::cout.rdbuf(old);
    unlink(tmpl);
    return s;
}

// Resolution-dependent camera fields, restating src/scene.cpp:159-173.
static void set_resolution(Scene *s, int W, int H, float fovy) {
    Camera &camera = s->state.camera;
    camera.resolution.x = W; camera.resolution.y = H;
    float yscaled = tan(fovy * (PI / 180));
    float xscaled = (yscaled * camera.resolution.x) / camera.resolution.y;
    float fovx = (atan(xscaled) * 180) / PI;
    camera.fov = glm::vec2(fovx, fovy);
    camera.pixelLength = glm::vec2(2 * xscaled / (float)camera.resolution.x, 2 * yscaled / (float)camera.resolution.y);
    s->state.image.resize((size_t)W * H);
    std::fill(s->state.image.begin(), s->state.image.end(), glm::vec3());
}

// src/main.cpp:77-101
static void harness_reset_camera() {
    Camera &cam = scene->state.camera;
    width = cam.resolution.x; height = cam.resolution.y;
    glm::vec3 view = cam.view;
    cameraPosition = cam.position;
    glm::vec3 viewXZ = glm::vec3(view.x, 0.0f, view.z);
    glm::vec3 viewZY = glm::vec3(0.0f, view.y, view.z);
    phi = glm::acos(glm::dot(glm::normalize(viewXZ), glm::vec3(0, 0, -1)));
    theta = glm::acos(glm::dot(glm::normalize(viewZY), glm::vec3(0, 1, 0)));
    ogLookAt = cam.lookAt;
    zoom = glm::length(cam.position - ogLookAt);
    camchanged = true;
    camera_tx = camera_ty = camera_tz = camera_ttheta = camera_tphi = 0.0f;
}

static void alloc_pbo() {
#ifdef REFH_CPU_EMU
    free(g_pbo);
    g_pbo_bytes = (size_t)2 * width * height * sizeof(uchar4);
    g_pbo = (uchar4 *)malloc(g_pbo_bytes);
#else
    cudaFree(g_pbo);
    g_pbo_bytes = (size_t)2 * width * height * sizeof(uchar4);
    cudaMalloc(&g_pbo, g_pbo_bytes);
#endif
}

static void finish_load(int W, int H) {
    harness_reset_camera();
    (void)W; (void)H;
    alloc_pbo();
    frame = 0;
    ui_reset_denoiser = false;
}

extern "C" int refh_load_scene_blob(const char *path, int W, int H) {
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    BlobHeader h;
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "SVGFSCN1", 8)) { fclose(f); return -2; }
    Scene *s = make_empty_scene();
    bool ok = true;
    ok &= fread(&s->state.camera, sizeof(Camera), 1, f) == 1;
    s->geoms.resize(h.n_geoms);          if (h.n_geoms) ok &= fread(s->geoms.data(), sizeof(Geom), h.n_geoms, f) == (size_t)h.n_geoms;
    s->materials.resize(h.n_materials);  if (h.n_materials) ok &= fread(s->materials.data(), sizeof(Material), h.n_materials, f) == (size_t)h.n_materials;
    s->triangles.resize(h.n_tris);       if (h.n_tris) ok &= fread(s->triangles.data(), sizeof(Triangle), h.n_tris, f) == (size_t)h.n_tris;
    s->Node_count = h.n_bvh;
    s->bvh_nodes = h.n_bvh ? new BVH_ArrNode[h.n_bvh] : NULL;
    if (h.n_bvh) ok &= fread(s->bvh_nodes, sizeof(BVH_ArrNode), h.n_bvh, f) == (size_t)h.n_bvh;
    s->BoudningBoxs.resize(h.n_boxes);   if (h.n_boxes) ok &= fread(s->BoudningBoxs.data(), sizeof(BoundingBox), h.n_boxes, f) == (size_t)h.n_boxes;
    for (int i = 0; i < h.n_textures && ok; i++) {
        int whc[3];
        ok &= fread(whc, sizeof(int), 3, f) == 3;
        Texture t; t.width = whc[0]; t.height = whc[1]; t.components = whc[2];
        size_t n = (size_t)whc[0] * whc[1] * whc[2];
        t.image = (unsigned char *)malloc(n); t.dev_image = NULL;
        ok &= fread(t.image, 1, n, f) == n;
        s->textures.push_back(t);
    }
    fclose(f);
    if (!ok) return -3;
    // src/scene.cpp:313-324 (loadLight)
    for (size_t i = 0; i < s->geoms.size(); i++) {
        if (s->materials[s->geoms[i].materialid].emittance > 0) {
            Light light; light.geomIdx = (int)i; light.matIdx = s->geoms[i].materialid;
            light.type = LightType::AREALIGHT; light.geom = s->geoms[i];
            s->lights.push_back(light);
        }
    }
    g_fovy = h.fovy;
    set_resolution(s, W, H, g_fovy);
    scene = s;
    finish_load(W, H);
    return 0;
}

#ifdef REFH_CPU_EMU
// Uses the reference's own text-scene loader (src/scene.cpp). Only possible where /root/reference exists:
// the loader hard-codes ../scenes/Models and ../scenes/Textures relative to the CWD (src/scene.cpp:220,236),
// so we chdir into <reference>/src for the duration of the load. The RES line is overridden through a
// temporary copy of the scene file, as the reference has no other way to set the resolution.
extern "C" int refh_load_scene_txt(const char *reference_root, const char *scene_name, int W, int H) {
    std::string src = std::string(reference_root) + "/scenes/" + scene_name;
    std::ifstream in(src);
    if (!in.is_open()) return -1;
    char tmpl[] = "/tmp/refh_scene_XXXXXX";
    int fd = mkstemp(tmpl);
    if (fd < 0) return -2;
    close(fd);
    {
        std::ofstream out(tmpl);
        std::string line;
        while (std::getline(in, line)) {
            if (line.rfind("RES", 0) == 0) {
                bool cr = !line.empty() && line.back() == '\r';
                out << "RES         " << W << " " << H << (cr ? "\r" : "") << "\n";
            } else out << line << "\n";
        }
    }
    char cwd[4096];
    if (!getcwd(cwd, sizeof(cwd))) return -3;
    if (chdir((std::string(reference_root) + "/src").c_str()) != 0) return -4;
    std::streambuf *old = std::cout.rdbuf();
    std::ostringstream sink; std::cout.rdbuf(sink.rdbuf());
    Scene *s = new Scene(std::string(tmpl));
    std::cout.rdbuf(old);
    if (chdir(cwd) != 0) return -5;
    unlink(tmpl);
    // FOVY is not kept by the loader; recover it from camera.fov.y (src/scene.cpp:162).
    g_fovy = s->state.camera.fov.y;
    scene = s;
    finish_load(W, H);
    return 0;
}

// Writes the loaded scene as a blob. Fields the reference leaves uninitialised and the hot path never
// reads (Material::norid, Material::specular.exponent is read from file, Triangle::normal/boundingbox,
// mesh indices of non-mesh geoms) are zeroed so blobs are reproducible.
extern "C" int refh_export_scene(const char *out_path) {
    if (!scene) return -1;
    FILE *f = fopen(out_path, "wb");
    if (!f) return -2;
    BlobHeader h; memset(&h, 0, sizeof(h));
    memcpy(h.magic, "SVGFSCN1", 8);
    h.n_geoms = (int)scene->geoms.size(); h.n_materials = (int)scene->materials.size();
    h.n_tris = (int)scene->triangles.size(); h.n_bvh = scene->Node_count > 0 ? scene->Node_count : 0;
    h.n_boxes = (int)scene->BoudningBoxs.size(); h.n_textures = (int)scene->textures.size();
    h.fovy = g_fovy;
    fwrite(&h, sizeof(h), 1, f);
    // camera as the *loader* leaves it, except right (NaN there, src/scene.cpp:164) which is zeroed.
    Camera cam = scene->state.camera;
    cam.right = glm::vec3(0.0f);
    cam.resolution = glm::ivec2(0, 0); cam.fov = glm::vec2(0.0f); cam.pixelLength = glm::vec2(0.0f);
    fwrite(&cam, sizeof(Camera), 1, f);
    for (Geom g : scene->geoms) {
        if (g.type != MESH) { g.T_startidx = 0; g.T_endidx = 0; g.BoundIdx = 0; }
        fwrite(&g, sizeof(Geom), 1, f);
    }
    for (Material m : scene->materials) { m.norid = 0; fwrite(&m, sizeof(Material), 1, f); }
    for (Triangle t : scene->triangles) {
        t.normal = glm::vec3(0.0f); t.boundingbox.maxCorner = glm::vec3(0.0f); t.boundingbox.minCorner = glm::vec3(0.0f);
        fwrite(&t, sizeof(Triangle), 1, f);
    }
    for (int i = 0; i < h.n_bvh; i++) {
        BVH_ArrNode n = scene->bvh_nodes[i];
        if (n.primitive_count > 0) { n.axis = 0; n.rightchildoffset = 0; } else { n.primitivesOffset = 0; }
        fwrite(&n, sizeof(BVH_ArrNode), 1, f);
    }
    if (h.n_boxes) fwrite(scene->BoudningBoxs.data(), sizeof(BoundingBox), h.n_boxes, f);
    for (const Texture &t : scene->textures) {
        int whc[3] = { t.width, t.height, t.components };
        fwrite(whc, sizeof(int), 3, f);
        fwrite(t.image, 1, (size_t)t.width * t.height * t.components, f);
    }
    fclose(f);
    return 0;
}
#endif

// ---- parameters ---------------------------------------------------------------------------------
extern "C" int refh_set_param(const char *name, double v) {
#define P_BOOL(n)  if (!strcmp(name, #n)) { ui_##n = (v != 0.0); return 0; }
#define P_INT(n)   if (!strcmp(name, #n)) { ui_##n = (int)v; return 0; }
#define P_FLT(n)   if (!strcmp(name, #n)) { ui_##n = (float)v; return 0; }
    P_INT(tracedepth) P_BOOL(shadowray) P_BOOL(reducevar) P_FLT(sintensity) P_FLT(lightradius)
    P_BOOL(denoise_enable) P_BOOL(temporal_enable) P_BOOL(spatial_enable) P_FLT(color_alpha) P_FLT(moment_alpha)
    P_BOOL(blurvariance) P_FLT(sigmal) P_FLT(sigmax) P_FLT(sigman) P_INT(atrous_nlevel) P_INT(history_level)
    P_BOOL(sepcolor) P_BOOL(addcolor) P_BOOL(automate_camera) P_FLT(camera_speed_x) P_FLT(camera_speed_y)
    P_FLT(camera_speed_z) P_FLT(camera_speed_theta) P_FLT(camera_speed_phi) P_INT(right_view_option)
    P_BOOL(reset_denoiser)
#undef P_BOOL
#undef P_INT
#undef P_FLT
    return -1;
}

// ---- one iteration of the reference's runCuda() (src/main.cpp:154-209), minus GL ------------------
extern "C" int refh_frame() {
    if (!scene) return -1;
    RenderState *renderState = &scene->state;
    if (ui_automate_camera) {
        Camera &cam = renderState->camera;
        camera_tx += ui_camera_speed_x;
        camera_ty += ui_camera_speed_y;
        camera_tz += ui_camera_speed_z;
        camera_ttheta += ui_camera_speed_theta;
        camera_tphi += ui_camera_speed_phi;
        cam.lookAt.x = 0.0f + 2.0f * sinf(camera_tx);
        cam.lookAt.y = 5.0f + 1.0f * sinf(camera_ty);
        cam.lookAt.z = 0.0f + 1.5f * sinf(camera_tz);
        theta = PI * 0.5f + PI / 18 * sinf(camera_ttheta);
        phi   = PI * 0.0f + PI / 12 * sinf(camera_tphi);
        camchanged = true;
    }
    if (camchanged) {
        if (!ui_denoise_enable) frame = 0;
        Camera &cam = renderState->camera;
        cameraPosition.x = zoom * sin(phi) * sin(theta);
        cameraPosition.y = zoom * cos(theta);
        cameraPosition.z = zoom * cos(phi) * sin(theta);
        cam.view = -glm::normalize(cameraPosition);
        glm::vec3 v = cam.view;
        glm::vec3 u = glm::vec3(0, 1, 0);
        glm::vec3 r = glm::cross(v, u);
        cam.up = glm::cross(r, v);
        cam.right = r;
        cam.position = cameraPosition;
        cameraPosition += cam.lookAt;
        cam.position = cameraPosition;
        camchanged = false;
    }
    ui_reset_denoiser |= (frame == 0);
    if (ui_reset_denoiser == true) {
        pathtraceFree();
        pathtraceInit(scene);
        denoiseFree();
        denoiseInit(scene);
        frame = 0;
        ui_reset_denoiser = false;
    }
    int rendered = frame;
    pathtrace(g_pbo, frame++);
    return rendered;
}

// ---- the public denoise() entry point on its own (src/denoise.h:8, denoise.cu:349-402) -------------
// refh_init: the reference's (re)initialisation order without rendering a frame. refh_denoise_host: uploads a 1-spp colour
// buffer and a G-buffer in the reference's AoS layouts, calls denoise(output, input, gbuffer) -- the reference's own, or the
// drop-in shim's, whichever this library was linked with -- and returns `output`. The camera is whatever
// refh_set_camera / refh_frame left in scene->state.camera (denoise.cu:350 re-reads it on every call).
extern "C" int refh_init() {
    if (!scene) return -1;
    pathtraceFree(); pathtraceInit(scene); denoiseFree(); denoiseInit(scene);
    frame = 0; ui_reset_denoiser = false;
    return 0;
}
extern "C" int refh_set_camera(const void *cam84) {
    if (!scene) return -1;
    memcpy(&scene->state.camera, cam84, sizeof(Camera));
    return 0;
}
extern "C" int refh_denoise_host(void *out, const void *in, const void *gbuffer) {
    if (!scene) return -1;
    const size_t px = (size_t)scene->state.camera.resolution.x * scene->state.camera.resolution.y;
    static glm::vec3 *d_in = NULL, *d_out = NULL; static GBufferTexel *d_g = NULL; static size_t cap = 0;
    if (cap != px) {
        cudaFree(d_in); cudaFree(d_out); cudaFree(d_g);
        cudaMalloc((void **)&d_in, px * sizeof(glm::vec3)); cudaMalloc((void **)&d_out, px * sizeof(glm::vec3));
        cudaMalloc((void **)&d_g, px * sizeof(GBufferTexel));
        cap = px;
    }
    cudaMemcpy(d_in, in, px * sizeof(glm::vec3), cudaMemcpyHostToDevice);
    cudaMemcpy(d_g, gbuffer, px * sizeof(GBufferTexel), cudaMemcpyHostToDevice);
    denoise(d_out, d_in, d_g);
    cudaDeviceSynchronize();
    cudaMemcpy(out, d_out, px * sizeof(glm::vec3), cudaMemcpyDeviceToHost);
    return 0;
}

extern "C" int refh_fetch(const char *name, void *host, size_t bytes) {
    if (!scene) return -1;
    if (!strcmp(name, "pbo")) {
        if (bytes != g_pbo_bytes) return -2;
        cudaMemcpy(host, g_pbo, bytes, cudaMemcpyDeviceToHost);
        return 0;
    }
    if (!strcmp(name, "host_image")) {
        size_t need = scene->state.image.size() * sizeof(glm::vec3);
        if (bytes != need) return -2;
        memcpy(host, scene->state.image.data(), need);
        return 0;
    }
    if (!strcmp(name, "camera")) {
        if (bytes != sizeof(Camera)) return -2;
        memcpy(host, &scene->state.camera, sizeof(Camera));
        return 0;
    }
    int r = refh_fetch_pathtrace(name, host, bytes);
    if (r != 1) return r;
    r = refh_fetch_denoise(name, host, bytes);
    if (r != 1) return r;
    return -3;
}

// Wall time of n consecutive frames (ms). pathtrace() ends with a blocking D2H copy
// (src/pathtrace.cu:450), so the host clock brackets all device work.
extern "C" double refh_time_frames(int n) {
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n; i++) refh_frame();
    cudaDeviceSynchronize();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

extern "C" int refh_scene_counts(int *out6) {
    if (!scene) return -1;
    out6[0] = (int)scene->geoms.size(); out6[1] = (int)scene->materials.size(); out6[2] = (int)scene->triangles.size();
    out6[3] = scene->Node_count; out6[4] = (int)scene->BoudningBoxs.size(); out6[5] = (int)scene->textures.size();
    return 0;
}
