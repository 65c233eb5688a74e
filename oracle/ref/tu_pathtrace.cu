// oracle/ref/tu_pathtrace.cu -- TEST INFRASTRUCTURE (oracle/), not product code.
// Compiles the reference's src/pathtrace.cu textually (staged copy, see stage.sh) and adds accessors
// for its file-static device buffers (src/pathtrace.cu:80-101).
#include "pathtrace.cu"
#include <cstring>

extern "C" int refh_fetch_pathtrace(const char *name, void *host, size_t bytes) {
    if (!hst_scene) return -1;
    const Camera &cam = hst_scene->state.camera;
    const size_t px = (size_t)cam.resolution.x * cam.resolution.y;
    const void *src = NULL; size_t need = 0;
    if      (!strcmp(name, "image"))         { src = dev_image;          need = px * 12; }
    else if (!strcmp(name, "denoised"))      { src = dev_denoised_image; need = px * 12; }
    else if (!strcmp(name, "gbuffer"))       { src = dev_gbuffer;        need = px * sizeof(GBufferTexel); }
    else if (!strcmp(name, "intersections")) { src = dev_intersections;  need = px * sizeof(ShadeableIntersection); }
    else if (!strcmp(name, "paths"))         { src = dev_paths;          need = px * sizeof(PathSegment); }
    else return 1;  // not ours
    if (bytes != need || !src) return -2;
    cudaMemcpy(host, src, need, cudaMemcpyDeviceToHost);
    return 0;
}

// Host entry to the reference's own closest-hit routine (src/pathtrace.cu:210-281 is __host__ __device__):
// used by unit tests of the oracle's intersection code, ray by ray.
extern "C" int refh_host_intersect(const float *origin, const float *dir, float *t, float *normal, float *uv,
                                   int *geomId, int *materialId) {
    if (!hst_scene) return -1;
    Ray r; r.origin = glm::vec3(origin[0], origin[1], origin[2]); r.direction = glm::vec3(dir[0], dir[1], dir[2]);
    ShadeableIntersection isect; memset(&isect, 0, sizeof(isect));
    bool hit = computeIntersection(r, isect, hst_scene->geoms.data(), (int)hst_scene->geoms.size(),
                                   hst_scene->triangles.data(), hst_scene->BoudningBoxs.data(), hst_scene->bvh_nodes);
    *t = isect.t; normal[0] = isect.surfaceNormal.x; normal[1] = isect.surfaceNormal.y; normal[2] = isect.surfaceNormal.z;
    uv[0] = isect.uv.x; uv[1] = isect.uv.y; *geomId = isect.geomId; *materialId = isect.materialId;
    return hit ? 1 : 0;
}
