// oracle/svgf_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Scalar CPU restatement of the reference's per-frame hot path. Every function names the reference
// lines it follows (paths relative to /root/reference). The arithmetic is written in the reference's
// expression order *including the order glm 0.9.6.3 evaluates its helpers in* (external/include/glm):
//   dot(vec3)      (x*x' + y*y') + z*z'                          detail/func_geometric.inl:65-72
//   normalize(v)   v * (1 / sqrt(dot(v,v)))                       detail/func_geometric.inl:154-159, func_exponential.inl:150-153
//   cross          (y z' - y' z, z x' - z' x, x y' - x' y)       detail/func_geometric.inl:134-142
//   mat4*vec4      (m0 v0 + m1 v1) + (m2 v2 + m3 v3)             detail/type_mat4x4.inl:617-628
//   min/max        x<y?x:y / x>y?x:y                             detail/func_common.inl:409-435
// so that, compiled with -ffp-contract=off, it is bit-identical to the reference's own code run on the CPU
// (oracle/_ref/libref_cpu*.so). Pinned by tests/test_oracle_vs_reference.py; GPU goldens in tests/golden/.
//
// Only tests/, bench.py (cpu_baseline / reference legs) and __graft_entry__.smoke() may use this library.
#include "svgf_oracle.h"

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

static_assert(sizeof(svgf_geom) == 248 && sizeof(svgf_material) == 56 && sizeof(svgf_triangle) == 136 &&
              sizeof(svgf_bvh_node) == 40 && sizeof(svgf_camera) == 84 && sizeof(svgf_gbuffer_texel) == 52 &&
              sizeof(svgf_path_segment) == 48 && sizeof(svgf_intersection) == 36, "reference ABI (SURVEY 8(a))");

namespace {

// ---- utilities.h:12-24 ----------------------------------------------------------------------------
const float PI_F = 3.1415926535897932384626422832795028841971f;
const float TWO_PI_F = 6.2831853071795864769252867665590057683943f;
const float SQRT_OF_ONE_THIRD_F = 0.5773502691896257645091487805019574556476f;
const float COLORDIVIDOR_F = 0.003921568627f;

struct V3 { float x, y, z; };
inline V3 mk(float x, float y, float z) { V3 r = {x, y, z}; return r; }
inline V3 ld(const float *p) { return mk(p[0], p[1], p[2]); }
inline void st(float *p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
inline V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
inline V3 operator/(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
inline float dot(V3 a, V3 b) { V3 t = a * b; return t.x + t.y + t.z; }
inline V3 cross(V3 x, V3 y) { return mk(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
inline float length(V3 v) { return sqrtf(dot(v, v)); }
inline V3 normalize(V3 v) { return v * (1.0f / sqrtf(dot(v, v))); }
inline float gmin(float x, float y) { return x < y ? x : y; }
inline float gmax(float x, float y) { return x > y ? x : y; }
inline float gabs(float x) { return x >= 0.0f ? x : -x; }

// glm mat4 (column-major float[16]) times vec4, keeping xyz (intersections.h:36-38 multiplyMV)
inline V3 multiplyMV(const float *m, V3 v, float w) {
    float o[3];
    for (int r = 0; r < 3; r++) {
        float mul0 = m[0 + r] * v.x, mul1 = m[4 + r] * v.y;
        float add0 = mul0 + mul1;
        float mul2 = m[8 + r] * v.z, mul3 = m[12 + r] * w;
        float add1 = mul2 + mul3;
        o[r] = add0 + add1;
    }
    return mk(o[0], o[1], o[2]);
}
inline void mulM4V4(const float *m, const float v[4], float o[4]) {
    for (int r = 0; r < 4; r++) {
        float add0 = m[0 + r] * v[0] + m[4 + r] * v[1];
        float add1 = m[8 + r] * v[2] + m[12 + r] * v[3];
        o[r] = add0 + add1;
    }
}

struct Ray { V3 origin, direction; };

// ---- interactions.h:10-30 -------------------------------------------------------------------------
inline unsigned int initRand(unsigned int val0, unsigned int val1, unsigned int backoff) {
    unsigned int v0 = val0, v1 = val1, s0 = 0;
    for (unsigned int n = 0; n < backoff; n++) {
        s0 += 0x9e3779b9;
        v0 += ((v1 << 4) + 0xa341316c) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4);
        v1 += ((v0 << 4) + 0xad90777d) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761e);
    }
    return v0;
}
inline float nextRand(unsigned int &s) {
    s = (1664525u * s + 1013904223u);
    return float(s & 0x00FFFFFF) / float(0x01000000);
}

}  // namespace

struct orc_scene {
    std::vector<svgf_geom> geoms;
    std::vector<svgf_material> materials;
    std::vector<svgf_triangle> tris;
    std::vector<svgf_bvh_node> bvh;
    struct Tex { int w, h, c; std::vector<unsigned char> px; };
    std::vector<Tex> textures;
    svgf_camera cam;
    float fovy;
    int n_boxes;
};

struct orc_state {
    const orc_scene *scene;
    int W, H;
    // pathtrace.cu:80-101
    std::vector<float> image, denoised;
    std::vector<svgf_gbuffer_texel> gbuffer;
    std::vector<svgf_intersection> intersections;
    std::vector<unsigned char> pbo;
    // denoise.cu:14-27
    std::vector<float> temp[2], color_history, color_acc, moment_history, moment_acc, variance, variance_tmp;
    std::vector<int> history_length, history_length_update;
    std::vector<svgf_gbuffer_texel> gbuffer_prev;
    float view_matrix_prev[16];
};

namespace {

inline V3 getPointOnRay(const Ray &r, float t) {   // intersections.h:29-31
    return r.origin + (t - .0001f) * normalize(r.direction);
}

// intersections.h:50-92
float boxIntersectionTest(const svgf_geom &box, const Ray &r, V3 &intersectionPoint, V3 &normal) {
    Ray q;
    q.origin = multiplyMV(box.inverseTransform, r.origin, 1.0f);
    q.direction = normalize(multiplyMV(box.inverseTransform, r.direction, 0.0f));
    float tmin = -1e38f, tmax = 1e38f;
    V3 tmin_n = mk(0, 0, 0), tmax_n = mk(0, 0, 0);
    const float qo[3] = {q.origin.x, q.origin.y, q.origin.z}, qd[3] = {q.direction.x, q.direction.y, q.direction.z};
    for (int xyz = 0; xyz < 3; ++xyz) {
        float qdxyz = qd[xyz];
        float t1 = (-0.5f - qo[xyz]) / qdxyz;
        float t2 = (+0.5f - qo[xyz]) / qdxyz;
        float ta = gmin(t1, t2);
        float tb = gmax(t1, t2);
        float nn[3] = {0, 0, 0};
        nn[xyz] = t2 < t1 ? +1 : -1;
        V3 n = mk(nn[0], nn[1], nn[2]);
        if (ta > 0 && ta > tmin) { tmin = ta; tmin_n = n; }
        if (tb < tmax) { tmax = tb; tmax_n = n; }
    }
    if (tmax >= tmin && tmax > 0) {
        if (tmin <= 0) { tmin = tmax; tmin_n = tmax_n; }
        intersectionPoint = multiplyMV(box.transform, getPointOnRay(q, tmin), 1.0f);
        normal = normalize(multiplyMV(box.transform, tmin_n, 0.0f));
        return length(r.origin - intersectionPoint);
    }
    return -1;
}

// intersections.h:104-146
float sphereIntersectionTest(const svgf_geom &sphere, const Ray &r, V3 &intersectionPoint, V3 &normal) {
    float radius = .5;
    Ray rt;
    rt.origin = multiplyMV(sphere.inverseTransform, r.origin, 1.0f);
    rt.direction = normalize(multiplyMV(sphere.inverseTransform, r.direction, 0.0f));
    float vDotDirection = dot(rt.origin, rt.direction);
    float radicand = vDotDirection * vDotDirection - (dot(rt.origin, rt.origin) - powf(radius, 2));
    if (radicand < 0) return -1;
    float squareRoot = sqrtf(radicand);
    float firstTerm = -vDotDirection;
    float t1 = firstTerm + squareRoot;
    float t2 = firstTerm - squareRoot;
    float t = 0;
    bool outside;
    if (t1 < 0 && t2 < 0) return -1;
    else if (t1 > 0 && t2 > 0) { t = fminf(t1, t2); outside = true; }
    else { t = fmaxf(t1, t2); outside = false; }
    V3 objspaceIntersection = getPointOnRay(rt, t);
    intersectionPoint = multiplyMV(sphere.transform, objspaceIntersection, 1.f);
    normal = normalize(multiplyMV(sphere.invTranspose, objspaceIntersection, 0.f));
    if (!outside) normal = -normal;
    return length(r.origin - intersectionPoint);
}

struct TriHit { float t; V3 n; float uv[2]; };

// sceneStructs.h:157-180 (Triangle::Intersect) over glm::intersectRayTriangle, gtx/intersect.inl:37-74
bool triangleIntersect(const svgf_triangle &tri, const Ray &r, TriHit &isect) {
    V3 v0 = ld(tri.verts[0].pos), v1 = ld(tri.verts[1].pos), v2 = ld(tri.verts[2].pos);
    V3 e1 = v1 - v0, e2 = v2 - v0;
    V3 p = cross(r.direction, e2);
    float a = dot(e1, p);
    float Epsilon = std::numeric_limits<float>::epsilon();
    if (a < Epsilon) { isect.t = -1.0f; return false; }
    float f = 1.0f / a;
    V3 s = r.origin - v0;
    float bx = f * dot(s, p);
    if (bx < 0.0f) { isect.t = -1.0f; return false; }
    if (bx > 1.0f) { isect.t = -1.0f; return false; }
    V3 q = cross(s, e1);
    float by = f * dot(r.direction, q);
    if (by < 0.0f) { isect.t = -1.0f; return false; }
    if (by + bx > 1.0f) { isect.t = -1.0f; return false; }
    float bz = f * dot(e2, q);
    if (!(bz >= 0.0f)) { isect.t = -1.0f; return false; }
    isect.t = bz;
    // uv: correct barycentric order; normal: permuted order (sceneStructs.h:162-170), both preserved
    const float w0 = 1.0f - bx - by;
    isect.uv[0] = (tri.verts[0].uv[0] * w0 + tri.verts[1].uv[0] * bx) + tri.verts[2].uv[0] * by;
    isect.uv[1] = (tri.verts[0].uv[1] * w0 + tri.verts[1].uv[1] * bx) + tri.verts[2].uv[1] * by;
    const float wn = 1.f - bx - by;
    V3 n = (ld(tri.verts[0].normal) * bx + ld(tri.verts[1].normal) * by) + ld(tri.verts[2].normal) * wn;
    isect.n = normalize(n);
    return true;
}

// boundingbox.h:62-79
inline bool AABBIntersect2(const svgf_bvh_node &nd, const Ray &ray, V3 invDir) {
    float txMin = (nd.bounds_min[0] - ray.origin.x) * invDir.x;
    float txMax = (nd.bounds_max[0] - ray.origin.x) * invDir.x;
    float tyMin = (nd.bounds_min[1] - ray.origin.y) * invDir.y;
    float tyMax = (nd.bounds_max[1] - ray.origin.y) * invDir.y;
    float tzMin = (nd.bounds_min[2] - ray.origin.z) * invDir.z;
    float tzMax = (nd.bounds_max[2] - ray.origin.z) * invDir.z;
    float tmin = gmax(gmax(gmin(txMin, txMax), gmin(tyMin, tyMax)), gmin(tzMin, tzMax));
    float tmax = gmin(gmin(gmax(txMin, txMax), gmax(tyMin, tyMax)), gmax(tzMin, tzMax));
    if (tmax < 0) return false;
    if (tmin > tmax) return false;
    return true;
}

// intersections.h:265-329
bool IntersectBVH(const Ray &ray, TriHit *isect, int &hit_tri_index, const svgf_bvh_node *nodes, const svgf_triangle *prims) {
    if (nodes == nullptr) return false;
    bool hit = false;
    int isDirNeg[3] = {ray.direction.x < 0.f, ray.direction.y < 0.f, ray.direction.z < 0.f};
    V3 invdir = mk(1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z);
    int toVisitOffset = 0, curr_ind = 0;
    int needToVisit[64];
    while (true) {
        const svgf_bvh_node *node = &nodes[curr_ind];
        if (AABBIntersect2(*node, ray, invdir)) {
            if (node->primitive_count > 0) {
                for (int i = 0; i < node->primitive_count; i++) {
                    TriHit inter;
                    if (triangleIntersect(prims[node->primitivesOffset + i], ray, inter)) {
                        hit = true;
                        if (isect->t == -1.0f) { *isect = inter; hit_tri_index = prims[node->primitivesOffset + i].id; }
                        else if (inter.t < isect->t) { *isect = inter; hit_tri_index = prims[node->primitivesOffset + i].id; }
                    }
                }
                if (toVisitOffset == 0) break;
                curr_ind = needToVisit[--toVisitOffset];
            } else {
                if (toVisitOffset == 64) { curr_ind = needToVisit[--toVisitOffset]; continue; }
                if (isDirNeg[node->axis]) { needToVisit[toVisitOffset++] = curr_ind + 1; curr_ind = node->rightchildoffset; }
                else { needToVisit[toVisitOffset++] = node->rightchildoffset; curr_ind = curr_ind + 1; }
            }
        } else {
            if (toVisitOffset == 0) break;
            curr_ind = needToVisit[--toVisitOffset];
        }
    }
    return hit;
}

// pathtrace.cu:210-281. Only t/geomId are written on a miss; the rest of `intersection` stays stale.
bool computeIntersection(const orc_scene &sc, const Ray &ray, svgf_intersection &intersection) {
    float t_min = FLT_MAX;
    int hit_geom_index = -1;
    int hit_tri_index = -1;
    V3 normal = mk(0, 0, 0);
    float uv[2] = {0, 0};
    float t;
    V3 tmp_intersect = mk(0, 0, 0), tmp_normal = mk(0, 0, 0);
    float tmp_uv[2] = {0, 0};
    const int geoms_size = (int)sc.geoms.size();
    for (int i = 0; i < geoms_size; i++) {
        const svgf_geom &geom = sc.geoms[i];
        if (geom.type == 1) t = boxIntersectionTest(geom, ray, tmp_intersect, tmp_normal);
        else if (geom.type == 0) t = sphereIntersectionTest(geom, ray, tmp_intersect, tmp_normal);
        else if (geom.type == 2) {
            TriHit isect; isect.t = FLT_MAX; isect.n = mk(0, 0, 0); isect.uv[0] = isect.uv[1] = 0;
            t = -1.0f;
            if (IntersectBVH(ray, &isect, hit_tri_index, sc.bvh.empty() ? nullptr : sc.bvh.data(), sc.tris.data())) {
                if (hit_tri_index >= geom.T_startidx && hit_tri_index < geom.T_endidx) {
                    t = isect.t; tmp_uv[0] = isect.uv[0]; tmp_uv[1] = isect.uv[1]; tmp_normal = isect.n;
                }
            }
        }
        if (t > 0.0f && t < t_min) {
            t_min = t; hit_geom_index = i; normal = tmp_normal; uv[0] = tmp_uv[0]; uv[1] = tmp_uv[1];
        }
    }
    if (hit_geom_index == -1) {
        intersection.t = -1.0f;
        intersection.geomId = -1;
        return false;
    }
    intersection.t = t_min;
    intersection.materialId = sc.geoms[hit_geom_index].materialid;
    st(intersection.surfaceNormal, normal);
    intersection.uv[0] = uv[0]; intersection.uv[1] = uv[1];
    intersection.geomId = hit_geom_index;
    return true;
}

// sceneStructs.h:208-221
V3 textureColor(const orc_scene::Tex &tx, const float uv[2]) {
    int X = (int)gmin(1.f * tx.w * uv[0], 1.f * tx.w - 1.0f);
    int Y = (int)gmin(1.f * tx.h * (1.0f - uv[1]), 1.f * tx.h - 1.0f);
    int texel_index = Y * tx.w + X;
    if (tx.c == 3) {
        // the reference reads out of bounds for uv < 0 (undefined); clamp the index so the oracle cannot crash
        long n = (long)tx.w * tx.h;
        if (texel_index < 0) texel_index = 0;
        if (texel_index >= n) texel_index = (int)n - 1;
        V3 col = mk((float)tx.px[texel_index * 3], (float)tx.px[texel_index * 3 + 1], (float)tx.px[texel_index * 3 + 2]);
        return COLORDIVIDOR_F * col;
    }
    return mk(0, 0, 0);
}
inline V3 materialAlbedo(const orc_scene &sc, const svgf_material &m, const float uv[2]) {
    return m.texid == -1 ? ld(m.color) : textureColor(sc.textures[m.texid], uv);
}

// pathtrace.cu:284-297 with glm::rotation (gtx/quaternion.inl:248-283) and quat*vec3 (gtc/quaternion.inl:319-326)
void computeShadowRay(Ray &shadowRay, V3 originPos, const svgf_geom &light, float lightRadius, float &expectDist, unsigned int &seed) {
    V3 lt = ld(light.translation);
    V3 directionToCenter = normalize(lt - originPos);
    const V3 orig = mk(0.0f, 0.0f, 1.0f);
    float qw, qx, qy, qz;
    float cosTheta = dot(orig, directionToCenter);
    const float eps = std::numeric_limits<float>::epsilon();
    if (cosTheta < -1.0f + eps) {
        V3 axis = cross(mk(0, 0, 1), orig);
        if (dot(axis, axis) < eps) axis = cross(mk(1, 0, 0), orig);
        axis = normalize(axis);
        const float a = 3.14159265358979323846264338327950288f;   // glm::pi<float>()
        const float s = sinf(a * 0.5f);
        qw = cosf(a * 0.5f); qx = axis.x * s; qy = axis.y * s; qz = axis.z * s;
    } else {
        V3 axis = cross(orig, directionToCenter);
        float s = sqrtf((1.0f + cosTheta) * 2.0f);
        float invs = 1.0f / s;
        qw = s * 0.5f; qx = axis.x * invs; qy = axis.y * invs; qz = axis.z * invs;
    }
    float theta = 2 * PI_F * nextRand(seed);
    V3 v = mk(cosf(theta), sinf(theta), 0.0f);
    V3 qv = mk(qx, qy, qz);
    V3 uv = cross(qv, v);
    V3 uuv = cross(qv, uv);
    V3 sampleDirection = v + ((uv * qw) + uuv) * 2.0f;
    float sampleRadius = nextRand(seed) * lightRadius;
    V3 samplePoint = lt + sampleDirection * sampleRadius;
    expectDist = length(samplePoint - originPos);
    shadowRay.origin = originPos;
    shadowRay.direction = normalize(samplePoint - originPos);
}

// interactions.h:37-67
V3 calculateRandomDirectionInHemisphere(V3 normal, unsigned int &seed) {
    float up = sqrtf(nextRand(seed));
    float over = sqrtf(1 - up * up);
    float around = nextRand(seed) * TWO_PI_F;
    V3 directionNotNormal;
    if (fabsf(normal.x) < SQRT_OF_ONE_THIRD_F) directionNotNormal = mk(1, 0, 0);
    else if (fabsf(normal.y) < SQRT_OF_ONE_THIRD_F) directionNotNormal = mk(0, 1, 0);
    else directionNotNormal = mk(0, 0, 1);
    V3 p1 = normalize(cross(normal, directionNotNormal));
    V3 p2 = normalize(cross(normal, p1));
    return (up * normal + cosf(around) * over * p1) + sinf(around) * over * p2;
}

struct Segment { Ray ray; V3 color; int pixelIndex; int remainingBounces; bool diffuse, specular; };

// interactions.h:94-136
void scatterRay(Segment &ps, V3 intersect, V3 normal, const svgf_material &m, unsigned int &seed) {
    ps.specular = false;
    ps.ray.origin = intersect + 1e-4f * normal;
    ps.remainingBounces--;
    if (m.hasRefractive) {
        float eta = 1.0f / m.indexOfRefraction;
        float unit_projection = dot(ps.ray.direction, normal);
        if (unit_projection > 0) eta = 1.0f / eta;
        float R0 = powf((1.0f - eta) / (1.0f + eta), 2.0f);
        float R = R0 + (1 - R0) * powf(1 - gabs(unit_projection), 5.0f);
        if (R < nextRand(seed)) {
            // glm::refract, detail/func_geometric.inl:189-198
            V3 I = ps.ray.direction, N = normal;
            float dotValue = dot(N, I);
            float k = 1.0f - eta * eta * (1.0f - dotValue * dotValue);
            ps.ray.direction = (eta * I - (eta * dotValue + sqrtf(k)) * N) * (float)(k >= 0.0f);
        } else {
            V3 I = ps.ray.direction, N = normal;
            ps.ray.direction = I - N * dot(N, I) * 2.0f;     // glm::reflect
            ps.color = ps.color * ld(m.specular_color);
            ps.specular = true;
        }
    } else if (nextRand(seed) < m.hasReflective) {
        V3 I = ps.ray.direction, N = normal;
        ps.ray.direction = I - N * dot(N, I) * 2.0f;
        ps.color = ps.color * ld(m.specular_color);
        ps.specular = true;
    } else {
        ps.ray.direction = calculateRandomDirectionInHemisphere(normal, seed);
        ps.diffuse = true;
    }
}

// pathtrace.cu:187-208 + 300-401, one pixel
void tracePixel(orc_state &S, const svgf_camera &cam, const svgf_params &P, int frame, int x, int y) {
    const orc_scene &sc = *S.scene;
    const int W = cam.resolution[0], Hh = cam.resolution[1];
    const int idx = x + y * W;
    Segment segment;
    segment.ray.origin = ld(cam.position);
    segment.color = mk(1.0f, 1.0f, 1.0f);
    segment.ray.direction = normalize(ld(cam.view)
        - ld(cam.right) * cam.pixelLength[0] * ((float)x - (float)(W * 0.5f - 0.5f))
        - ld(cam.up) * cam.pixelLength[1] * ((float)y - (float)(Hh * 0.5f - 0.5f)));
    segment.pixelIndex = idx;
    segment.remainingBounces = P.tracedepth;
    segment.diffuse = false; segment.specular = false;

    svgf_intersection &intersection = S.intersections[idx];
    V3 accumulatedColor = mk(0, 0, 0);
    bool hit = computeIntersection(sc, segment.ray, intersection);
    {
        const svgf_material &material = sc.materials[intersection.materialId];
        svgf_gbuffer_texel &g = S.gbuffer[idx];
        st(g.position, segment.ray.origin + intersection.t * segment.ray.direction);
        st(g.normal, ld(intersection.surfaceNormal));
        g.geomId = intersection.geomId;
        st(g.albedo, materialAlbedo(sc, material, intersection.uv));
        st(g.ialbedo, mk(1.0f, 1.0f, 1.0f));
    }
    const bool trace_shadowray = P.shadowray, reduce_var = P.reducevar, denoise = P.denoise_enable, sepcolor = P.sepcolor;
    for (int depth = 1; depth <= P.tracedepth; depth++) {
        if (!hit) break;
        unsigned int seed = initRand(idx, frame + depth, 16);
        const svgf_material &material = sc.materials[intersection.materialId];
        if (material.emittance > 0.0f) {
            if (!trace_shadowray || !reduce_var || !segment.diffuse)
                accumulatedColor = accumulatedColor + segment.color * ld(material.color) * material.emittance;
            break;
        } else {
            V3 intersectionPos = segment.ray.origin + intersection.t * segment.ray.direction;
            V3 intersectionNormal = ld(intersection.surfaceNormal);
            bool materialIsDiffuse = material.hasReflective < 1e-6 && material.hasRefractive < 1e-6;
            if (denoise && sepcolor) {
                if (depth > 1) segment.color = segment.color * materialAlbedo(sc, material, intersection.uv);
            } else {
                segment.color = segment.color * materialAlbedo(sc, material, intersection.uv);
            }
            if (trace_shadowray && materialIsDiffuse) {
                const int lightIdx = 0;
                const svgf_geom &light = sc.geoms[lightIdx];
                Ray shadowRay; float shadowRayExpectDist = 0.0f;
                computeShadowRay(shadowRay, intersectionPos + 1e-4f * intersectionNormal, light, P.lightradius, shadowRayExpectDist, seed);
                svgf_intersection sh;   // uninitialised in the reference; only read after a hit on geoms[0]
                memset(&sh, 0, sizeof(sh)); sh.geomId = -2;
                computeIntersection(sc, shadowRay, sh);
                if (sh.geomId == lightIdx) {
                    const svgf_material &sm = sc.materials[sh.materialId];
                    if (sm.emittance > 0.0f) {
                        float diffuse = gmax(0.0f, dot(shadowRay.direction, intersectionNormal));
                        float shadowIntensity = P.sintensity / powf(shadowRayExpectDist, 2.0f);
                        accumulatedColor = accumulatedColor + segment.color * sm.emittance * ld(sm.color) * shadowIntensity * diffuse;
                    }
                }
            }
            if (depth < P.tracedepth) {
                scatterRay(segment, intersectionPos, intersectionNormal, material, seed);
                hit = computeIntersection(sc, segment.ray, intersection);
            }
        }
    }
    float *img = &S.image[(size_t)segment.pixelIndex * 3];
    if (denoise) st(img, accumulatedColor);
    else st(img, ld(img) * (float)frame / (float)(frame + 1) + accumulatedColor / (float)(frame + 1));
}


// ---------------------------------------------------------------------------------------------------
// denoise.cu

// denoise.cu:77-170, one pixel. `variance` is what the pixel READS, `variance_w` what it WRITES
// (same array for the in-place reference behaviour, a second array for the Jacobi variant).
inline void atrousPixel(const float *colorin, float *colorout, const float *variance, float *variance_w,
                        const svgf_gbuffer_texel *gBuffer, int resx, int resy, int x, int y, int level, bool is_last,
                        float sigma_c, float sigma_n, float sigma_x, bool blur_variance, bool addcolor) {
    static const float h[25] = {1.0 / 256.0, 1.0 / 64.0, 3.0 / 128.0, 1.0 / 64.0, 1.0 / 256.0,
                                1.0 / 64.0, 1.0 / 16.0, 3.0 / 32.0, 1.0 / 16.0, 1.0 / 64.0,
                                3.0 / 128.0, 3.0 / 32.0, 9.0 / 64.0, 3.0 / 32.0, 3.0 / 128.0,
                                1.0 / 64.0, 1.0 / 16.0, 3.0 / 32.0, 1.0 / 16.0, 1.0 / 64.0,
                                1.0 / 256.0, 1.0 / 64.0, 3.0 / 128.0, 1.0 / 64.0, 1.0 / 256.0};
    static const float gaussian[9] = {1.0 / 16.0, 1.0 / 8.0, 1.0 / 16.0, 1.0 / 8.0, 1.0 / 4.0, 1.0 / 8.0, 1.0 / 16.0, 1.0 / 8.0, 1.0 / 16.0};
    const int p = x + y * resx;
    const int step = 1 << level;
    float var;
    if (blur_variance) {
        float sum = 0.0f, sumw = 0.0f;
        static const int gx[9] = {-1, 0, 1, -1, 0, 1, -1, 0, 1}, gy[9] = {-1, -1, -1, 0, 0, 0, 1, 1, 1};
        for (int s = 0; s < 9; s++) {
            int lx = x + gx[s], ly = y + gy[s];
            if (lx >= 0 && ly >= 0 && lx < resx && ly < resy) {
                sum += gaussian[s] * variance[lx + ly * resx];
                sumw += gaussian[s];
            }
        }
        var = fmaxf(sum / sumw, 0.0f);
    } else {
        var = fmaxf(variance[p], 0.0f);
    }
    float lp = 0.2126 * colorin[p * 3] + 0.7152 * colorin[p * 3 + 1] + 0.0722 * colorin[p * 3 + 2];
    V3 pp = ld(gBuffer[p].position), np = ld(gBuffer[p].normal);
    V3 color_sum = mk(0, 0, 0);
    float variance_sum = 0.0f, weights_sum = 0, weights_squared_sum = 0;
    for (int i = -2; i <= 2; i++) {
        for (int j = -2; j <= 2; j++) {
            int xq = x + step * i, yq = y + step * j;
            if (xq >= 0 && xq < resx && yq >= 0 && yq < resy) {
                int q = xq + yq * resx;
                float lq = 0.2126 * colorin[q * 3] + 0.7152 * colorin[q * 3 + 1] + 0.0722 * colorin[q * 3 + 2];
                V3 pq = ld(gBuffer[q].position), nq = ld(gBuffer[q].normal);
                float wl = expf(-gabs(lq - lp) / (sqrtf(var) * sigma_c + 1e-6));
                float wn = fminf(1.0f, expf(-length(nq - np) / (sigma_n + 1e-6)));
                float wx = fminf(1.0f, expf(-length(pq - pp) / (sigma_x + 1e-6)));
                int k = (2 + i) + (2 + j) * 5;
                float weight = h[k] * wl * wn * wx;
                weights_sum += weight;
                weights_squared_sum += weight * weight;
                color_sum = color_sum + (ld(&colorin[q * 3]) * weight);
                variance_sum += (variance[q] * weight * weight);
            }
        }
    }
    V3 out;
    if (weights_sum > 10e-6) {
        out = color_sum / weights_sum;
        variance_w[p] = variance_sum / weights_squared_sum;
    } else {
        out = ld(&colorin[p * 3]);
    }
    if (is_last && addcolor) out = out * (ld(gBuffer[p].albedo) * ld(gBuffer[p].ialbedo));
    st(&colorout[p * 3], out);
}

void atrousLevel(const float *colorin, float *colorout, float *variance, float *variance_tmp,
                 const svgf_gbuffer_texel *g, int W, int H, int level, bool is_last, float sc, float sn, float sx,
                 bool blur, bool addcolor, int variance_mode, int threads) {
    if (variance_mode == ORC_VAR_INPLACE_SEQ) {
        // launch order of oracle/ref/cuda_emu: 8x8 blocks (denoise.cu:354-357) row-major, threads row-major
        const int gx = (W + 7) / 8, gy = (H + 7) / 8;
        for (int by = 0; by < gy; by++) for (int bx = 0; bx < gx; bx++)
            for (int ty = 0; ty < 8; ty++) for (int tx = 0; tx < 8; tx++) {
                int x = bx * 8 + tx, y = by * 8 + ty;
                if (x < W && y < H) atrousPixel(colorin, colorout, variance, variance, g, W, H, x, y, level, is_last, sc, sn, sx, blur, addcolor);
            }
    } else {
        memcpy(variance_tmp, variance, sizeof(float) * (size_t)W * H);
        (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
                atrousPixel(colorin, colorout, variance, variance_tmp, g, W, H, x, y, level, is_last, sc, sn, sx, blur, addcolor);
        memcpy(variance, variance_tmp, sizeof(float) * (size_t)W * H);
    }
}

// denoise.cu:172-182. Coordinates arrive as glm::vec2, indices are computed in float.
inline bool isReprjValid(int resx, int resy, float cx, float cy, float px, float py,
                         const svgf_gbuffer_texel *cur, const svgf_gbuffer_texel *prev) {
    if (px < 0 || px >= resx || py < 0 || py >= resy) return false;
    int p = (int)(cx + cy * resx);
    int q = (int)(px + py * resx);
    if (prev[q].geomId == -1 || prev[q].geomId != cur[p].geomId) return false;
    if (length(ld(cur[p].normal) - ld(prev[q].normal)) > 1e-1f) return false;
    return true;
}
// float -> int as the reference's glm::ivec2(floorx, floory) does it. Out-of-range values are undefined in
// C++ (x86 and the GPU disagree); they only occur for positions in the camera plane and every tap is then
// rejected by the bounds test either way, so saturate to keep the oracle well defined.
inline int f2i(float v) {
    if (!(v > -2147483000.0f)) return -2147483647 - 1;
    if (!(v < 2147483000.0f)) return 2147483647;
    return (int)v;
}

// denoise.cu:185-317, one pixel
void backProjectPixel(orc_state &S, const float *current_color, const svgf_gbuffer_texel *current_gbuffer,
                      int resx, int resy, int x, int y, float color_alpha_min, float moment_alpha_min) {
    const float *vm = S.view_matrix_prev;
    const svgf_gbuffer_texel *prev_gbuffer = S.gbuffer_prev.data();
    const int *history_length = S.history_length.data();
    const float *color_history = S.color_history.data(), *moment_history = S.moment_history.data();
    const int p = x + y * resx;
    int N = history_length[p];
    V3 sample = ld(&current_color[p * 3]);
    float luminance = 0.2126 * sample.x + 0.7152 * sample.y + 0.0722 * sample.z;
    if (N > 0 && current_gbuffer[p].geomId != -1) {
        float pos4[4] = {current_gbuffer[p].position[0], current_gbuffer[p].position[1], current_gbuffer[p].position[2], 1.0f};
        float vs[4];
        mulM4V4(vm, pos4, vs);
        float clipx = vs[0] / vs[2];
        float clipy = vs[1] / vs[2];
        float ndcx = -clipx * 0.5f + 0.5f;
        float ndcy = -clipy * 0.5f + 0.5f;
        float prevx = ndcx * resx - 0.5f;
        float prevy = ndcy * resy - 0.5f;
        bool v[4];
        float floorx = floorf(prevx), floory = floorf(prevy);
        float fracx = prevx - floorx, fracy = prevy - floory;
        bool valid = (floorx >= 0 && floory >= 0 && floorx < resx && floory < resy);
        static const int ox[4] = {0, 1, 0, 1}, oy[4] = {0, 0, 1, 1};
        const int ifx = f2i(floorx), ify = f2i(floory);
        for (int s = 0; s < 4; s++) {
            int lx = (int)((unsigned)ifx + (unsigned)ox[s]), ly = (int)((unsigned)ify + (unsigned)oy[s]);
            v[s] = isReprjValid(resx, resy, (float)x, (float)y, (float)lx, (float)ly, current_gbuffer, prev_gbuffer);
            valid = valid && v[s];
        }
        V3 prevColor = mk(0, 0, 0);
        float prevMoments[2] = {0, 0};
        float prevHistoryLength = 0.0f;
        if (valid) {
            float sumw = 0.0f;
            float w[4] = {(1 - fracx) * (1 - fracy), fracx * (1 - fracy), (1 - fracx) * fracy, fracx * fracy};
            for (int s = 0; s < 4; s++) {
                int lx = ifx + ox[s], ly = ify + oy[s];
                int locq = lx + ly * resx;
                if (v[s]) {
                    prevColor = prevColor + w[s] * ld(&color_history[locq * 3]);
                    prevMoments[0] += w[s] * moment_history[locq * 2];
                    prevMoments[1] += w[s] * moment_history[locq * 2 + 1];
                    prevHistoryLength += w[s] * (float)history_length[locq];
                    sumw += w[s];
                }
            }
            if (sumw >= 0.01) {
                prevColor = prevColor / sumw;
                prevMoments[0] /= sumw; prevMoments[1] /= sumw;
                prevHistoryLength /= sumw;
                valid = true;
            }
        }
        if (!valid) {
            float cnt = 0.0f;
            for (int yy = -1; yy <= 1; yy++) {
                for (int xx = -1; xx <= 1; xx++) {
                    float lx = floorx + (float)xx, ly = floory + (float)yy;
                    if (isReprjValid(resx, resy, (float)x, (float)y, lx, ly, current_gbuffer, prev_gbuffer)) {
                        int q = (int)(lx + resx * ly);
                        prevColor = prevColor + ld(&color_history[q * 3]);
                        prevMoments[0] += moment_history[q * 2];
                        prevMoments[1] += moment_history[q * 2 + 1];
                        prevHistoryLength += history_length[q];
                        cnt += 1.0f;
                    }
                }
            }
            if (cnt > 0.0f) {
                prevColor = prevColor / cnt;
                prevMoments[0] /= cnt; prevMoments[1] /= cnt;
                prevHistoryLength /= cnt;
                valid = true;
            }
        }
        if (valid) {
            float color_alpha = fmaxf(1.0f / (float)(N + 1), color_alpha_min);
            float moment_alpha = fmaxf(1.0f / (float)(N + 1), moment_alpha_min);
            S.history_length_update[p] = (int)prevHistoryLength + 1;
            st(&S.color_acc[p * 3], ld(&current_color[p * 3]) * color_alpha + prevColor * (1.0f - color_alpha));
            float first_moment = moment_alpha * prevMoments[0] + (1.0f - moment_alpha) * luminance;
            float second_moment = moment_alpha * prevMoments[1] + (1.0f - moment_alpha) * luminance * luminance;
            S.moment_acc[p * 2] = first_moment; S.moment_acc[p * 2 + 1] = second_moment;
            float variance = second_moment - first_moment * first_moment;
            S.variance[p] = variance > 0.0f ? variance : 0.0f;
            return;
        }
    }
    S.history_length_update[p] = 1;
    st(&S.color_acc[p * 3], ld(&current_color[p * 3]));
    S.moment_acc[p * 2] = luminance; S.moment_acc[p * 2 + 1] = luminance * luminance;
    S.variance[p] = 100.0f;
}

// glm::inverse(mat4), external/include/glm/detail/type_mat4x4.inl:37-92 (cofactor expansion, float)
void inverseM4(const float *m, float *out) {
#define M(c, r) m[(c) * 4 + (r)]
    float Coef00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3);
    float Coef02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
    float Coef03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
    float Coef04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
    float Coef06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3);
    float Coef07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
    float Coef08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
    float Coef10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
    float Coef11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
    float Coef12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
    float Coef14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3);
    float Coef15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
    float Coef16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2);
    float Coef18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
    float Coef19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
    float Coef20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
    float Coef22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1);
    float Coef23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
    const float Fac0[4] = {Coef00, Coef00, Coef02, Coef03}, Fac1[4] = {Coef04, Coef04, Coef06, Coef07};
    const float Fac2[4] = {Coef08, Coef08, Coef10, Coef11}, Fac3[4] = {Coef12, Coef12, Coef14, Coef15};
    const float Fac4[4] = {Coef16, Coef16, Coef18, Coef19}, Fac5[4] = {Coef20, Coef20, Coef22, Coef23};
    const float Vec0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, Vec1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
    const float Vec2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, Vec3[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
    static const float SignA[4] = {+1, -1, +1, -1}, SignB[4] = {-1, +1, -1, +1};
    float Inv[4][4];
    for (int i = 0; i < 4; i++) {
        Inv[0][i] = ((Vec1[i] * Fac0[i] - Vec2[i] * Fac1[i]) + Vec3[i] * Fac2[i]) * SignA[i];
        Inv[1][i] = ((Vec0[i] * Fac0[i] - Vec2[i] * Fac3[i]) + Vec3[i] * Fac4[i]) * SignB[i];
        Inv[2][i] = ((Vec0[i] * Fac1[i] - Vec1[i] * Fac3[i]) + Vec3[i] * Fac5[i]) * SignA[i];
        Inv[3][i] = ((Vec0[i] * Fac2[i] - Vec1[i] * Fac4[i]) + Vec2[i] * Fac5[i]) * SignB[i];
    }
    const float Dot0[4] = {M(0, 0) * Inv[0][0], M(0, 1) * Inv[1][0], M(0, 2) * Inv[2][0], M(0, 3) * Inv[3][0]};
    float Dot1 = (Dot0[0] + Dot0[1]) + (Dot0[2] + Dot0[3]);
    float OneOverDeterminant = 1.0f / Dot1;
    for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) out[c * 4 + r] = Inv[c][r] * OneOverDeterminant;
#undef M
}

// denoise.cu:349-402
void denoiseFrame(orc_state &S, float *output, const float *input, const svgf_gbuffer_texel *gbuffer,
                  const svgf_camera &cam, const svgf_params &P, int variance_mode, int threads) {
    const int W = cam.resolution[0], H = cam.resolution[1];
    const size_t px = (size_t)W * H;
    float color_alpha = P.temporal_enable ? P.color_alpha : 1.0f;
    float moment_alpha = P.temporal_enable ? P.moment_alpha : 1.0f;
    if (P.temporal_enable) {
#pragma omp parallel for schedule(static) num_threads(threads)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) backProjectPixel(S, input, gbuffer, W, H, x, y, color_alpha, moment_alpha);
        memcpy(S.color_history.data(), S.color_acc.data(), px * 12);
    } else {
        for (size_t i = 0; i < px; i++) S.variance[i] = 10.0f;      // EstimateVariance, denoise.cu:320-329
        memcpy(S.color_history.data(), input, px * 12);
    }
    if (P.right_view_option == 1) {         // DebugView<int>, denoise.cu:331-340, 374
        for (size_t i = 0; i < px; i++) { float v = (float)S.history_length[i] / 100.0f; output[i * 3] = output[i * 3 + 1] = output[i * 3 + 2] = v; }
    } else if (P.right_view_option == 2) {
        for (size_t i = 0; i < px; i++) { float v = (float)S.variance[i] / 0.1f; output[i * 3] = output[i * 3 + 1] = output[i * 3 + 2] = v; }
    } else if (P.atrous_nlevel == 0 || !P.spatial_enable) {
        memcpy(output, S.color_history.data(), px * 12);
    } else {
        for (int level = 1; level <= P.atrous_nlevel; level++) {
            float *src = (level == 1) ? S.color_history.data() : S.temp[level % 2].data();
            float *dst = (level == P.atrous_nlevel) ? output : S.temp[(level + 1) % 2].data();
            atrousLevel(src, dst, S.variance.data(), S.variance_tmp.data(), gbuffer, W, H, level, level == P.atrous_nlevel,
                        P.sigmal, P.sigman, P.sigmax, P.blurvariance, (P.sepcolor && P.addcolor), variance_mode, threads);
            if (level == P.history_level) memcpy(S.color_history.data(), dst, px * 12);
        }
    }
    memcpy(S.gbuffer_prev.data(), gbuffer, px * sizeof(svgf_gbuffer_texel));
    memcpy(S.moment_history.data(), S.moment_acc.data(), px * 8);
    memcpy(S.history_length.data(), S.history_length_update.data(), px * 4);
    orc_view_matrix(&cam, S.view_matrix_prev);
}

// pathtrace.cu:46-78
void packPBO(orc_state &S, int W, int H) {
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        const int index = x + y * W;
        for (int side = 0; side < 2; side++) {
            const float *pix = side == 0 ? &S.image[index * 3] : &S.denoised[index * 3];
            unsigned char *o = &S.pbo[((size_t)x + (size_t)y * W * 2 + (side ? W : 0)) * 4];
            for (int c = 0; c < 3; c++) {
                int v = (int)(pix[c] * 255.0);
                o[c] = (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
            }
            o[3] = 0;
        }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// C API

extern "C" {

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

orc_scene *orc_scene_load(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    struct { char magic[8]; int n_geoms, n_materials, n_tris, n_bvh, n_boxes, n_textures; float fovy; int reserved; } h;
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "SVGFSCN1", 8)) { fclose(f); return NULL; }
    orc_scene *s = new orc_scene();
    bool ok = fread(&s->cam, sizeof(svgf_camera), 1, f) == 1;
    s->geoms.resize(h.n_geoms); s->materials.resize(h.n_materials); s->tris.resize(h.n_tris); s->bvh.resize(h.n_bvh);
    if (h.n_geoms) ok &= fread(s->geoms.data(), sizeof(svgf_geom), h.n_geoms, f) == (size_t)h.n_geoms;
    if (h.n_materials) ok &= fread(s->materials.data(), sizeof(svgf_material), h.n_materials, f) == (size_t)h.n_materials;
    if (h.n_tris) ok &= fread(s->tris.data(), sizeof(svgf_triangle), h.n_tris, f) == (size_t)h.n_tris;
    if (h.n_bvh) ok &= fread(s->bvh.data(), sizeof(svgf_bvh_node), h.n_bvh, f) == (size_t)h.n_bvh;
    if (h.n_boxes) ok &= fseek(f, 24L * h.n_boxes, SEEK_CUR) == 0;   // world bounds: unused by the hot path
    for (int i = 0; i < h.n_textures && ok; i++) {
        int whc[3];
        ok &= fread(whc, sizeof(int), 3, f) == 3;
        orc_scene::Tex t; t.w = whc[0]; t.h = whc[1]; t.c = whc[2];
        t.px.resize((size_t)t.w * t.h * t.c);
        ok &= fread(t.px.data(), 1, t.px.size(), f) == t.px.size();
        s->textures.push_back(t);
    }
    fclose(f);
    s->fovy = h.fovy; s->n_boxes = h.n_boxes;
    if (!ok) { delete s; return NULL; }
    return s;
}
void orc_scene_free(orc_scene *s) { delete s; }
int orc_scene_counts(const orc_scene *s, int *o) {
    o[0] = (int)s->geoms.size(); o[1] = (int)s->materials.size(); o[2] = (int)s->tris.size();
    o[3] = (int)s->bvh.size(); o[4] = s->n_boxes; o[5] = (int)s->textures.size();
    return 0;
}
int orc_scene_camera(const orc_scene *s, svgf_camera *cam, float *fovy) { *cam = s->cam; *fovy = s->fovy; return 0; }
int orc_scene_desc(const orc_scene *s, svgf_scene_desc *d, svgf_texture_desc *tex, int max_tex) {
    memset(d, 0, sizeof(*d));
    d->geoms = s->geoms.data(); d->n_geoms = (int)s->geoms.size();
    d->materials = s->materials.data(); d->n_materials = (int)s->materials.size();
    d->triangles = s->tris.data(); d->n_triangles = (int)s->tris.size();
    d->bvh_nodes = s->bvh.data(); d->n_bvh_nodes = (int)s->bvh.size();
    int nt = (int)s->textures.size();
    if (nt > max_tex) return -1;
    for (int i = 0; i < nt; i++) { tex[i].width = s->textures[i].w; tex[i].height = s->textures[i].h; tex[i].components = s->textures[i].c; tex[i].pixels = s->textures[i].px.data(); }
    d->textures = tex; d->n_textures = nt;
    return 0;
}

orc_state *orc_create(const orc_scene *sc, int W, int H) {
    orc_state *S = new orc_state();
    S->scene = sc; S->W = W; S->H = H;
    memset(S->view_matrix_prev, 0, sizeof(S->view_matrix_prev));   // glm::mat4() = identity (type_mat4x4.inl:99-107)
    S->view_matrix_prev[0] = S->view_matrix_prev[5] = S->view_matrix_prev[10] = S->view_matrix_prev[15] = 1.0f;
    orc_reset(S);
    return S;
}
void orc_destroy(orc_state *S) { delete S; }

// pathtraceInit (pathtrace.cu:103-158) + denoiseInit (denoise.cu:31-61). Buffers the reference leaves
// uninitialised (color_history, gbuffer_prev, temp[], gbuffer, color_acc) are zero here.
void orc_reset(orc_state *S) {
    const size_t px = (size_t)S->W * S->H;
    S->image.assign(px * 3, 0.0f); S->denoised.assign(px * 3, 0.0f);
    svgf_gbuffer_texel zt; memset(&zt, 0, sizeof(zt));
    svgf_intersection zi; memset(&zi, 0, sizeof(zi));
    S->gbuffer.assign(px, zt); S->gbuffer_prev.assign(px, zt); S->intersections.assign(px, zi);
    S->pbo.assign(px * 8, 0);
    S->temp[0].assign(px * 3, 0.0f); S->temp[1].assign(px * 3, 0.0f);
    S->color_history.assign(px * 3, 0.0f); S->color_acc.assign(px * 3, 0.0f);
    S->moment_history.assign(px * 2, 0.0f); S->moment_acc.assign(px * 2, 0.0f);
    S->variance.assign(px, 0.0f); S->variance_tmp.assign(px, 0.0f);
    S->history_length.assign(px, 0); S->history_length_update.assign(px, 0);
}

int orc_frame(orc_state *S, const svgf_camera *cam, const svgf_params *P, int frame, int variance_mode, int threads) {
    const int W = cam->resolution[0], H = cam->resolution[1];
    if (W != S->W || H != S->H) return -1;
    if (threads <= 0) threads = orc_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) tracePixel(*S, *cam, *P, frame, x, y);
    if (P->denoise_enable) denoiseFrame(*S, S->denoised.data(), S->image.data(), S->gbuffer.data(), *cam, *P, variance_mode, threads);
    else memcpy(S->denoised.data(), S->image.data(), (size_t)W * H * 12);
    packPBO(*S, W, H);
    return 0;
}

// denoise(output, input, gbuffer) (src/denoise.h:8, denoise.cu:349-402) on caller buffers, against the state's histories: the
// public entry point on its own, without the path tracer in front of it.
int orc_denoise(orc_state *S, float *output, const float *input, const svgf_gbuffer_texel *gbuffer, const svgf_camera *cam,
                const svgf_params *P, int variance_mode, int threads) {
    if (cam->resolution[0] != S->W || cam->resolution[1] != S->H) return -1;
    if (threads <= 0) threads = orc_max_threads();
    denoiseFrame(*S, output, input, gbuffer, *cam, *P, variance_mode, threads);
    return 0;
}

int orc_fetch(orc_state *S, const char *name, void *host, size_t bytes) {
    const void *src = NULL; size_t need = 0;
#define F(n, vec) if (!strcmp(name, n)) { src = S->vec.data(); need = S->vec.size() * sizeof(S->vec[0]); }
    F("image", image) F("denoised", denoised) F("gbuffer", gbuffer) F("intersections", intersections) F("pbo", pbo)
    F("variance", variance) F("color_acc", color_acc) F("color_history", color_history) F("moment_acc", moment_acc)
    F("moment_history", moment_history) F("history_length", history_length) F("history_length_update", history_length_update)
    F("gbuffer_prev", gbuffer_prev) F("temp0", temp[0]) F("temp1", temp[1]) F("host_image", denoised)
#undef F
    if (!strcmp(name, "view_matrix_prev")) { src = S->view_matrix_prev; need = 64; }
    if (!src) return -3;
    if (bytes != need) return -2;
    memcpy(host, src, need);
    return 0;
}

int orc_host_intersect(const orc_scene *sc, const float *origin, const float *dir, float *t, float *normal, float *uv,
                       int *geomId, int *materialId) {
    Ray r; r.origin = ld(origin); r.direction = ld(dir);
    svgf_intersection is; memset(&is, 0, sizeof(is));
    bool hit = computeIntersection(*sc, r, is);
    *t = is.t; normal[0] = is.surfaceNormal[0]; normal[1] = is.surfaceNormal[1]; normal[2] = is.surfaceNormal[2];
    uv[0] = is.uv[0]; uv[1] = is.uv[1]; *geomId = is.geomId; *materialId = is.materialId;
    return hit ? 1 : 0;
}

int orc_atrous_level(float *color_out, float *variance_out, const float *color_in, const float *variance_in,
                     const svgf_gbuffer_texel *gbuffer, int W, int H, int level, int is_last,
                     float sigma_c, float sigma_n, float sigma_x, int blur_variance, int addcolor,
                     int variance_mode, int threads) {
    if (threads <= 0) threads = orc_max_threads();
    const size_t px = (size_t)W * H;
    std::vector<float> var(variance_in, variance_in + px), tmp(px);
    atrousLevel(color_in, color_out, var.data(), tmp.data(), gbuffer, W, H, level, is_last != 0, sigma_c, sigma_n, sigma_x,
                blur_variance != 0, addcolor != 0, variance_mode, threads);
    memcpy(variance_out, var.data(), px * 4);
    return 0;
}

// GetViewMatrix, denoise.cu:342-347
void orc_view_matrix(const svgf_camera *cam, float *out16) {
    float m[16] = {cam->right[0], cam->right[1], cam->right[2], 0.f, cam->up[0], cam->up[1], cam->up[2], 0.f,
                   cam->view[0], cam->view[1], cam->view[2], 0.f, cam->position[0], cam->position[1], cam->position[2], 1.f};
    inverseM4(m, out16);
}

// scene.cpp:159-168 + main.cpp:77-101 (resetCamera)
void orc_camera_init(svgf_camera *cam, svgf_camera_rig *rig, const float eye[3], const float lookat[3],
                     const float up[3], float fovy, int W, int H) {
    memset(cam, 0, sizeof(*cam));
    cam->resolution[0] = W; cam->resolution[1] = H;
    float yscaled = tanf(fovy * (PI_F / 180));
    float xscaled = (yscaled * W) / H;
    float fovx = (atanf(xscaled) * 180) / PI_F;
    cam->fov[0] = fovx; cam->fov[1] = fovy;
    cam->pixelLength[0] = 2 * xscaled / (float)W; cam->pixelLength[1] = 2 * yscaled / (float)H;
    st(cam->position, ld(eye)); st(cam->lookAt, ld(lookat)); st(cam->up, ld(up));
    V3 view = normalize(ld(lookat) - ld(eye));
    st(cam->view, view);
    memset(rig, 0, sizeof(*rig));
    rig->fovy = fovy;
    V3 viewXZ = mk(view.x, 0.0f, view.z), viewZY = mk(0.0f, view.y, view.z);
    rig->phi = acosf(dot(normalize(viewXZ), mk(0, 0, -1)));
    rig->theta = acosf(dot(normalize(viewZY), mk(0, 1, 0)));
    rig->zoom = length(ld(eye) - ld(lookat));
}

// main.cpp:156-190
void orc_camera_step(svgf_camera *cam, svgf_camera_rig *rig, int automate, const float sp[5]) {
    if (automate) {
        rig->tx += sp[0]; rig->ty += sp[1]; rig->tz += sp[2]; rig->ttheta += sp[3]; rig->tphi += sp[4];
        cam->lookAt[0] = 0.0f + 2.0f * sinf(rig->tx);
        cam->lookAt[1] = 5.0f + 1.0f * sinf(rig->ty);
        cam->lookAt[2] = 0.0f + 1.5f * sinf(rig->tz);
        rig->theta = PI_F * 0.5f + PI_F / 18 * sinf(rig->ttheta);
        rig->phi = PI_F * 0.0f + PI_F / 12 * sinf(rig->tphi);
    }
    V3 cp;
    cp.x = rig->zoom * sinf(rig->phi) * sinf(rig->theta);
    cp.y = rig->zoom * cosf(rig->theta);
    cp.z = rig->zoom * cosf(rig->phi) * sinf(rig->theta);
    V3 v = -normalize(cp);
    V3 r = cross(v, mk(0, 1, 0));
    st(cam->view, v); st(cam->up, cross(r, v)); st(cam->right, r);
    st(cam->position, cp + ld(cam->lookAt));
}

}  // extern "C"
