// pathtrace.cu -- 1-spp path-trace feeder, sm_100a.
//
// Replaces generateRayFromCamera + rt of the reference (src/pathtrace.cu:187-208, 300-401) and the
// device libraries they call (src/intersections.h, src/interactions.h, sceneStructs.h:157-221,
// boundingbox.h:62-79). Per-pixel results are a pure function of (pixel index, frame, depth) because the
// reference re-seeds its RNG as initRand(idx, frame + depth, 16) at every bounce (pathtrace.cu:328), so any
// execution order reproduces the reference. The arithmetic follows the reference's (and glm 0.9.6.3's)
// expression order so results agree to rounding; what is re-designed is everything around it:
//   * scene tables live in shared memory (geoms, materials) or compact 16-byte records (BVH nodes 32 B instead
//     of 40 B AoS, triangles as {v0,e1,e2} + a cold shading record instead of one 136 B struct);
//   * the global BVH is traversed ONCE per query instead of once per MESH geom (pathtrace.cu:244-255): every
//     mesh geom receives the same closest triangle and only the owner of its id accepts it, so one traversal
//     and a range test per geom give the same answer;
//   * normal/uv interpolation is deferred to the winning triangle;
//   * the G-buffer is written as float4 SoA planes, the persistent per-pixel intersection record is reduced to
//     the 24 bytes that can be observed (stale normal/material/uv on a primary miss, pathtrace.cu:316-322).
#include "svgf_internal.h"
#include "halo_sync.cuh"
#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <cstring>

namespace {

#define PI_F 3.1415926535897932384626422832795028841971f
#define TWO_PI_F 6.2831853071795864769252867665590057683943f
#define SQRT_OF_ONE_THIRD_F 0.5773502691896257645091487805019574556476f
#define COLORDIVIDOR_F 0.003921568627f

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 mk(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ F3 operator-(F3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ F3 operator*(F3 a, F3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ F3 operator*(F3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ F3 operator*(float s, F3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ F3 operator/(F3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ float dot(F3 a, F3 b) { F3 t = a * b; return t.x + t.y + t.z; }
__device__ __forceinline__ F3 cross(F3 x, F3 y) { return mk(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
__device__ __forceinline__ float length(F3 v) { return sqrtf(dot(v, v)); }
__device__ __forceinline__ F3 normalize(F3 v) { return v * (1.0f / sqrtf(dot(v, v))); }
__device__ __forceinline__ float gmin(float x, float y) { return x < y ? x : y; }
__device__ __forceinline__ float gmax(float x, float y) { return x > y ? x : y; }
__device__ __forceinline__ float gabs(float x) { return x >= 0.0f ? x : -x; }

// glm mat4 * vec4 -> xyz, (m0 v0 + m1 v1) + (m2 v2 + m3 v3)   (intersections.h:36-38)
#ifndef SVGF_RT_MV_PACKED       // A/B (tools/build_rt_ab.sh): rows x and y of the product as one register pair (FMUL2/FFMA2/FADD2), the same
#define SVGF_RT_MV_PACKED 0     // operations with the same roundings as the scalar form the compiler makes of the loop below
#endif
__device__ __forceinline__ F3 multiplyMV(const float *m, F3 v, float w) {
#if SVGF_RT_MV_PACKED
    const float2 c0 = make_float2(m[0], m[1]), c1 = make_float2(m[4], m[5]), c2 = make_float2(m[8], m[9]), c3 = make_float2(m[12], m[13]);
    const float2 a0 = __ffma2_rn(c0, make_float2(v.x, v.x), __fmul2_rn(c1, make_float2(v.y, v.y)));
    const float2 a1 = __ffma2_rn(c2, make_float2(v.z, v.z), __fmul2_rn(c3, make_float2(w, w)));
    const float2 xy = __fadd2_rn(a0, a1);
    const float z0 = __fmaf_rn(m[2], v.x, __fmul_rn(m[6], v.y)), z1 = __fmaf_rn(m[10], v.z, __fmul_rn(m[14], w));
    return mk(xy.x, xy.y, __fadd_rn(z0, z1));
#else
    float o[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        float add0 = m[0 + r] * v.x + m[4 + r] * v.y;
        float add1 = m[8 + r] * v.z + m[12 + r] * w;
        o[r] = add0 + add1;
    }
    return mk(o[0], o[1], o[2]);
#endif
}

struct Ray { F3 origin, direction; };

// interactions.h:10-30
#ifndef SVGF_RT_TEA_UNROLL      // A/B (tools/build_rt_ab.sh): rounds per trip of the hash loop
#define SVGF_RT_TEA_UNROLL 1
#endif
constexpr int kTeaUnroll = SVGF_RT_TEA_UNROLL;
__device__ __forceinline__ unsigned int initRand(unsigned int val0, unsigned int val1) {
    unsigned int v0 = val0, v1 = val1, s0 = 0;
#pragma unroll kTeaUnroll
    for (unsigned int n = 0; n < 16; n++) {
        s0 += 0x9e3779b9;
        v0 += ((v1 << 4) + 0xa341316c) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4);
        v1 += ((v0 << 4) + 0xad90777d) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761e);
    }
    return v0;
}
__device__ __forceinline__ float nextRand(unsigned int &s) {
    s = (1664525u * s + 1013904223u);
    return float(s & 0x00FFFFFF) / float(0x01000000);
}

__device__ __forceinline__ F3 getPointOnRay(const Ray &r, float t) {   // intersections.h:29-31
    return r.origin + (t - .0001f) * normalize(r.direction);
}

// intersections.h:50-92, the slab part: object-space ray q against the unit cube. Yields the object-space t of the hit
// and the face (axis, sign) the deferred normal is made from; the shared head (ray into object space) and tail (hit point
// back to world space, distance) live in computeIntersection, common to cubes and spheres.
__device__ __forceinline__ bool box_slabs(const Ray &q, float &tobj, int &axis, float &sign) {
    float tmin = -1e38f, tmax = 1e38f;
    int tmin_axis = -1, tmax_axis = -1; float tmin_s = 0.f, tmax_s = 0.f;   // normal = sign on one axis, zero elsewhere
    const float qo[3] = {q.origin.x, q.origin.y, q.origin.z}, qd[3] = {q.direction.x, q.direction.y, q.direction.z};
#pragma unroll
    for (int xyz = 0; xyz < 3; ++xyz) {
        float qdxyz = qd[xyz];
        float t1 = (-0.5f - qo[xyz]) / qdxyz;
        float t2 = (+0.5f - qo[xyz]) / qdxyz;
        float ta = gmin(t1, t2);
        float tb = gmax(t1, t2);
        float s = t2 < t1 ? +1.f : -1.f;
        if (ta > 0 && ta > tmin) { tmin = ta; tmin_axis = xyz; tmin_s = s; }
        if (tb < tmax) { tmax = tb; tmax_axis = xyz; tmax_s = s; }
    }
    if (tmax >= tmin && tmax > 0) {
        if (tmin <= 0) { tmin = tmax; tmin_axis = tmax_axis; tmin_s = tmax_s; }
        tobj = tmin; axis = tmin_axis; sign = tmin_s;
        return true;
    }
    return false;
}
__device__ __forceinline__ F3 box_normal(const GeomD &box, int axis, float sign) {      // intersections.h:67-68,90
    const F3 n = mk(axis == 0 ? sign : 0.f, axis == 1 ? sign : 0.f, axis == 2 ? sign : 0.f);
    return normalize(multiplyMV(box.transform, n, 0.0f));
}

// intersections.h:104-131, the quadratic: object-space ray against the sphere of radius 0.5
__device__ __forceinline__ bool sphere_roots(const Ray &rt, float &tobj, bool &outside) {
    float vDotDirection = dot(rt.origin, rt.direction);
    float radicand = vDotDirection * vDotDirection - (dot(rt.origin, rt.origin) - 0.25f);
    if (radicand < 0) return false;
    float squareRoot = sqrtf(radicand);
    float firstTerm = -vDotDirection;
    float t1 = firstTerm + squareRoot;
    float t2 = firstTerm - squareRoot;
    if (t1 < 0 && t2 < 0) return false;
    else if (t1 > 0 && t2 > 0) { tobj = fminf(t1, t2); outside = true; }
    else { tobj = fmaxf(t1, t2); outside = false; }
    return true;
}
__device__ __forceinline__ F3 sphere_normal(const GeomD &sphere, F3 osi, bool outside) {    // intersections.h:140-143
    F3 n = normalize(multiplyMV(sphere.invTranspose, osi, 0.f));
    return outside ? n : -n;
}

// A Geom record in shared memory. In the exact-test loop every lane reads the SAME field of a DIFFERENT geom; at the record's
// natural 256-byte stride all those words share one bank (ncu: 32 % of the kernel's shared-memory wavefronts were conflicts).
// 16 bytes of padding per record (68-word stride) rotate the records by four banks each -- a 16-byte matrix row of eight
// different geoms is conflict-free -- and keep the 16-byte alignment the compiler's LDS.128 of the matrices relies on (a
// 65-word stride was measured first: it forced scalar loads and DOUBLED the kernel time).
#ifndef SVGF_RT_NO_PAD
struct alignas(16) GeomS : GeomD { int bank_pad_[4]; };
static_assert(sizeof(GeomS) == 272, "GeomS");
#else       // A/B build (tools/build_rt_ab.sh): the record at its natural stride
struct GeomS : GeomD {};
#endif

// Copies the scene tables into shared memory (all threads of the block; caller synchronises).
__device__ __forceinline__ void stage_scene(unsigned char *smem, const GeomD *g_geoms, int n_geoms, const svgf_material *g_materials,
                                            int n_materials, int tid, int nt, GeomS *&s_geoms, svgf_material *&s_mats) {
    s_geoms = reinterpret_cast<GeomS *>(smem);
    s_mats = reinterpret_cast<svgf_material *>(smem + sizeof(GeomS) * n_geoms);
    const int gw = sizeof(GeomD) / 4 * n_geoms, mw = sizeof(svgf_material) / 4 * n_materials;
    const int *src = reinterpret_cast<const int *>(g_geoms); int *dst = reinterpret_cast<int *>(s_geoms);
    for (int i = tid; i < gw; i += nt) dst[(i >> 6) * (int)(sizeof(GeomS) / 4) + (i & 63)] = src[i];
    src = reinterpret_cast<const int *>(g_materials); dst = reinterpret_cast<int *>(s_mats);
    for (int i = tid; i < mw; i += nt) dst[i] = src[i];
}
__host__ __device__ inline size_t scene_smem_bytes(int n_geoms, int n_materials) { return sizeof(GeomS) * n_geoms + sizeof(svgf_material) * n_materials; }

struct SceneView {
    // The shadow query may stop at the first occluder (see computeIntersection) when geoms[0] is a cube or a sphere and every
    // triangle belongs to a MESH geom's range (always so for scenes of the reference's loader; checked at upload).
    bool light_query_ok;
    const GeomS *geoms; int n_geoms;            // shared memory
    const svgf_material *materials;             // shared memory
    const float4 *bvh; int n_nodes;
    const float4 *tri_hot, *tri_cold;
    const TexD *textures;
};

// IntersectBVH (intersections.h:265-329) with Triangle::Intersect (sceneStructs.h:157-180) over
// glm::intersectRayTriangle (gtx/intersect.inl:37-74); same near-first order, same 64-entry stack
// (overflow drops the subtree), same "first strictly smaller t wins".
struct TriBest { float t, bx, by; int slot; };
// `t_bound`: hits at or beyond it cannot matter to the caller (it already holds a closer surface), so subtrees whose
// slab entry lies beyond it -- or beyond the closest triangle found so far -- are skipped. The reference visits them
// (no t-culling in AABBIntersect2) and then discards what it finds there: a triangle inside a node cannot be hit before
// the ray enters the node's box, so with a conservative margin the closest triangle is the same.
// `t_any` > 0: occlusion query -- return at the first triangle hit strictly in front of t_any (0 < t < t_any); which one is
// irrelevant to the caller. (The closest-hit search would reject the mesh if its CLOSEST triangle sat at exactly t == 0; that
// this differs needs a triangle hit at exactly 0 and another one before t_any on the same ray.)
// Slab test of one node record against the ray, with the caller's cull distance (boundingbox.h:62-79 has no t-culling).
#ifndef SVGF_RT_SLAB_PACKED
#define SVGF_RT_SLAB_PACKED 2
#endif
#ifndef SVGF_RT_SLAB_UNROLL     // A/B (tools/build_rt_ab.sh): unrolling of the candidate loop over the geoms
#define SVGF_RT_SLAB_UNROLL 1
#endif
constexpr int kSlabUnroll = SVGF_RT_SLAB_UNROLL;
// Entry and exit parameters of a ray against a box {lo.xyz, hi.xyz}: (bound - origin) * invdir per axis. SVGF_RT_SLAB_PACKED (A/B,
// tools/build_rt_ab.sh): the x and y axes travel as one register pair (FADD2/FMUL2, the same IEEE operations, two per issue slot).
__device__ __forceinline__ void slab_params(const float4 &lo, const float4 &hi, const Ray &ray, const F3 invdir, float &tn, float &tf) {
#if SVGF_RT_SLAB_PACKED
    const float2 no = make_float2(-ray.origin.x, -ray.origin.y), iv = make_float2(invdir.x, invdir.y);
    const float2 l = __fmul2_rn(__fadd2_rn(make_float2(lo.x, lo.y), no), iv), h = __fmul2_rn(__fadd2_rn(make_float2(hi.x, hi.y), no), iv);
    const float ax = l.x, ay = l.y, bx = h.x, by = h.y;
#else
    const float ax = (lo.x - ray.origin.x) * invdir.x, bx = (hi.x - ray.origin.x) * invdir.x;
    const float ay = (lo.y - ray.origin.y) * invdir.y, by = (hi.y - ray.origin.y) * invdir.y;
#endif
    const float az = (lo.z - ray.origin.z) * invdir.z, bz = (hi.z - ray.origin.z) * invdir.z;
    tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
}
__device__ __forceinline__ bool bvh_box_hit(const float4 &a, const float4 &b, const Ray &ray, const F3 invdir, float t_cull) {
#if SVGF_RT_SLAB_PACKED >= 2
    const float2 no = make_float2(-ray.origin.x, -ray.origin.y), iv = make_float2(invdir.x, invdir.y);
    const float2 l = __fmul2_rn(__fadd2_rn(make_float2(a.x, a.y), no), iv), h = __fmul2_rn(__fadd2_rn(make_float2(b.x, b.y), no), iv);
    const float tzMin = (a.z - ray.origin.z) * invdir.z, tzMax = (b.z - ray.origin.z) * invdir.z;
    const float tmin = gmax(gmax(gmin(l.x, h.x), gmin(l.y, h.y)), gmin(tzMin, tzMax));
    const float tmax = gmin(gmin(gmax(l.x, h.x), gmax(l.y, h.y)), gmax(tzMin, tzMax));
    return !(tmax < 0) && !(tmin > tmax) && !(tmin > t_cull * 1.0001f + 1e-4f);
#else
    const float txMin = (a.x - ray.origin.x) * invdir.x, txMax = (b.x - ray.origin.x) * invdir.x;
    const float tyMin = (a.y - ray.origin.y) * invdir.y, tyMax = (b.y - ray.origin.y) * invdir.y;
    const float tzMin = (a.z - ray.origin.z) * invdir.z, tzMax = (b.z - ray.origin.z) * invdir.z;
    const float tmin = gmax(gmax(gmin(txMin, txMax), gmin(tyMin, tyMax)), gmin(tzMin, tzMax));
    const float tmax = gmin(gmin(gmax(txMin, txMax), gmax(tyMin, tyMax)), gmax(tzMin, tzMax));
    return !(tmax < 0) && !(tmin > tmax) && !(tmin > t_cull * 1.0001f + 1e-4f);
#endif
}

#ifndef SVGF_RT_BVH_PAIRED
#define SVGF_RT_BVH_PAIRED 0
#endif
#if SVGF_RT_BVH_PAIRED
// A/B variant (tools/build_rt_ab.sh, -DSVGF_RT_BVH_PAIRED=1), MEASURED SLOWER on B200 and therefore not the default: C3 room 3.11 vs
// 2.80 ms, C5 bunny 1.52 vs 1.43, C2 0.87 vs 0.77 (profiles/r2_ab_rt_bvh_paired.txt) -- the two extra node records in flight cost
// more in the 64-register kernel than the saved round trips return. BOTH children of an interior node are fetched and tested
// together (four independent loads, one round trip to memory per level of the tree), the near one is entered with its record
// already in registers and the far one is stacked only if the ray hits its box. The default loop (below) enters the near child,
// stacks the far one untested and pays a dependent fetch for every node it pops, missed ones included. The set of nodes and triangles visited and their order are
// exactly the same -- a far child rejected here would have been rejected when popped (the cull distance only shrinks), one that
// passes is tested again when popped, with the cull distance of that moment, as before -- so the results are bit-identical.
// (The reference's 64-entry stack drops subtrees when it overflows; with fewer entries stacked that could only differ for a tree
// more than 64 levels deep.)
__device__ bool intersectBVH(const SceneView &sc, const Ray &ray, const F3 invdir, float t_bound, TriBest &best, float t_any = 0.f) {
    if (sc.n_nodes == 0) return false;
    bool hit = false;
    const int neg[3] = {ray.direction.x < 0.f, ray.direction.y < 0.f, ray.direction.z < 0.f};
    int top = 0, cur = 0;
    int stack[64];
    best.t = FLT_MAX; best.slot = -1; best.bx = best.by = 0.f;
    float4 a = __ldg(&sc.bvh[0]), b = __ldg(&sc.bvh[1]);
    bool pop = !bvh_box_hit(a, b, ray, invdir, fminf(best.t, t_bound));      // the root is tested like any node
    while (true) {
        if (pop) {      // next stacked node; it is tested (again) with the cull distance of NOW
            if (top == 0) break;
            cur = stack[--top];
            a = __ldg(&sc.bvh[2 * cur]); b = __ldg(&sc.bvh[2 * cur + 1]);
            pop = !bvh_box_hit(a, b, ray, invdir, fminf(best.t, t_bound));
            continue;
        }
        // node `cur` with record (a, b) has passed its box test
        const int meta = __float_as_int(a.w), off = __float_as_int(b.w);
        const int count = meta & 0xffff;
        if (count > 0) {
            for (int i = 0; i < count; i++) {
                const int slot = off + i;
                const float4 h0 = __ldg(&sc.tri_hot[3 * slot]), h1 = __ldg(&sc.tri_hot[3 * slot + 1]), h2 = __ldg(&sc.tri_hot[3 * slot + 2]);
                const F3 v0 = mk(h0.x, h0.y, h0.z), e1 = mk(h1.x, h1.y, h1.z), e2 = mk(h2.x, h2.y, h2.z);
                const F3 p = cross(ray.direction, e2);
                const float det = dot(e1, p);
                if (det < FLT_EPSILON) continue;
                const float f = 1.0f / det;
                const F3 s = ray.origin - v0;
                const float bx = f * dot(s, p);
                if (bx < 0.0f) continue;
                if (bx > 1.0f) continue;
                const F3 q = cross(s, e1);
                const float by = f * dot(ray.direction, q);
                if (by < 0.0f) continue;
                if (by + bx > 1.0f) continue;
                const float bz = f * dot(e2, q);
                if (!(bz >= 0.0f)) continue;
                hit = true;
                if (bz < best.t) { best.t = bz; best.bx = bx; best.by = by; best.slot = slot; }
                if (bz > 0.0f && bz < t_any) { top = 0; t_bound = -FLT_MAX; }      // (light query, compiled out by default: ends the search through data)
            }
            pop = true;
        } else if (top == 64) {
            pop = true;                 // the reference drops both children of a node it meets with a full stack
        } else {
            const int L = cur + 1, R = off;
            const float4 aL = __ldg(&sc.bvh[2 * L]), bL = __ldg(&sc.bvh[2 * L + 1]), aR = __ldg(&sc.bvh[2 * R]), bR = __ldg(&sc.bvh[2 * R + 1]);
            const float t_cull = fminf(best.t, t_bound);
            const bool hL = bvh_box_hit(aL, bL, ray, invdir, t_cull), hR = bvh_box_hit(aR, bR, ray, invdir, t_cull);
            const bool r_first = neg[meta >> 16] != 0;                     // near child: the right one for a ray running against the split axis
            const bool h_first = r_first ? hR : hL, h_second = r_first ? hL : hR;
            const int first = r_first ? R : L, second = r_first ? L : R;
            if (h_first & h_second) stack[top++] = second;
            const bool go_first = h_first;
            cur = go_first ? first : second;
            const bool take_r = go_first == r_first;                       // the record that goes with `cur`
            a = take_r ? aR : aL; b = take_r ? bR : bL;
            pop = !(h_first | h_second);
        }
    }
    return hit;
}
#else
#ifndef SVGF_RT_TRI_FLAT
#define SVGF_RT_TRI_FLAT 0
#endif
#ifndef SVGF_RT_BVH_WW
#define SVGF_RT_BVH_WW 0
#endif
// Triangle::Intersect over glm::intersectRayTriangle for the triangle in `slot`. SVGF_RT_TRI_FLAT (A/B, tools/build_rt_ab.sh): the
// same expressions without the early exits -- every lane of a leaf runs the whole test and one predicate decides -- so that lanes
// rejected at different stages do not drift apart inside the loop.
__device__ __forceinline__ void tri_test(const SceneView &sc, const Ray &ray, int slot, bool &hit, TriBest &best, float t_any, int &top,
                                         float &t_bound) {
    const float4 h0 = __ldg(&sc.tri_hot[3 * slot]), h1 = __ldg(&sc.tri_hot[3 * slot + 1]), h2 = __ldg(&sc.tri_hot[3 * slot + 2]);
    const F3 v0 = mk(h0.x, h0.y, h0.z), e1 = mk(h1.x, h1.y, h1.z), e2 = mk(h2.x, h2.y, h2.z);
    const F3 p = cross(ray.direction, e2);
    const float det = dot(e1, p);
#if SVGF_RT_TRI_FLAT
    const float f = 1.0f / det;
    const F3 s = ray.origin - v0;
    const float bx = f * dot(s, p);
    const F3 q = cross(s, e1);
    const float by = f * dot(ray.direction, q);
    const float bz = f * dot(e2, q);
    // the reference's chain of rejections, comparison for comparison (a NaN passes where it passes there)
    const bool ok = !(det < FLT_EPSILON) & !(bx < 0.0f) & !(bx > 1.0f) & !(by < 0.0f) & !(by + bx > 1.0f) & (bz >= 0.0f);
    if (ok) {
        hit = true;
        if (bz < best.t) { best.t = bz; best.bx = bx; best.by = by; best.slot = slot; }
        if (bz > 0.0f && bz < t_any) { top = 0; t_bound = -FLT_MAX; }
    }
#else
    if (det < FLT_EPSILON) return;
    const float f = 1.0f / det;
    const F3 s = ray.origin - v0;
    const float bx = f * dot(s, p);
    if (bx < 0.0f) return;
    if (bx > 1.0f) return;
    const F3 q = cross(s, e1);
    const float by = f * dot(ray.direction, q);
    if (by < 0.0f) return;
    if (by + bx > 1.0f) return;
    const float bz = f * dot(e2, q);
    if (!(bz >= 0.0f)) return;
    hit = true;
    if (bz < best.t) { best.t = bz; best.bx = bx; best.by = by; best.slot = slot; }
    if (bz > 0.0f && bz < t_any) { top = 0; t_bound = -FLT_MAX; }      // (light query, compiled out by default: ends the search through data)
#endif
}
#if SVGF_RT_BVH_WW
// A/B variant (tools/build_rt_ab.sh): "while-while" form of the same walk. The inner loop runs down to the next leaf whose box the
// ray hits, the lanes of a warp meet again at its exit, and only then are the triangles tested -- in the default loop a lane at a
// leaf runs its triangle tests while the lanes at interior nodes wait, and the other way round. Nodes and triangles are visited
// in the same order with the same cull distances, so the results are the same bits.
__device__ bool intersectBVH(const SceneView &sc, const Ray &ray, const F3 invdir, float t_bound, TriBest &best, float t_any = 0.f) {
    if (sc.n_nodes == 0) return false;
    bool hit = false;
    const int neg[3] = {ray.direction.x < 0.f, ray.direction.y < 0.f, ray.direction.z < 0.f};
    int top = 0, cur = 0;
    int stack[64];
    best.t = FLT_MAX; best.slot = -1; best.bx = best.by = 0.f;
    while (true) {
        int off = 0, count = 0;
        while (true) {
            const float4 a = __ldg(&sc.bvh[2 * cur]), b = __ldg(&sc.bvh[2 * cur + 1]);
            if (bvh_box_hit(a, b, ray, invdir, fminf(best.t, t_bound))) {
                const int meta = __float_as_int(a.w), o = __float_as_int(b.w);
                if ((meta & 0xffff) > 0) { off = o; count = meta & 0xffff; break; }
                if (top == 64) { cur = stack[--top]; continue; }
                if (neg[meta >> 16]) { stack[top++] = cur + 1; cur = o; }
                else { stack[top++] = o; cur = cur + 1; }
            } else {
                if (top == 0) break;
                cur = stack[--top];
            }
        }
        if (count == 0) break;
        for (int i = 0; i < count; i++) tri_test(sc, ray, off + i, hit, best, t_any, top, t_bound);
        if (top == 0) break;
        cur = stack[--top];
    }
    return hit;
}
#else
__device__ bool intersectBVH(const SceneView &sc, const Ray &ray, const F3 invdir, float t_bound, TriBest &best, float t_any = 0.f) {
    if (sc.n_nodes == 0) return false;
    bool hit = false;
    const int neg[3] = {ray.direction.x < 0.f, ray.direction.y < 0.f, ray.direction.z < 0.f};
    int top = 0, cur = 0;
    int stack[64];
    best.t = FLT_MAX; best.slot = -1; best.bx = best.by = 0.f;
    while (true) {
        const float4 a = __ldg(&sc.bvh[2 * cur]), b = __ldg(&sc.bvh[2 * cur + 1]);
        const bool box_hit = bvh_box_hit(a, b, ray, invdir, fminf(best.t, t_bound));
        if (box_hit) {
            const int meta = __float_as_int(a.w), off = __float_as_int(b.w);
            const int count = meta & 0xffff;
            if (count > 0) {
                for (int i = 0; i < count; i++) tri_test(sc, ray, off + i, hit, best, t_any, top, t_bound);
                if (top == 0) break;
                cur = stack[--top];
            } else {
                if (top == 64) { cur = stack[--top]; continue; }
                const int axis = meta >> 16;
                if (neg[axis]) { stack[top++] = cur + 1; cur = off; }
                else { stack[top++] = off; cur = cur + 1; }
            }
        } else {
            if (top == 0) break;
            cur = stack[--top];
        }
    }
    return hit;
}
#endif
#endif

struct Isect {      // the live part of ShadeableIntersection (sceneStructs.h:104-111)
    float t; F3 n; int materialId, geomId; float u, v;
};

// computeIntersection, pathtrace.cu:210-281: the closest t > 0 over all geoms, ties going to the lowest geom index
// (the reference loops in index order with a strict `<`). On a miss only t and geomId change (267-271).
//
// Evaluation order is re-designed for SIMT: the reference's loop makes all 32 lanes visit geom i together, and only the
// lanes whose ray can reach it do any work in the expensive exact test. Here each lane first collects the cubes and
// spheres its own ray can reach (slab test against conservative world bounds), then pops them one by one, so in every trip
// of the exact-test loops all lanes are busy -- each with a different geom. The tie rule is applied explicitly, so the
// winner is the reference's. Normals are computed for the winner only.
//
// `light_query`: the shadow ray of pathtrace.cu:358-384 only asks "is geoms[0], the light, the closest hit?" (lightIdx == 0;
// geom 0 wins every tie by index). For such a query the light is tested FIRST -- a miss answers the question -- and its
// distance then bounds everything else: candidates entered beyond it are dropped by the bounds test, and the first hit strictly
// in front of it ends the query (an any-hit search instead of a closest-hit one; no normal is computed). The answer to the
// question, and with it every output bit, is the reference's. The flag is per-lane DATA, so path and shadow queries still
// share the one call site.
// `light` < 0: path query. >= 0: light query for that geom (0 in the reference; any emissive cube/sphere under the
// "light_sampling_all" option, where a geom of LOWER index at exactly the light's distance wins the tie as in the closest-hit
// search; the measure-zero tie between the light and a triangle is not reproduced there).
__device__ bool computeIntersection(const SceneView &sc, const Ray &ray, Isect &is, int light) {
#ifdef SVGF_RT_LIGHT_QUERY      // A/B build (tools/build_rt_ab.sh). MEASURED SLOWER and therefore compiled out of the product, see below.
    const bool lq = light >= 0 && light < 32 && sc.light_query_ok;
#else
    const bool lq = false;
#endif
    float t_min = FLT_MAX;
    int hit_geom = -1;
    bool lq_miss = false;       // light query answered "no": the light is missed or something lies in front of it
    // what the winner's deferred normal needs
    int w_axis = 0; float w_sign = 0.f; F3 w_osi = mk(0, 0, 0); bool w_outside = true; int w_kind = -1;   // 1 cube, 0 sphere, 2 mesh
    // 1 / direction exactly as IntersectBVH forms it (intersections.h:276); also feeds the conservative bounds pre-test
    const F3 invdir = mk(1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z);
    bool any_mesh = false;
    // |direction| turns the slab parameter of the bounds pre-test into a distance comparable with the reference's t
    const float dlen = length(ray.direction);
    for (int base = 0; base < sc.n_geoms; base += 32) {
        unsigned cubes = 0, spheres = 0;
        float near_tn = FLT_MAX; int near_j = -1;       // candidate whose bounds the ray enters first
        const int n = min(32, sc.n_geoms - base);
#pragma unroll kSlabUnroll
        for (int j = 0; j < n; j++) {
            const GeomD &g = sc.geoms[base + j];
            if (g.type == 2) { any_mesh = true; continue; }
            // Slab test against the inflated world bounds. A NaN direction keeps every term NaN (no reject: the exact test
            // then runs as in the reference); 0 * inf only arises for a ray lying IN a padded face plane, which is outside
            // the real surface by ~50x the rounding of the exact test, so dropping that NaN (fminf/fmaxf) can only reject
            // true misses.
#if SVGF_RT_SLAB_PACKED
            float tn, tf;
            slab_params(*reinterpret_cast<const float4 *>(g.aabb_min), *reinterpret_cast<const float4 *>(g.aabb_max), ray, invdir, tn, tf);
#else
            const float ax = (g.aabb_min[0] - ray.origin.x) * invdir.x, bx = (g.aabb_max[0] - ray.origin.x) * invdir.x;
            const float ay = (g.aabb_min[1] - ray.origin.y) * invdir.y, by = (g.aabb_max[1] - ray.origin.y) * invdir.y;
            const float az = (g.aabb_min[2] - ray.origin.z) * invdir.z, bz = (g.aabb_max[2] - ray.origin.z) * invdir.z;
            const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
            const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
#endif
            if (tf < tn || tf < 0.f) continue;      // the exact test would return -1 (no hit)
            if (g.type == 1) cubes |= 1u << j; else spheres |= 1u << j;
            if (tn < near_tn) { near_tn = tn; near_j = j; }
        }
        // One loop for both primitive kinds: head (ray into object space) and tail (hit point back to world space, distance
        // along the ray: intersections.h:52-55,86-91 / 106-111,133-146) are the same code, only the middle differs, so a lane
        // holding a sphere works in the same trips as the lanes holding cubes.
        // Order: the candidate entered first, then index order; once a hit is known, candidates whose bounds the ray enters
        // beyond it are dropped without the exact test -- their hit point lies inside the bounds, so its distance could
        // neither beat nor tie t_min (margin far above rounding). The reference tests them and discards the result; the
        // explicit tie rule below makes the winner independent of the order.
        unsigned cand = cubes | spheres;
        // (a finished light query leaves through the loop's ordinary exit, cand == 0, and is answered after the loops: early
        // `return`s from inside these loops moved the warp's reconvergence points and doubled the kernel time for ALL queries)
        if (lq && base == 0) {
            if (!((cand >> light) & 1u)) { lq_miss = true; cand = 0; }        // the ray misses the light's bounds
            near_j = light;
        }
        if (lq_miss) cand = 0;
#ifdef SVGF_RT_LIGHT_FIRST      // A/B build: closest-hit search as ever, but a shadow query tests the light first so that its distance culls the rest
        if (!lq && light >= 0 && light < 32 && base == 0 && ((cand >> light) & 1u)) near_j = light;
#endif
        while (true) {
            int j = -1;
            while (cand) {      // next candidate that can still matter (cheap; lanes reconverge before the exact test)
                int jj = near_j >= 0 ? near_j : __ffs(cand) - 1;
                near_j = -1;
                cand &= ~(1u << jj);
                if (t_min < FLT_MAX) {
                    const GeomD &g = sc.geoms[base + jj];
                    const float ax = (g.aabb_min[0] - ray.origin.x) * invdir.x, bx = (g.aabb_max[0] - ray.origin.x) * invdir.x;
                    const float ay = (g.aabb_min[1] - ray.origin.y) * invdir.y, by = (g.aabb_max[1] - ray.origin.y) * invdir.y;
                    const float az = (g.aabb_min[2] - ray.origin.z) * invdir.z, bz = (g.aabb_max[2] - ray.origin.z) * invdir.z;
                    const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
                    if (tn * dlen > t_min * 1.0001f + 1e-4f) continue;
                }
                j = jj;
                break;
            }
            if (j < 0) break;
            const int i = base + j;
            const GeomD &g = sc.geoms[i];
            Ray q;
            q.origin = multiplyMV(g.inverseTransform, ray.origin, 1.0f);
            q.direction = normalize(multiplyMV(g.inverseTransform, ray.direction, 0.0f));
            const bool is_cube = (cubes >> j) & 1u;
            float tobj = 0.f; int axis = 0; float sign = 0.f; bool outside = true;
            const bool ok = is_cube ? box_slabs(q, tobj, axis, sign) : sphere_roots(q, tobj, outside);
            if (!ok) {
                if (lq && i == light) { lq_miss = true; cand = 0; }
                continue;
            }
            const F3 osi = getPointOnRay(q, tobj);
            const F3 ip = multiplyMV(g.transform, osi, 1.0f);
            const float t = length(ray.origin - ip);
            if (lq) {
                if (i == light) {
                    if (!(t > 0.0f)) { lq_miss = true; cand = 0; }
                    t_min = t; hit_geom = light;
                } else if (t > 0.0f && (t < t_min || (t == t_min && i < light))) { lq_miss = true; cand = 0; }     // occluded
                continue;
            }
            if (t > 0.0f && (t < t_min || (t == t_min && i < hit_geom))) {
                t_min = t; hit_geom = i; w_kind = is_cube ? 1 : 0; w_axis = axis; w_sign = sign; w_osi = osi; w_outside = outside;
            }
        }
    }
    // Meshes: every MESH geom runs the same traversal of the one global BVH and accepts the closest triangle only if its id is
    // in the geom's range (pathtrace.cu:244-255), so one traversal serves them all. Triangles at or beyond the closest cube/
    // sphere cannot win (margin for the tie rule), which bounds the traversal.
    float mesh_u = 0.f, mesh_v = 0.f; int mesh_owner = -1;
    TriBest tb; tb.t = FLT_MAX; tb.slot = -1; tb.bx = tb.by = 0.f;
    // One traversal call site for both kinds of query. A light query that is still open (the light is hit at t_min) asks for
    // any triangle strictly in front of it; an answered one skips the traversal.
#ifndef SVGF_RT_LQ_NO_BVH_ANY
    const float t_any = lq ? t_min : 0.f;
#else
    const float t_any = 0.f;
#endif
    const bool tri_hit = any_mesh && !lq_miss && intersectBVH(sc, ray, invdir, t_min * 1.0001f + 1e-4f, tb, t_any);
    if (lq) {
        // (closest-hit rule for the bounded traversal: a triangle wins with 0 < t < t_min; t == t_min loses to geom 0 by index)
        if (lq_miss || (tri_hit && tb.t > 0.0f && tb.t < t_min)) { is.t = -1.0f; is.geomId = -1; return false; }
        is.t = t_min; is.geomId = light; is.materialId = sc.geoms[light].materialid;     // normal and uv are not read by the caller
        return true;
    }
    if (tri_hit) {
        const int tri_id = __float_as_int(__ldg(&sc.tri_hot[3 * tb.slot]).w);
#pragma unroll 1
        for (int i = 0; i < sc.n_geoms && mesh_owner < 0; i++) {
            const GeomD &g = sc.geoms[i];
            if (g.type == 2 && tri_id >= g.tri_begin && tri_id < g.tri_end) mesh_owner = i;
        }
        if (mesh_owner >= 0) {
            // uv in the correct barycentric order (sceneStructs.h:162-164); the reference's tmp_uv keeps this value for later geoms
            const float4 c0 = __ldg(&sc.tri_cold[4 * tb.slot]), c1 = __ldg(&sc.tri_cold[4 * tb.slot + 1]);
            const float4 c2 = __ldg(&sc.tri_cold[4 * tb.slot + 2]), c3 = __ldg(&sc.tri_cold[4 * tb.slot + 3]);
            const float w0 = 1.0f - tb.bx - tb.by;
            mesh_u = (c0.w * w0 + c2.w * tb.bx) + c3.y * tb.by;
            mesh_v = (c1.w * w0 + c3.x * tb.bx) + c3.z * tb.by;
            const float t = tb.t;
            if (t > 0.0f && (t < t_min || (t == t_min && mesh_owner < hit_geom))) { t_min = t; hit_geom = mesh_owner; w_kind = 2; }
        }
    }
    if (hit_geom == -1) { is.t = -1.0f; is.geomId = -1; return false; }
    F3 normal;
    if (w_kind == 1) normal = box_normal(sc.geoms[hit_geom], w_axis, w_sign);
    else if (w_kind == 0) normal = sphere_normal(sc.geoms[hit_geom], w_osi, w_outside);
    else {      // Triangle::Intersect's normal, in the reference's permuted barycentric order (sceneStructs.h:168-172)
        const float4 c0 = __ldg(&sc.tri_cold[4 * tb.slot]), c1 = __ldg(&sc.tri_cold[4 * tb.slot + 1]), c2 = __ldg(&sc.tri_cold[4 * tb.slot + 2]);
        const float wn = 1.f - tb.bx - tb.by;
        normal = normalize((mk(c0.x, c0.y, c0.z) * tb.bx + mk(c1.x, c1.y, c1.z) * tb.by) + mk(c2.x, c2.y, c2.z) * wn);
    }
    is.t = t_min; is.materialId = sc.geoms[hit_geom].materialid; is.n = normal; is.geomId = hit_geom;
    // uv: a mesh hit's own; for a cube/sphere the reference hands over its stale tmp_uv (pathtrace.cu:226,251,263) = the uv of a
    // mesh candidate processed EARLIER in index order, else nothing (uninitialised there, 0 here)
    const bool uv_from_mesh = mesh_owner >= 0 && mesh_owner <= hit_geom;
    is.u = uv_from_mesh ? mesh_u : 0.f; is.v = uv_from_mesh ? mesh_v : 0.f;
    return true;
}

// Texture::getColor, sceneStructs.h:208-221 (nearest texel, RGB8)
__device__ F3 textureColor(const TexD &tx, float u, float v) {
    int X = (int)gmin(1.f * tx.w * u, 1.f * tx.w - 1.0f);
    int Y = (int)gmin(1.f * tx.h * (1.0f - v), 1.f * tx.h - 1.0f);
    int texel = Y * tx.w + X;
    if (tx.comp != 3) return mk(0, 0, 0);
    // uv < 0 reads out of bounds in the reference (undefined); clamp so the kernel cannot fault
    const int n = tx.w * tx.h;
    texel = texel < 0 ? 0 : (texel >= n ? n - 1 : texel);
    const unsigned char *p = tx.px + 3 * (size_t)texel;
    F3 col = mk((float)p[0], (float)p[1], (float)p[2]);
    return COLORDIVIDOR_F * col;
}
__device__ __forceinline__ F3 materialAlbedo(const SceneView &sc, const svgf_material &m, float u, float v) {
    return m.texid == -1 ? mk(m.color[0], m.color[1], m.color[2]) : textureColor(sc.textures[m.texid], u, v);
}

// computeShadowRay, pathtrace.cu:284-297; glm::rotation (gtx/quaternion.inl:248-283), quat*vec3 (gtc/quaternion.inl:319-326)
// (noinline: one compiled copy shared by the state-machine kernel and the wavefront stages, see add_direct_light)
__device__ __noinline__ void computeShadowRay(Ray &sr, F3 ipos, F3 inrm, F3 lightPos, float lightRadius, float &expectDist, unsigned int &seed) {
    const F3 originPos = ipos + 1e-4f * inrm;       // pathtrace.cu:366
    F3 dirToCenter = normalize(lightPos - originPos);
    const F3 orig = mk(0.0f, 0.0f, 1.0f);
    float qw; F3 qv;
    float cosTheta = dot(orig, dirToCenter);
    if (cosTheta < -1.0f + FLT_EPSILON) {
        F3 axis = cross(mk(0, 0, 1), orig);
        if (dot(axis, axis) < FLT_EPSILON) axis = cross(mk(1, 0, 0), orig);
        axis = normalize(axis);
        const float a = 3.14159265358979323846264338327950288f;
        const float s = sinf(a * 0.5f);
        qw = cosf(a * 0.5f); qv = mk(axis.x * s, axis.y * s, axis.z * s);
    } else {
        F3 axis = cross(orig, dirToCenter);
        float s = sqrtf((1.0f + cosTheta) * 2.0f);
        float invs = 1.0f / s;
        qw = s * 0.5f; qv = mk(axis.x * invs, axis.y * invs, axis.z * invs);
    }
    float theta = 2 * PI_F * nextRand(seed);
    float st, ct; sincosf(theta, &st, &ct);
    F3 v = mk(ct, st, 0.0f);
    F3 uv = cross(qv, v);
    F3 uuv = cross(qv, uv);
    F3 sampleDirection = v + ((uv * qw) + uuv) * 2.0f;
    float sampleRadius = nextRand(seed) * lightRadius;
    F3 samplePoint = lightPos + sampleDirection * sampleRadius;
    expectDist = length(samplePoint - originPos);
    sr.origin = originPos;
    sr.direction = normalize(samplePoint - originPos);
}

// interactions.h:37-67
__device__ F3 calculateRandomDirectionInHemisphere(F3 normal, unsigned int &seed) {
    float up = sqrtf(nextRand(seed));
    float over = sqrtf(1 - up * up);
    float around = nextRand(seed) * TWO_PI_F;
    F3 notNormal;
    if (fabsf(normal.x) < SQRT_OF_ONE_THIRD_F) notNormal = mk(1, 0, 0);
    else if (fabsf(normal.y) < SQRT_OF_ONE_THIRD_F) notNormal = mk(0, 1, 0);
    else notNormal = mk(0, 0, 1);
    F3 p1 = normalize(cross(normal, notNormal));
    F3 p2 = normalize(cross(normal, p1));
    float sa, ca; sincosf(around, &sa, &ca);
    return (up * normal + ca * over * p1) + sa * over * p2;
}

struct PathState { Ray ray; F3 color; bool diffuse; };

// scatterRay, interactions.h:94-136
__device__ __noinline__ void scatterRay(PathState &ps, F3 intersect, F3 normal, const svgf_material &m, unsigned int &seed) {
    ps.ray.origin = intersect + 1e-4f * normal;
    const F3 spec = mk(m.specular_color[0], m.specular_color[1], m.specular_color[2]);
    if (m.hasRefractive) {
        float eta = 1.0f / m.indexOfRefraction;
        float unit_projection = dot(ps.ray.direction, normal);
        if (unit_projection > 0) eta = 1.0f / eta;
        // powf(x, 2) and powf(x, 5) as products: <= 2 ulp from libdevice's powf, far inside the parity budget, and ~300
        // instructions less code in a kernel whose instruction-cache footprint matters
        const float r0 = (1.0f - eta) / (1.0f + eta), c1 = 1 - gabs(unit_projection), c2 = c1 * c1;
        float R0 = r0 * r0;
        float R = R0 + (1 - R0) * (c2 * c2 * c1);
        if (R < nextRand(seed)) {
            F3 I = ps.ray.direction, N = normal;        // glm::refract, detail/func_geometric.inl:189-198
            float dotValue = dot(N, I);
            float k = 1.0f - eta * eta * (1.0f - dotValue * dotValue);
            ps.ray.direction = (eta * I - (eta * dotValue + sqrtf(k)) * N) * (float)(k >= 0.0f);
        } else {
            F3 I = ps.ray.direction, N = normal;        // glm::reflect
            ps.ray.direction = I - N * dot(N, I) * 2.0f;
            ps.color = ps.color * spec;
        }
    } else if (nextRand(seed) < m.hasReflective) {
        F3 I = ps.ray.direction, N = normal;
        ps.ray.direction = I - N * dot(N, I) * 2.0f;
        ps.color = ps.color * spec;
    } else {
        ps.ray.direction = calculateRandomDirectionInHemisphere(normal, seed);
        ps.diffuse = true;
    }
}

// Radiance accumulation, compiled ONCE (noinline) so that the state-machine kernel and the wavefront stages contract
// the same multiplies into the same FMAs and stay bit-identical to each other.
__device__ __noinline__ F3 add_direct_light(F3 acc, F3 throughput, const svgf_material &light, float sintensity, float expectDist,
                                            F3 shadow_dir, F3 normal) {        // pathtrace.cu:377-382
    const float diffuse = gmax(0.0f, dot(shadow_dir, normal));
    const float shadowIntensity = sintensity / (expectDist * expectDist);    // pow(d, 2.0f): x*x is the correctly rounded square
    return acc + throughput * light.emittance * mk(light.color[0], light.color[1], light.color[2]) * shadowIntensity * diffuse;
}
__device__ __noinline__ F3 add_emission(F3 acc, F3 throughput, const svgf_material &m) {       // pathtrace.cu:333
    return acc + throughput * mk(m.color[0], m.color[1], m.color[2]) * m.emittance;
}

__device__ __noinline__ F3 hit_point(F3 origin, F3 direction, float t) { return origin + t * direction; }     // pathtrace.cu:317,338

// One thread per pixel, block = 8x16 pixel tile (a warp covers an 8x4 patch: coherent primary rays).
//
// The reference's rt kernel inlines its closest-hit routine three times (primary, shadow, bounce: pathtrace.cu:314,370,390);
// compiled that way the kernel is ~80 KB of SASS and spends half its time waiting for instruction fetches, because the
// lanes of a warp sit in different copies. Here the path is a small state machine around ONE closest-hit call site:
// every trip of the loop intersects "the current ray" -- the path ray or the shadow ray -- so lanes that are in
// different phases of their path still execute the same intersection code together, and the whole kernel fits the
// instruction cache. Per-pixel results are unchanged: the RNG is re-seeded from (pixel, frame + depth) at every depth.
// Block (and with it warp) shape in pixels: a warp covers RT_BX x 32/RT_BX pixels. 8 x 4 measured best on B200 (tools/build_rt_ab.sh
// builds 4 x 8, 16 x 2 and 32 x 1 through these macros; profiles/r2_ab_rt_warp_shape.jsonl).
#ifndef SVGF_RT_BX
#define SVGF_RT_BX 8
#define SVGF_RT_BY 16
#endif
constexpr int RT_BX = SVGF_RT_BX, RT_BY = SVGF_RT_BY;
enum { Q_PATH = 0, Q_SHADOW = 1 };

// Sharded frames: the a-trous view of the G-buffer is needed up to 2 * 2^levels rows beyond a strip. The rows of MY strip
// within that reach of a neighbour's go into the neighbour's planes as they are produced (dual stores over NVLink), so no
// copy kernel runs between the path tracer and the temporal pass.
struct RtPush {
    HaloPeers peers;
    float4 *gnp[SVGF_MAX_RANKS - 1]; float2 *gzl[SVGF_MAX_RANKS - 1];
};

// ML: "light_sampling_all" (the reference samples geoms[0] only, pathtrace.cu:359-361): every shadow ray
// picks one of the emissive cubes/spheres uniformly (one more random number per shadow ray) and its contribution is scaled by
// their number. A separate instantiation: the default kernel carries none of it.
// CP: "compaction" (A/B, SVGF_RT_COMPACT=1). Paths end at different depths (the light, the open side of the room, the depth limit), and
// a warp runs until its longest path has ended: on cornell a third of the lanes of the warps that still run are dead. With CP the
// block meets after every bounce; when its live paths fit into fewer warps than they occupy, they move (20 words each, through shared
// memory) into the lowest threads of the block and the freed warps only attend the barriers from then on. Which thread carries a
// path cannot change its pixel: everything is keyed by the pixel index, the RNG is re-seeded from (pixel, frame + depth).
constexpr int CP_WORDS = 20, CP_WARPS = RT_BX * RT_BY / 32;
__host__ __device__ inline size_t cp_smem_offset(size_t scene_bytes) { return (scene_bytes + 15) & ~(size_t)15; }
__host__ __device__ inline size_t cp_smem_bytes() { return 2 * CP_WARPS * sizeof(int) + (size_t)CP_WORDS * RT_BX * RT_BY * sizeof(unsigned); }

template <int MINB, bool PUSH, bool ML, bool CP = false>
__global__ void __launch_bounds__(RT_BX *RT_BY, MINB)
rt_kernel(RtParams P, const GeomD *__restrict__ g_geoms, int n_geoms, const svgf_material *__restrict__ g_materials,
          int n_materials, const float4 *__restrict__ bvh, int n_nodes, const float4 *__restrict__ tri_hot,
          const float4 *__restrict__ tri_cold, const TexD *__restrict__ textures,
          float4 *__restrict__ nrm_out, float4 *__restrict__ pos_out, float4 *__restrict__ alb_out,
          float *__restrict__ image, float4 *__restrict__ stale_nm, float2 *__restrict__ stale_uv,
          float4 *__restrict__ gnp_out, float2 *__restrict__ gzl_out, const __grid_constant__ RtPush push) {
    extern __shared__ __align__(16) unsigned char smem[];
    GeomS *s_geoms; svgf_material *s_mats;
    static_assert(!(CP && PUSH), "the compacting kernel does not push halo rows");
    const int tid = threadIdx.y * RT_BX + threadIdx.x;
    stage_scene(smem, g_geoms, n_geoms, g_materials, n_materials, tid, RT_BX * RT_BY, s_geoms, s_mats);
    __syncthreads();
    const int x = blockIdx.x * RT_BX + threadIdx.x;
    const int y = P.row_begin + blockIdx.y * RT_BY + threadIdx.y;
    bool alive = !(x >= P.W || y >= P.row_end);
    if (!CP && !alive) return;
    int idx = x + y * P.W;
    int *cp_cnt = reinterpret_cast<int *>(smem + cp_smem_offset(scene_smem_bytes(n_geoms, n_materials)));
    unsigned *cp_state = reinterpret_cast<unsigned *>(cp_cnt + 2 * CP_WARPS);

    SceneView sc;
    sc.geoms = s_geoms; sc.n_geoms = n_geoms; sc.materials = s_mats; sc.bvh = bvh; sc.n_nodes = n_nodes;
    sc.light_query_ok = s_geoms[0].pad_ != 0.f;      // set at upload (api.cu: upload_scene)
    sc.tri_hot = tri_hot; sc.tri_cold = tri_cold; sc.textures = textures;

    // generateRayFromCamera, pathtrace.cu:187-208
    const svgf_camera &cam = P.cam;
    PathState seg;
    seg.ray.origin = mk(cam.position[0], cam.position[1], cam.position[2]);
    seg.color = mk(1.0f, 1.0f, 1.0f);
    seg.ray.direction = normalize(mk(cam.view[0], cam.view[1], cam.view[2])
        - mk(cam.right[0], cam.right[1], cam.right[2]) * cam.pixelLength[0] * ((float)x - (float)(P.W * 0.5f - 0.5f))
        - mk(cam.up[0], cam.up[1], cam.up[2]) * cam.pixelLength[1] * ((float)y - (float)(P.H * 0.5f - 0.5f)));
    seg.diffuse = false;

    Isect is;               // the persistent per-pixel ShadeableIntersection (pathtrace.cu:85,310)
    is.t = 0.f; is.n = mk(0, 0, 0); is.materialId = 0; is.geomId = 0; is.u = is.v = 0.f;
    bool any_hit = false, stale_loaded = false;
    F3 acc = mk(0, 0, 0);
    int depth = 0;          // 0: the primary query is in flight
    int kind = Q_PATH;
    Ray cur = seg.ray;
    // shading context carried across a shadow query (pathtrace.cu:358-392 uses them after the shadow test)
    unsigned int seed = 0; F3 ipos = mk(0, 0, 0), inrm = mk(0, 0, 0); float expectDist = 0.f;
    int shadow_light = 0;   // the geom the shadow ray in flight aims at (ML only; otherwise geoms[0])
    bool pushed_rows = false;   // PUSH: this thread stored G-buffer rows into a neighbour's planes

    for (int round = 0;; round++) {     // CP: one trip per bounce of the block; otherwise a single trip
    if (!CP || alive) {
    bool finished = true;               // the path has ended (every exit of the loop below but the CP one at its bottom)
    while (true) {
        Isect res; res.geomId = -2; res.materialId = 0; res.t = 0.f; res.n = mk(0, 0, 0); res.u = res.v = 0.f;
        const int lightIdx = ML ? shadow_light : 0;
        const bool hit = computeIntersection(sc, cur, res, kind == Q_SHADOW ? lightIdx : -1);      // <- the only closest-hit call site
        if (kind == Q_SHADOW) {
            if (res.geomId == lightIdx) {                        // pathtrace.cu:374-384 (lightIdx == 0)
                const svgf_material &sm = sc.materials[res.materialId];
                const float si = ML ? P.sintensity * (float)P.n_lights : P.sintensity;
                if (sm.emittance > 0.0f) acc = add_direct_light(acc, seg.color, sm, si, expectDist, cur.direction, inrm);
            }
        } else {
            // the path ray's result lands in the persistent record; a miss only touches t and geomId (pathtrace.cu:267-271)
            if (hit) { is = res; any_hit = true; }
            else {
                is.t = -1.0f; is.geomId = -1;
                if (depth == 0 && !stale_loaded) {               // stale record of earlier frames (pathtrace.cu:119-120,316-322)
                    const float4 s4 = stale_nm[idx]; const float2 suv = stale_uv[idx];
                    is.n = mk(s4.x, s4.y, s4.z); is.materialId = __float_as_int(s4.w); is.u = suv.x; is.v = suv.y;
                    stale_loaded = true;
                }
            }
            if (depth == 0) {                                    // G-buffer from the primary hit, pathtrace.cu:316-323
                const svgf_material &material = sc.materials[is.materialId];
                const F3 p = hit_point(seg.ray.origin, seg.ray.direction, is.t);
                const F3 a = materialAlbedo(sc, material, is.u, is.v);
                nrm_out[idx] = make_float4(is.n.x, is.n.y, is.n.z, __int_as_float(is.geomId));
                pos_out[idx] = make_float4(p.x, p.y, p.z, 0.f);
                alb_out[idx] = make_float4(a.x, a.y, a.z, 0.f);
                const float4 gnp = make_float4(is.n.x * P.kn, p.x * P.kx, is.n.y * P.kn, p.y * P.kx);
                const float2 gzl = make_float2(is.n.z * P.kn, p.z * P.kx);
                gnp_out[idx] = gnp; gzl_out[idx] = gzl;
                if (PUSH) {     // compile-time: the single-GPU kernel carries none of this
                    unsigned m = halo_targets(push.peers, y);
                    // (the system-scope fence that publishes these rows is taken once, after the path has ended: inside this
                    // divergent loop it held up the whole warp for an NVLink round trip, +13 % kernel time on 2 GPUs at 4K)
                    for (; m; m &= m - 1) { const int i = __ffs(m) - 1; push.gnp[i][idx] = gnp; push.gzl[i][idx] = gzl; pushed_rows = true; }
                }
            }
            depth++;
            // ---- top of the reference's bounce loop for `depth` (pathtrace.cu:325-356) ----
            if (depth > P.max_depth || !hit) break;
            seed = initRand(idx, P.frame + depth);
            const svgf_material &material = sc.materials[is.materialId];
            if (material.emittance > 0.0f) {
                if (!P.trace_shadowray || !P.reduce_var || !seg.diffuse) acc = add_emission(acc, seg.color, material);
                break;
            }
            ipos = hit_point(seg.ray.origin, seg.ray.direction, is.t);
            inrm = is.n;
            const bool materialIsDiffuse = material.hasReflective < 1e-6 && material.hasRefractive < 1e-6;
            if (!(P.denoise && P.sepcolor) || depth > 1) seg.color = seg.color * materialAlbedo(sc, material, is.u, is.v);
            if (P.trace_shadowray && materialIsDiffuse) {        // pathtrace.cu:358-371
                if (ML) shadow_light = P.lights[min((int)(nextRand(seed) * (float)P.n_lights), P.n_lights - 1)];
                const GeomD &light = sc.geoms[ML ? shadow_light : 0];
                computeShadowRay(cur, ipos, inrm, mk(light.translation[0], light.translation[1], light.translation[2]),
                                 P.lightradius, expectDist, seed);
                kind = Q_SHADOW;
                continue;
            }
        }
        // ---- bounce (pathtrace.cu:387-392) ----
        if (depth >= P.max_depth) break;
        scatterRay(seg, ipos, inrm, sc.materials[is.materialId], seed);
        cur = seg.ray;
        kind = Q_PATH;
        if (CP) { finished = false; break; }        // bounce boundary: the block takes stock
    }
    if (finished) {
        float *img = image + 3 * (size_t)idx;
        if (P.denoise) { img[0] = acc.x; img[1] = acc.y; img[2] = acc.z; }
        else {          // running mean, pathtrace.cu:398
            const float f = (float)P.frame, f1 = (float)(P.frame + 1);
            F3 old = mk(img[0], img[1], img[2]);
            F3 nw = old * f / f1 + acc / f1;
            img[0] = nw.x; img[1] = nw.y; img[2] = nw.z;
        }
        if (any_hit) {
            stale_nm[idx] = make_float4(is.n.x, is.n.y, is.n.z, __int_as_float(is.materialId));
            stale_uv[idx] = make_float2(is.u, is.v);
        }
        alive = false;
    }
    }
    if (!CP) break;
    {   // every thread of the block, alive or not, comes through here once per round
        const unsigned bal = __ballot_sync(0xffffffffu, alive);
        const int lane = tid & 31, w = tid >> 5;
        int *cnt = cp_cnt + (round & 1) * CP_WARPS;     // two sets: a fast warp may post round r+1 while a slow one still reads round r
        if (lane == 0) cnt[w] = __popc(bal);
        __syncthreads();
        int total = 0, before = 0, warps_now = 0;
#pragma unroll
        for (int i = 0; i < CP_WARPS; i++) { const int n = cnt[i]; total += n; before += i < w ? n : 0; warps_now += n > 0; }
        if (total == 0) break;
        if (((total + 31) >> 5) < warps_now) {
            constexpr int NT = RT_BX * RT_BY;
            if (alive) {
                unsigned *d = cp_state + before + __popc(bal & ((1u << lane) - 1u));
                d[0 * NT] = (unsigned)idx;
                d[1 * NT] = __float_as_uint(cur.origin.x); d[2 * NT] = __float_as_uint(cur.origin.y); d[3 * NT] = __float_as_uint(cur.origin.z);
                d[4 * NT] = __float_as_uint(cur.direction.x); d[5 * NT] = __float_as_uint(cur.direction.y); d[6 * NT] = __float_as_uint(cur.direction.z);
                d[7 * NT] = __float_as_uint(seg.color.x); d[8 * NT] = __float_as_uint(seg.color.y); d[9 * NT] = __float_as_uint(seg.color.z);
                d[10 * NT] = __float_as_uint(acc.x); d[11 * NT] = __float_as_uint(acc.y); d[12 * NT] = __float_as_uint(acc.z);
                d[13 * NT] = ((unsigned)depth & 0x0fffffffu) | (seg.diffuse ? 0x10000000u : 0u) | (any_hit ? 0x20000000u : 0u) | (stale_loaded ? 0x40000000u : 0u);
                d[14 * NT] = __float_as_uint(is.n.x); d[15 * NT] = __float_as_uint(is.n.y); d[16 * NT] = __float_as_uint(is.n.z);
                d[17 * NT] = (unsigned)is.materialId; d[18 * NT] = __float_as_uint(is.u); d[19 * NT] = __float_as_uint(is.v);
            }
            __syncthreads();
            alive = tid < total;
            if (alive) {
                const unsigned *d = cp_state + tid;
                idx = (int)d[0 * NT];
                cur.origin = mk(__uint_as_float(d[1 * NT]), __uint_as_float(d[2 * NT]), __uint_as_float(d[3 * NT]));
                cur.direction = mk(__uint_as_float(d[4 * NT]), __uint_as_float(d[5 * NT]), __uint_as_float(d[6 * NT]));
                seg.ray = cur;
                seg.color = mk(__uint_as_float(d[7 * NT]), __uint_as_float(d[8 * NT]), __uint_as_float(d[9 * NT]));
                acc = mk(__uint_as_float(d[10 * NT]), __uint_as_float(d[11 * NT]), __uint_as_float(d[12 * NT]));
                const unsigned f = d[13 * NT];
                depth = (int)(f & 0x0fffffffu); seg.diffuse = (f & 0x10000000u) != 0; any_hit = (f & 0x20000000u) != 0; stale_loaded = (f & 0x40000000u) != 0;
                is.n = mk(__uint_as_float(d[14 * NT]), __uint_as_float(d[15 * NT]), __uint_as_float(d[16 * NT]));
                is.materialId = (int)d[17 * NT]; is.u = __uint_as_float(d[18 * NT]); is.v = __uint_as_float(d[19 * NT]);
                kind = Q_PATH;
            }
            // (the next stores into cp_state follow the next round's barrier, which every thread reaches after these loads)
        }
    }
    }
    if (PUSH && pushed_rows) __threadfence_system();     // in the neighbours' memory before any later flag of this rank
}


// ---------------------------------------------------------------------------------------------------------------
// Persistent variant of the state machine (SVGF_RT_VARIANT=persistent, A/B only): paths have different lengths (a mirror
// bounce skips the shadow query, a path that reaches the light or leaves the scene stops), so in the one-pixel-per-thread
// kernel a quarter of the lanes of a warp sit idle waiting for the longest path. Here a lane that finishes its pixel
// immediately takes the next unrendered pixel (warp-aggregated atomic on a work counter, pixels enumerated tile by tile) and
// joins the others at the single closest-hit call site with its primary ray. Which lane renders a pixel cannot change the
// pixel (everything is keyed by the pixel index; tests check bit-equality). MEASURED ON B200: 20-25 % SLOWER than the plain
// state machine (C2 1.58 vs 1.25 ms, C3 4.71 vs 4.04 ms): once lanes of a warp are at different depths of different
// pixels their rays stop being coherent, and the extra divergence inside the intersection and shading code costs more
// than the idle lanes did. Kept as a documented negative result.
__global__ void __launch_bounds__(128, 4)
rt_persistent_kernel(RtParams P, const GeomD *__restrict__ g_geoms, int n_geoms, const svgf_material *__restrict__ g_materials,
                     int n_materials, const float4 *__restrict__ bvh, int n_nodes, const float4 *__restrict__ tri_hot,
                     const float4 *__restrict__ tri_cold, const TexD *__restrict__ textures,
                     float4 *__restrict__ nrm_out, float4 *__restrict__ pos_out, float4 *__restrict__ alb_out,
                     float *__restrict__ image, float4 *__restrict__ stale_nm, float2 *__restrict__ stale_uv,
                     float4 *__restrict__ gnp_out, float2 *__restrict__ gzl_out, unsigned int *__restrict__ work_counter) {
    extern __shared__ __align__(16) unsigned char smem[];
    GeomS *s_geoms; svgf_material *s_mats;
    stage_scene(smem, g_geoms, n_geoms, g_materials, n_materials, threadIdx.x, blockDim.x, s_geoms, s_mats);
    __syncthreads();
    SceneView sc;
    sc.geoms = s_geoms; sc.n_geoms = n_geoms; sc.materials = s_mats; sc.bvh = bvh; sc.n_nodes = n_nodes;
    sc.light_query_ok = s_geoms[0].pad_ != 0.f;      // set at upload (api.cu: upload_scene)
    sc.tri_hot = tri_hot; sc.tri_cold = tri_cold; sc.textures = textures;
    const svgf_camera &cam = P.cam;
    const int lane = threadIdx.x & 31;
    const int tiles_x = (P.W + 7) >> 3, tiles_y = (P.row_end - P.row_begin + 3) >> 2;
    const unsigned int n_items = (unsigned int)tiles_x * (unsigned int)tiles_y * 32u;       // 8x4-pixel tiles, 32 slots each

    bool active = false, exhausted = false;
    int idx = 0;
    PathState seg; seg.ray.origin = seg.ray.direction = seg.color = mk(0, 0, 0); seg.diffuse = false;
    Isect is; is.t = 0.f; is.n = mk(0, 0, 0); is.materialId = 0; is.geomId = 0; is.u = is.v = 0.f;
    bool any_hit = false;
    F3 acc = mk(0, 0, 0);
    int depth = 0, kind = Q_PATH;
    Ray cur = seg.ray;
    unsigned int seed = 0; F3 ipos = mk(0, 0, 0), inrm = mk(0, 0, 0); float expectDist = 0.f;

    while (true) {
        // ---- idle lanes take new pixels (warp-synchronous) ----
        if (!exhausted) {
            const unsigned need = __ballot_sync(0xffffffffu, !active);
            if (need) {
                const int leader = __ffs(need) - 1;
                unsigned int base = 0;
                if (lane == leader) base = atomicAdd(work_counter, (unsigned int)__popc(need));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (base >= n_items) exhausted = true;
                if (!active) {
                    const unsigned int n = base + __popc(need & ((1u << lane) - 1u));
                    const unsigned int tile = n >> 5, w = n & 31u;
                    const int x = (int)(tile % (unsigned int)tiles_x) * 8 + (int)(w & 7u);
                    const int y = P.row_begin + (int)(tile / (unsigned int)tiles_x) * 4 + (int)(w >> 3);
                    if (n < n_items && x < P.W && y < P.row_end) {
                        active = true; idx = x + y * P.W;
                        // generateRayFromCamera, pathtrace.cu:187-208
                        seg.ray.origin = mk(cam.position[0], cam.position[1], cam.position[2]);
                        seg.color = mk(1.0f, 1.0f, 1.0f);
                        seg.ray.direction = normalize(mk(cam.view[0], cam.view[1], cam.view[2])
                            - mk(cam.right[0], cam.right[1], cam.right[2]) * cam.pixelLength[0] * ((float)x - (float)(P.W * 0.5f - 0.5f))
                            - mk(cam.up[0], cam.up[1], cam.up[2]) * cam.pixelLength[1] * ((float)y - (float)(P.H * 0.5f - 0.5f)));
                        seg.diffuse = false;
                        any_hit = false; acc = mk(0, 0, 0); depth = 0; kind = Q_PATH; cur = seg.ray;
                    }
                }
            }
        }
        if (!__any_sync(0xffffffffu, active)) { if (exhausted) break; else continue; }
        if (!active) continue;

        // ---- one step of the path's state machine: intersect the current ray, act on the result ----
        Isect res; res.geomId = -2; res.materialId = 0; res.t = 0.f; res.n = mk(0, 0, 0); res.u = res.v = 0.f;
        const bool hit = computeIntersection(sc, cur, res, kind == Q_SHADOW ? 0 : -1);      // <- the only closest-hit call site
        bool finished = false, bounce = false;
        if (kind == Q_SHADOW) {
            if (res.geomId == 0) {                               // pathtrace.cu:374-384 (lightIdx == 0)
                const svgf_material &sm = sc.materials[res.materialId];
                if (sm.emittance > 0.0f) acc = add_direct_light(acc, seg.color, sm, P.sintensity, expectDist, cur.direction, inrm);
            }
            bounce = true;
        } else {
            if (hit) { is = res; any_hit = true; }
            else {
                is.t = -1.0f; is.geomId = -1;
                if (depth == 0) {                                // stale record of earlier frames (pathtrace.cu:119-120,316-322)
                    const float4 s4 = stale_nm[idx]; const float2 suv = stale_uv[idx];
                    is.n = mk(s4.x, s4.y, s4.z); is.materialId = __float_as_int(s4.w); is.u = suv.x; is.v = suv.y;
                }
            }
            if (depth == 0) {                                    // G-buffer from the primary hit, pathtrace.cu:316-323
                const svgf_material &material = sc.materials[is.materialId];
                const F3 p = hit_point(seg.ray.origin, seg.ray.direction, is.t);
                const F3 a = materialAlbedo(sc, material, is.u, is.v);
                nrm_out[idx] = make_float4(is.n.x, is.n.y, is.n.z, __int_as_float(is.geomId));
                pos_out[idx] = make_float4(p.x, p.y, p.z, 0.f);
                alb_out[idx] = make_float4(a.x, a.y, a.z, 0.f);
                gnp_out[idx] = make_float4(is.n.x * P.kn, p.x * P.kx, is.n.y * P.kn, p.y * P.kx);
                gzl_out[idx] = make_float2(is.n.z * P.kn, p.z * P.kx);
            }
            depth++;
            if (depth > P.max_depth || !hit) finished = true;    // top of the reference's bounce loop (pathtrace.cu:325-326)
            else {
                seed = initRand(idx, P.frame + depth);
                const svgf_material &material = sc.materials[is.materialId];
                if (material.emittance > 0.0f) {
                    if (!P.trace_shadowray || !P.reduce_var || !seg.diffuse) acc = add_emission(acc, seg.color, material);
                    finished = true;
                } else {
                    ipos = hit_point(seg.ray.origin, seg.ray.direction, is.t);
                    inrm = is.n;
                    const bool materialIsDiffuse = material.hasReflective < 1e-6 && material.hasRefractive < 1e-6;
                    if (!(P.denoise && P.sepcolor) || depth > 1) seg.color = seg.color * materialAlbedo(sc, material, is.u, is.v);
                    if (P.trace_shadowray && materialIsDiffuse) {        // pathtrace.cu:358-371
                        const GeomD &light = sc.geoms[0];
                        computeShadowRay(cur, ipos, inrm, mk(light.translation[0], light.translation[1], light.translation[2]),
                                         P.lightradius, expectDist, seed);
                        kind = Q_SHADOW;
                    } else bounce = true;
                }
            }
        }
        if (bounce) {                                            // pathtrace.cu:387-392
            if (depth >= P.max_depth) finished = true;
            else { scatterRay(seg, ipos, inrm, sc.materials[is.materialId], seed); cur = seg.ray; kind = Q_PATH; }
        }
        if (finished) {
            float *img = image + 3 * (size_t)idx;
            if (P.denoise) { img[0] = acc.x; img[1] = acc.y; img[2] = acc.z; }
            else {          // running mean, pathtrace.cu:398
                const float f = (float)P.frame, f1 = (float)(P.frame + 1);
                const F3 nw = mk(img[0], img[1], img[2]) * f / f1 + acc / f1;
                img[0] = nw.x; img[1] = nw.y; img[2] = nw.z;
            }
            if (any_hit) {
                stale_nm[idx] = make_float4(is.n.x, is.n.y, is.n.z, __int_as_float(is.materialId));
                stale_uv[idx] = make_float2(is.u, is.v);
            }
            active = false;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Wavefront variant (SVGF_RT_VARIANT=wavefront): the same per-ray functions, split into stages with SoA ray/hit
// buffers in HBM and queues of live paths compacted with a warp ballot. One "slot" per pixel holds a path's state;
// queues hold slot indices, so compaction moves 4 bytes per path, never the state. Stage kernels are launched for the
// worst case and exit on the device-side queue length, so a frame needs no host synchronisation:
//     generate -> [ intersect(path queue) -> shade -> intersect(shadow queue) -> shade_shadow ] x max_depth
// Every result equals the state-machine kernel's bit for bit (same functions, same inputs, RNG keyed by pixel/frame/depth);
// tests/test_gpu_parity.py::test_wavefront_equals_megakernel checks that. It exists for scenes where many paths die
// early (open scenes); in closed boxes like cornell.txt hardly any path terminates before max depth and the extra HBM
// round trips of the ray state make it slower than the state machine, which stays the default.
struct WfBuffers {
    float4 *ray_o, *ray_d;      // {origin, -} {direction, -}   current query ray of the slot (path or shadow)
    float4 *seg_o, *seg_d;      // the path segment's own ray (kept while a shadow query is in flight)
    float4 *color;              // {throughput rgb, diffuse flag}
    float4 *acc;                // {radiance rgb, depth}
    float4 *ctx0, *ctx1;        // {ipos, seed bits} {inrm, expectDist}
    float4 *hit0, *hit1;        // {t, n} {u, v, materialId bits, geomId bits}; hit1.w == -2 marks "no result yet"
    int *q_path[2], *q_shadow;  // queues of slots
    int *counts;                // [0],[1]: path queue lengths (ping-pong), [2]: shadow queue length
};

__device__ __forceinline__ void wf_push(int *queue, int *count, int slot, bool pred) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) queue[base + __popc(m & ((1u << lane) - 1))] = slot;
}

__device__ __forceinline__ SceneView wf_scene(unsigned char *smem, const GeomD *g_geoms, int n_geoms, const svgf_material *g_materials,
                                              int n_materials, const float4 *bvh, int n_nodes, const float4 *tri_hot,
                                              const float4 *tri_cold, const TexD *textures) {
    GeomS *s_geoms; svgf_material *s_mats;
    stage_scene(smem, g_geoms, n_geoms, g_materials, n_materials, threadIdx.x, blockDim.x, s_geoms, s_mats);
    __syncthreads();
    SceneView sc;
    sc.geoms = s_geoms; sc.n_geoms = n_geoms; sc.materials = s_mats; sc.bvh = bvh; sc.n_nodes = n_nodes;
    sc.light_query_ok = s_geoms[0].pad_ != 0.f;      // set at upload (api.cu: upload_scene)
    sc.tri_hot = tri_hot; sc.tri_cold = tri_cold; sc.textures = textures;
    return sc;
}

struct WfScene { const GeomD *geoms; int n_geoms; const svgf_material *materials; int n_materials; const float4 *bvh; int n_nodes;
                 const float4 *tri_hot, *tri_cold; const TexD *textures; };

__global__ void __launch_bounds__(128)
wf_generate_kernel(RtParams P, WfBuffers B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int npx = P.W * (P.row_end - P.row_begin);
    if (i == 0) { B.counts[0] = npx; B.counts[1] = 0; B.counts[2] = 0; }
    if (i >= npx) return;
    const int x = i % P.W, y = P.row_begin + i / P.W;
    const svgf_camera &cam = P.cam;
    const F3 o = mk(cam.position[0], cam.position[1], cam.position[2]);
    const F3 d = normalize(mk(cam.view[0], cam.view[1], cam.view[2])
        - mk(cam.right[0], cam.right[1], cam.right[2]) * cam.pixelLength[0] * ((float)x - (float)(P.W * 0.5f - 0.5f))
        - mk(cam.up[0], cam.up[1], cam.up[2]) * cam.pixelLength[1] * ((float)y - (float)(P.H * 0.5f - 0.5f)));
    B.ray_o[i] = make_float4(o.x, o.y, o.z, 0.f); B.ray_d[i] = make_float4(d.x, d.y, d.z, 0.f);
    B.seg_o[i] = B.ray_o[i]; B.seg_d[i] = B.ray_d[i];
    B.color[i] = make_float4(1.f, 1.f, 1.f, 0.f);
    B.acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    B.q_path[0][i] = i;
}

__global__ void __launch_bounds__(128)
wf_intersect_kernel(WfScene S, WfBuffers B, const int *queue, const int *count, int light_query) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SceneView sc = wf_scene(smem, S.geoms, S.n_geoms, S.materials, S.n_materials, S.bvh, S.n_nodes, S.tri_hot, S.tri_cold, S.textures);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count) return;
    const int slot = queue[i];
    const float4 o = B.ray_o[slot], d = B.ray_d[slot];
    Ray r; r.origin = mk(o.x, o.y, o.z); r.direction = mk(d.x, d.y, d.z);
    Isect res; res.geomId = -2; res.materialId = 0; res.t = 0.f; res.n = mk(0, 0, 0); res.u = res.v = 0.f;
    const bool hit = computeIntersection(sc, r, res, light_query != 0 ? 0 : -1);
    B.hit0[slot] = make_float4(res.t, res.n.x, res.n.y, res.n.z);
    B.hit1[slot] = make_float4(res.u, res.v, __int_as_float(res.materialId), __int_as_float(hit ? res.geomId : -1));
}

__device__ __forceinline__ void wf_finish(const RtParams &P, float *image, int pix, F3 acc) {
    float *img = image + 3 * (size_t)pix;
    if (P.denoise) { img[0] = acc.x; img[1] = acc.y; img[2] = acc.z; }
    else {
        const float f = (float)P.frame, f1 = (float)(P.frame + 1);
        const F3 nw = mk(img[0], img[1], img[2]) * f / f1 + acc / f1;
        img[0] = nw.x; img[1] = nw.y; img[2] = nw.z;
    }
}

// Processes the result of a PATH query (the body of the reference's bounce loop up to the shadow ray, pathtrace.cu:314-371)
__global__ void __launch_bounds__(128)
wf_shade_kernel(RtParams P, WfScene S, WfBuffers B, int in_q, float4 *__restrict__ nrm_out, float4 *__restrict__ pos_out,
                float4 *__restrict__ alb_out, float4 *__restrict__ gnp_out, float2 *__restrict__ gzl_out, float *__restrict__ image,
                float4 *__restrict__ stale_nm, float2 *__restrict__ stale_uv) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SceneView sc = wf_scene(smem, S.geoms, S.n_geoms, S.materials, S.n_materials, S.bvh, S.n_nodes, S.tri_hot, S.tri_cold, S.textures);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < B.counts[in_q];
    int slot = 0; bool to_shadow = false, to_path = false;
    if (live) {
        slot = B.q_path[in_q][i];
        const int pix = (P.row_begin + slot / P.W) * P.W + slot % P.W;
        const float4 h0 = B.hit0[slot], h1 = B.hit1[slot];
        const bool hit = __float_as_int(h1.w) != -1;
        float4 acc4 = B.acc[slot]; float4 col4 = B.color[slot];
        int depth = (int)acc4.w;
        const float4 so = B.seg_o[slot], sd = B.seg_d[slot];
        PathState seg; seg.ray.origin = mk(so.x, so.y, so.z); seg.ray.direction = mk(sd.x, sd.y, sd.z);
        seg.color = mk(col4.x, col4.y, col4.z); seg.diffuse = col4.w != 0.f;
        F3 acc = mk(acc4.x, acc4.y, acc4.z);
        // persistent record (pathtrace.cu:85): a hit rewrites it, a miss only t/geomId
        Isect is;
        if (hit) { is.t = h0.x; is.n = mk(h0.y, h0.z, h0.w); is.u = h1.x; is.v = h1.y; is.materialId = __float_as_int(h1.z); is.geomId = __float_as_int(h1.w);
                   stale_nm[pix] = make_float4(is.n.x, is.n.y, is.n.z, h1.z); stale_uv[pix] = make_float2(is.u, is.v); }
        else { const float4 s4 = stale_nm[pix]; const float2 suv = stale_uv[pix];
               is.t = -1.0f; is.geomId = -1; is.n = mk(s4.x, s4.y, s4.z); is.materialId = __float_as_int(s4.w); is.u = suv.x; is.v = suv.y; }
        if (depth == 0) {
            const svgf_material &material = sc.materials[is.materialId];
            const F3 p = hit_point(seg.ray.origin, seg.ray.direction, is.t);
            const F3 a = materialAlbedo(sc, material, is.u, is.v);
            nrm_out[pix] = make_float4(is.n.x, is.n.y, is.n.z, __int_as_float(is.geomId));
            pos_out[pix] = make_float4(p.x, p.y, p.z, 0.f);
            alb_out[pix] = make_float4(a.x, a.y, a.z, 0.f);
            gnp_out[pix] = make_float4(is.n.x * P.kn, p.x * P.kx, is.n.y * P.kn, p.y * P.kx);
            gzl_out[pix] = make_float2(is.n.z * P.kn, p.z * P.kx);
        }
        depth++;
        bool done = depth > P.max_depth || !hit;
        if (!done) {
            unsigned int seed = initRand(pix, P.frame + depth);
            const svgf_material &material = sc.materials[is.materialId];
            if (material.emittance > 0.0f) {
                if (!P.trace_shadowray || !P.reduce_var || !seg.diffuse) acc = add_emission(acc, seg.color, material);
                done = true;
            } else {
                const F3 ipos = hit_point(seg.ray.origin, seg.ray.direction, is.t), inrm = is.n;
                const bool materialIsDiffuse = material.hasReflective < 1e-6 && material.hasRefractive < 1e-6;
                if (!(P.denoise && P.sepcolor) || depth > 1) seg.color = seg.color * materialAlbedo(sc, material, is.u, is.v);
                if (P.trace_shadowray && materialIsDiffuse) {
                    const GeomD &light = sc.geoms[0];
                    Ray sr; float expectDist = 0.f;
                    computeShadowRay(sr, ipos, inrm, mk(light.translation[0], light.translation[1], light.translation[2]),
                                     P.lightradius, expectDist, seed);
                    B.ray_o[slot] = make_float4(sr.origin.x, sr.origin.y, sr.origin.z, 0.f);
                    B.ray_d[slot] = make_float4(sr.direction.x, sr.direction.y, sr.direction.z, 0.f);
                    B.ctx0[slot] = make_float4(ipos.x, ipos.y, ipos.z, __uint_as_float(seed));
                    B.ctx1[slot] = make_float4(inrm.x, inrm.y, inrm.z, expectDist);
                    to_shadow = true;
                } else if (depth < P.max_depth) {
                    scatterRay(seg, ipos, inrm, material, seed);
                    B.ray_o[slot] = make_float4(seg.ray.origin.x, seg.ray.origin.y, seg.ray.origin.z, 0.f);
                    B.ray_d[slot] = make_float4(seg.ray.direction.x, seg.ray.direction.y, seg.ray.direction.z, 0.f);
                    B.seg_o[slot] = B.ray_o[slot]; B.seg_d[slot] = B.ray_d[slot];
                    to_path = true;
                } else done = true;
            }
        }
        B.color[slot] = make_float4(seg.color.x, seg.color.y, seg.color.z, seg.diffuse ? 1.f : 0.f);
        B.acc[slot] = make_float4(acc.x, acc.y, acc.z, (float)depth);
        if (done) wf_finish(P, image, pix, acc);
    }
    wf_push(B.q_shadow, &B.counts[2], slot, to_shadow);
    wf_push(B.q_path[in_q ^ 1], &B.counts[in_q ^ 1], slot, to_path);
}

// Processes the result of a SHADOW query, then bounces (pathtrace.cu:373-392)
__global__ void __launch_bounds__(128)
wf_shade_shadow_kernel(RtParams P, WfScene S, WfBuffers B, int out_q, float *__restrict__ image, const float4 *__restrict__ stale_nm) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SceneView sc = wf_scene(smem, S.geoms, S.n_geoms, S.materials, S.n_materials, S.bvh, S.n_nodes, S.tri_hot, S.tri_cold, S.textures);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < B.counts[2];
    int slot = 0; bool to_path = false;
    if (live) {
        slot = B.q_shadow[i];
        const int pix = (P.row_begin + slot / P.W) * P.W + slot % P.W;
        const float4 h1 = B.hit1[slot];
        float4 acc4 = B.acc[slot]; const float4 col4 = B.color[slot];
        const int depth = (int)acc4.w;
        const float4 c0 = B.ctx0[slot], c1 = B.ctx1[slot];
        const F3 ipos = mk(c0.x, c0.y, c0.z), inrm = mk(c1.x, c1.y, c1.z);
        unsigned int seed = __float_as_uint(c0.w);
        PathState seg; seg.color = mk(col4.x, col4.y, col4.z); seg.diffuse = col4.w != 0.f;
        const float4 so = B.seg_o[slot], sd = B.seg_d[slot];
        seg.ray.origin = mk(so.x, so.y, so.z); seg.ray.direction = mk(sd.x, sd.y, sd.z);
        F3 acc = mk(acc4.x, acc4.y, acc4.z);
        if (__float_as_int(h1.w) == 0) {
            const svgf_material &sm = sc.materials[__float_as_int(h1.z)];
            if (sm.emittance > 0.0f) {
                const float4 rd = B.ray_d[slot];
                acc = add_direct_light(acc, seg.color, sm, P.sintensity, c1.w, mk(rd.x, rd.y, rd.z), inrm);
            }
        }
        if (depth < P.max_depth) {
            const int materialId = __float_as_int(stale_nm[pix].w);     // the persistent record's material
            scatterRay(seg, ipos, inrm, sc.materials[materialId], seed);
            B.ray_o[slot] = make_float4(seg.ray.origin.x, seg.ray.origin.y, seg.ray.origin.z, 0.f);
            B.ray_d[slot] = make_float4(seg.ray.direction.x, seg.ray.direction.y, seg.ray.direction.z, 0.f);
            B.seg_o[slot] = B.ray_o[slot]; B.seg_d[slot] = B.ray_d[slot];
            B.color[slot] = make_float4(seg.color.x, seg.color.y, seg.color.z, seg.diffuse ? 1.f : 0.f);
            to_path = true;
        }
        B.acc[slot] = make_float4(acc.x, acc.y, acc.z, (float)depth);
        if (!to_path) wf_finish(P, image, pix, acc);
    }
    wf_push(B.q_path[out_q], &B.counts[out_q], slot, to_path);
}

__global__ void wf_reset_counts_kernel(int *counts, int which_a, int which_b) {
    if (threadIdx.x == 0) { counts[which_a] = 0; if (which_b >= 0) counts[which_b] = 0; }
}

}  // namespace

static cudaError_t launch_pathtrace_wavefront(svgf_ctx *c, const RtParams &p, float4 *nrm_out) {
    const DeviceScene &s = c->scene;
    const int rows = p.row_end - p.row_begin;
    if (rows <= 0) return cudaSuccess;
    const int npx = p.W * rows;
    if (!c->wf_mem) {       // lazily: 10 float4 planes + 3 int queues + counters, sized for the whole frame
        const size_t n = c->px;
        cudaError_t e = cudaMalloc(&c->wf_mem, n * (10 * sizeof(float4) + 3 * sizeof(int)) + 64);
        if (e != cudaSuccess) return e;
    }
    WfBuffers B;
    float4 *f4 = static_cast<float4 *>(c->wf_mem);
    const size_t n = c->px;
    B.ray_o = f4; B.ray_d = f4 + n; B.seg_o = f4 + 2 * n; B.seg_d = f4 + 3 * n; B.color = f4 + 4 * n; B.acc = f4 + 5 * n;
    B.ctx0 = f4 + 6 * n; B.ctx1 = f4 + 7 * n; B.hit0 = f4 + 8 * n; B.hit1 = f4 + 9 * n;
    int *qi = reinterpret_cast<int *>(f4 + 10 * n);
    B.q_path[0] = qi; B.q_path[1] = qi + n; B.q_shadow = qi + 2 * n; B.counts = qi + 3 * n;
    WfScene S{s.geoms, s.n_geoms, s.materials, s.n_materials, s.bvh, s.n_nodes, s.tri_hot, s.tri_cold, s.textures};
    const size_t smem = scene_smem_bytes(s.n_geoms, s.n_materials);
    if (smem > 48 * 1024) {
        cudaFuncSetAttribute(wf_intersect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(wf_shade_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(wf_shade_shadow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    const int blocks = (npx + 127) / 128;
    cudaStream_t st = c->stream;
    wf_generate_kernel<<<blocks, 128, 0, st>>>(p, B);
    int q = 0;
    // bounce k shades path depth k + 1 (incl. its shadow ray); depth == max_depth never scatters, so max_depth rounds
    // finish every path (one round even for max_depth == 0: the primary hit still has to produce the G-buffer)
    const int rounds = p.max_depth > 1 ? p.max_depth : 1;
    for (int k = 0; k < rounds; k++) {
        wf_intersect_kernel<<<blocks, 128, smem, st>>>(S, B, B.q_path[q], &B.counts[q], 0);
        wf_reset_counts_kernel<<<1, 32, 0, st>>>(B.counts, q ^ 1, 2);
        wf_shade_kernel<<<blocks, 128, smem, st>>>(p, S, B, q, nrm_out, c->pos, c->alb, c->gnp, c->gzl, c->image, c->stale_nm, c->stale_uv);
        if (p.trace_shadowray) {
            wf_intersect_kernel<<<blocks, 128, smem, st>>>(S, B, B.q_shadow, &B.counts[2], 1);
            wf_shade_shadow_kernel<<<blocks, 128, smem, st>>>(p, S, B, q ^ 1, c->image, c->stale_nm);
        }
        q ^= 1;
    }
    return cudaGetLastError();
}

// Loads every kernel of this file now (see svgf_preload_kernels, api.cu).
void preload_pathtrace_kernels() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, rt_kernel<8, false, false>); cudaFuncGetAttributes(&a, rt_kernel<8, true, false>); cudaFuncGetAttributes(&a, rt_kernel<7, false, false>);
    cudaFuncGetAttributes(&a, rt_kernel<4, false, false>); cudaFuncGetAttributes(&a, rt_kernel<4, true, false>);
    cudaFuncGetAttributes(&a, rt_kernel<8, false, true>); cudaFuncGetAttributes(&a, rt_kernel<8, true, true>);
    cudaFuncGetAttributes(&a, rt_kernel<7, false, false, true>); cudaFuncGetAttributes(&a, rt_kernel<8, false, false, true>);
    cudaFuncGetAttributes(&a, rt_persistent_kernel);
    cudaFuncGetAttributes(&a, wf_generate_kernel); cudaFuncGetAttributes(&a, wf_intersect_kernel); cudaFuncGetAttributes(&a, wf_reset_counts_kernel);
    cudaFuncGetAttributes(&a, wf_shade_kernel); cudaFuncGetAttributes(&a, wf_shade_shadow_kernel);
    (void)cudaGetLastError();
}

// gbuf_reach > 0 (sharded frames, push mode): rows of the a-trous G-buffer view within that many rows of a neighbour's strip
// also go into the neighbour's planes. The state-machine kernel stores them itself; *pushed says whether it did.
cudaError_t launch_pathtrace(svgf_ctx *c, const RtParams &p, float4 *nrm_out, int gbuf_reach, bool *pushed) {
    if (pushed) *pushed = false;
    if (c->rt_variant == 1 && p.n_lights <= 1) return launch_pathtrace_wavefront(c, p, nrm_out);
    const DeviceScene &s = c->scene;
    if (c->rt_variant == 2 && p.n_lights <= 1) {       // persistent state machine with work refill (A/B; slower: see the kernel's comment)
        const size_t smem = scene_smem_bytes(s.n_geoms, s.n_materials);
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(rt_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        if (p.row_end <= p.row_begin) return cudaSuccess;
        if (!c->rt_counter) {
            cudaError_t e = cudaMalloc((void **)&c->rt_counter, sizeof(unsigned int));
            if (e != cudaSuccess) return e;
            int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            c->rt_blocks = sms * 4;
        }
        cudaError_t e = cudaMemsetAsync(c->rt_counter, 0, sizeof(unsigned int), c->stream);
        if (e != cudaSuccess) return e;
        const long long items = (long long)((p.W + 7) / 8) * ((p.row_end - p.row_begin + 3) / 4);   // tiles of 32 pixels = one warp's first fetch
        const int blocks = (int)std::min<long long>(c->rt_blocks, (items + 3) / 4);
        rt_persistent_kernel<<<blocks, 128, smem, c->stream>>>(p, s.geoms, s.n_geoms, s.materials, s.n_materials, s.bvh, s.n_nodes,
                                                               s.tri_hot, s.tri_cold, s.textures, nrm_out, c->pos, c->alb, c->image,
                                                               c->stale_nm, c->stale_uv, c->gnp, c->gzl, c->rt_counter);
        return cudaGetLastError();
    }
    const size_t smem = scene_smem_bytes(s.n_geoms, s.n_materials);
    const int rows = p.row_end - p.row_begin;
    if (rows <= 0) return cudaSuccess;
    dim3 block(RT_BX, RT_BY), grid((p.W + RT_BX - 1) / RT_BX, (rows + RT_BY - 1) / RT_BY);
    RtPush push; memset(&push, 0, sizeof(push));
    if (gbuf_reach > 0) {
        push.peers = halo_peers(c, gbuf_reach);
        for (int i = 0; i < push.peers.n; i++) { push.gnp[i] = c->p_gnp.p[push.peers.rank[i]]; push.gzl[i] = c->p_gzl.p[push.peers.rank[i]]; }
        if (pushed) *pushed = true;
    }
#define RT_LAUNCH_X(MINB, PUSH, ML, CP)                                                                                          \
    do {                                                                                                                         \
        auto kern = rt_kernel<MINB, PUSH, ML, CP>;                                                                               \
        const size_t sm = CP ? cp_smem_offset(smem) + cp_smem_bytes() : smem;                                                    \
        if (sm > 48 * 1024) {                                                                                                    \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);                    \
            if (e != cudaSuccess) return e;                                                                                      \
        }                                                                                                                        \
        kern<<<grid, block, sm, c->rt_launch_stream ? c->rt_launch_stream : c->stream>>>(p, s.geoms, s.n_geoms, s.materials, s.n_materials, s.bvh, \
                                                                    s.n_nodes, s.tri_hot, s.tri_cold, s.textures, nrm_out, c->pos, \
                                                                    c->alb, c->image, c->stale_nm, c->stale_uv, c->gnp, c->gzl, push); \
    } while (0)
#define RT_LAUNCH(MINB, PUSH, ML) RT_LAUNCH_X(MINB, PUSH, ML, false)
    // Occupancy beats registers here: the kernel waits on dependent fp32 chains and BVH loads, so 8 blocks/SM (64 registers,
    // 84 B of spills) run 15-30 % faster than 4 blocks/SM (110 registers, none); 10 and 12 were slower again (measured on
    // B200: C2 0.93/0.80/0.86/0.89 ms, C3 3.73/2.82/2.85/2.92 ms for 4/8/10/12). SVGF_RT_MINBLOCKS=4 keeps the A/B.
    static const int minb = getenv("SVGF_RT_MINBLOCKS") ? atoi(getenv("SVGF_RT_MINBLOCKS")) : 8;
    // A/B (SVGF_RT_COMPACT, read at svgf_create): 1 = compacting kernel for the scenes that run at 7 blocks/SM, 2 = for every
    // frame that pushes no halo rows (see rt_kernel, CP)
    const int rt_compact = c->rt_compact;
    const bool do_push = push.peers.n > 0;
    if (p.n_lights > 1) { if (do_push) RT_LAUNCH(8, true, true); else RT_LAUNCH(8, false, true); }
    else if (minb == 4) { if (do_push) RT_LAUNCH(4, true, false); else RT_LAUNCH(4, false, false); }
#ifdef SVGF_RT_MINB_AB      // A/B build (tools/build_rt_ab.sh): another occupancy target for the default kernel
    else { if (do_push) RT_LAUNCH(8, true, false); else RT_LAUNCH(SVGF_RT_MINB_AB, false, false); }
#else
    // 7 blocks/SM (72 registers) against 8 (64), measured on B200 (profiles/r2_ab_rt_blocks_per_sm.jsonl): a scene of cubes and
    // spheres with next to no mesh (cornell: 38 triangles, 11 BVH nodes) is short of registers (C2 774 vs 805 us), a scene with
    // real meshes is short of warps to hide the BVH loads behind (room, 819 nodes: 2909 vs 2800 us). 5 and 6 lose on both.
    else if (do_push) RT_LAUNCH(8, true, false);
    else if (rt_compact == 1 && s.n_nodes <= 64 && minb == 8) RT_LAUNCH_X(7, false, false, true);
    else if (rt_compact == 2) RT_LAUNCH_X(8, false, false, true);
    else if (s.n_nodes <= 64 && minb == 8) RT_LAUNCH(7, false, false);
    else RT_LAUNCH(8, false, false);
#endif
#undef RT_LAUNCH
#undef RT_LAUNCH_X
    return cudaGetLastError();
}
