// lbvh_core.h -- SURVEY.md 8(f) N3: a linear BVH (Morton codes + Karras' 2012 radix-tree construction) built on the device
// over the triangles the path tracer already holds, in the node format its traversal reads (csrc/pathtrace.cu:intersectBVH --
// the reference's BVH_ArrNode semantics, src/bvhtree.h:48-54: pre-order array, left child = index + 1, `rightchildoffset`,
// leaves with a triangle range). The reference builds its SAH tree once on the host (src/bvhtree.cpp); this is for geometry
// that changes on the device. Closest-hit results do not depend on the tree (only ties between equal t and the 64-deep
// traversal stack do), so a frame rendered with this tree matches one rendered with the reference's.
//
// Every step is a plain function of (index, arrays), compiled for the device (csrc/lbvh.cu launches them one thread per index)
// and for the host, where tests/emu/lbvh_emu.cpp runs the same code index by index: the CPU suite checks the tree (structure,
// bounds, pre-order layout, and a render through the oracle) and the GPU suite checks that the device build produces the
// very same arrays. All float steps use explicitly rounded operations so that both give identical bits.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define LBVH_FN __host__ __device__ __forceinline__
#else
#include <cmath>
#include <cstring>
#define LBVH_FN inline
#endif

struct LbvhF4 { float x, y, z, w; };            // same layout as float4 (node halves, triangle records)

LBVH_FN float lbvh_as_float(int32_t i) {
#ifdef __CUDA_ARCH__
    return __int_as_float(i);
#else
    float f; memcpy(&f, &i, 4); return f;
#endif
}
LBVH_FN int32_t lbvh_as_int(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    int32_t i; memcpy(&i, &f, 4); return i;
#endif
}
// one ulp towards -inf / +inf (finite inputs): bounds are formed from v0, v0 + e1, v0 + e2, whose sums round
LBVH_FN float lbvh_down(float x) { const int32_t i = lbvh_as_int(x); return x == 0.0f ? -1.4e-45f : lbvh_as_float(i > 0 ? i - 1 : i + 1); }
LBVH_FN float lbvh_up(float x) { const int32_t i = lbvh_as_int(x); return x == 0.0f ? 1.4e-45f : lbvh_as_float(i > 0 ? i + 1 : i - 1); }
LBVH_FN float lbvh_min(float a, float b) { return b < a ? b : a; }
LBVH_FN float lbvh_max(float a, float b) { return a < b ? b : a; }
LBVH_FN float lbvh_add(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
LBVH_FN float lbvh_mul(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
LBVH_FN float lbvh_div(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
LBVH_FN int lbvh_clz64(uint64_t v) {
#ifdef __CUDA_ARCH__
    return __clzll((long long)v);
#else
    return v ? __builtin_clzll(v) : 64;
#endif
}

// ---- step 1: triangle bounds {min, max} (6 floats) from the hot record {v0, id} {e1, .} {e2, .} ----
LBVH_FN void lbvh_tri_bounds(const LbvhF4 *tri_hot, int k, float *b6) {
    const LbvhF4 v0 = tri_hot[3 * k], e1 = tri_hot[3 * k + 1], e2 = tri_hot[3 * k + 2];
    const float p0[3] = {v0.x, v0.y, v0.z};
    const float p1[3] = {lbvh_add(v0.x, e1.x), lbvh_add(v0.y, e1.y), lbvh_add(v0.z, e1.z)};
    const float p2[3] = {lbvh_add(v0.x, e2.x), lbvh_add(v0.y, e2.y), lbvh_add(v0.z, e2.z)};
    for (int a = 0; a < 3; a++) {
        b6[a] = lbvh_down(lbvh_min(p0[a], lbvh_min(p1[a], p2[a])));
        b6[3 + a] = lbvh_up(lbvh_max(p0[a], lbvh_max(p1[a], p2[a])));
    }
}

// ---- step 2: 30-bit Morton code of the box centre inside the scene box, and the sort key (code << 32 | triangle slot) ----
LBVH_FN uint32_t lbvh_spread10(uint32_t v) {        // 10 bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
LBVH_FN uint64_t lbvh_key(const float *b6, const float *scene6, int k) {
    uint32_t q[3];
    for (int a = 0; a < 3; a++) {
        const float c = lbvh_mul(0.5f, lbvh_add(b6[a], b6[3 + a]));
        const float ext = lbvh_add(scene6[3 + a], -scene6[a]);
        float u = ext > 0.0f ? lbvh_div(lbvh_add(c, -scene6[a]), ext) : 0.0f;
        u = lbvh_mul(u, 1024.0f);
        q[a] = u >= 1023.0f ? 1023u : (u > 0.0f ? (uint32_t)u : 0u);
    }
    const uint32_t code = (lbvh_spread10(q[0]) << 2) | (lbvh_spread10(q[1]) << 1) | lbvh_spread10(q[2]);
    return ((uint64_t)code << 32) | (uint32_t)k;
}

// ---- step 3: Karras 2012. Node ids: internal 0 .. n-2 (root = 0), leaf of sorted position k = n - 1 + k ----
LBVH_FN int lbvh_delta(const uint64_t *keys, int n, int i, int j) {
    return (j < 0 || j >= n) ? -1 : lbvh_clz64(keys[i] ^ keys[j]);      // keys are unique (they end in the triangle slot)
}
// internal node i: its two children, and the axis its split bit belongs to (x y z x y z ... from the top Morton bit down)
LBVH_FN void lbvh_internal(const uint64_t *keys, int n, int i, int *left, int *right, int *parent, int *axis) {
    const int d = lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1) < 0 ? -1 : 1;
    const int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + (d < 0 ? -1 : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const int lc = lo == gamma ? n - 1 + gamma : gamma, rc = hi == gamma + 1 ? n - 1 + gamma + 1 : gamma + 1;
    left[i] = lc; right[i] = rc; parent[lc] = i; parent[rc] = i;
    // key bit `dnode` (0 = top of the 64): bits 2..31 hold the code, x first; below them the tie-break -> any axis
    axis[i] = (dnode >= 2 && dnode < 32) ? (dnode - 2) % 3 : 0;
}

// ---- step 4: leaf k climbs towards the root; the SECOND child to arrive at a node completes it (bounds, subtree size) ----
// `arrive(node)` returns how many arrivals the node had before this one (atomicAdd on the device, plain increment on the host).
template <class Arrive>
LBVH_FN void lbvh_climb(int k, int n, const uint64_t *keys, const float *tri_b6, const int *left, const int *right, const int *parent,
                        float *node_b6, int *size, Arrive arrive) {
    const int slot = (int)(uint32_t)keys[k];
    int v = n - 1 + k;
    for (int a = 0; a < 6; a++) node_b6[6 * v + a] = tri_b6[6 * slot + a];
    size[v] = 1;
    if (n == 1) return;
    // what the sibling's climber wrote is read through volatile loads: on the device it comes from another SM
    const volatile float *vb = node_b6;
    const volatile int *vs = size;
    while (v != 0) {
        const int p = parent[v];
        if (arrive(p) == 0) return;             // the sibling subtree is not complete yet; its climber will finish this node
        const int lc = left[p], rc = right[p];
        for (int a = 0; a < 3; a++) {
            node_b6[6 * p + a] = lbvh_min(vb[6 * lc + a], vb[6 * rc + a]);
            node_b6[6 * p + 3 + a] = lbvh_max(vb[6 * lc + 3 + a], vb[6 * rc + 3 + a]);
        }
        size[p] = 1 + vs[lc] + vs[rc];
        v = p;
    }
}

// ---- step 5: pre-order position of node v (root 0; a left child follows its parent, a right child follows the left subtree) ----
LBVH_FN int lbvh_preorder(int v, const int *left, const int *parent, const int *size) {
    int idx = 0;
    for (int cur = v; cur != 0;) {
        const int p = parent[cur];
        idx += cur == left[p] ? 1 : 1 + size[left[p]];
        cur = p;
    }
    return idx;
}

// ---- step 6: write node v in the traversal's packed format: {min, count | axis << 16}, {max, first slot | right child} ----
LBVH_FN void lbvh_emit(int v, int n, const int *left, const int *right, const int *parent, const int *size, const int *axis,
                       const float *node_b6, LbvhF4 *out_nodes) {
    const int idx = n == 1 ? 0 : lbvh_preorder(v, left, parent, size);
    const bool leaf = v >= n - 1;
    const int meta = leaf ? 1 : (axis[v] << 16);
    const int off = leaf ? v - (n - 1) : idx + 1 + size[left[v]];       // leaf: sorted position = slot in the reordered triangles
    out_nodes[2 * idx] = LbvhF4{node_b6[6 * v], node_b6[6 * v + 1], node_b6[6 * v + 2], lbvh_as_float(meta)};
    out_nodes[2 * idx + 1] = LbvhF4{node_b6[6 * v + 3], node_b6[6 * v + 4], node_b6[6 * v + 5], lbvh_as_float(off)};
}
