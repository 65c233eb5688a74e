// lbvh.cu -- SURVEY.md 8(f) N3: svgf_rebuild_bvh(): a linear BVH built ON THE DEVICE over the context's triangles, swapped in for
// the tree svgf_create uploaded. One thread per index runs the steps of csrc/lbvh_core.h (which the CPU suite runs index by
// index through tests/emu/lbvh_emu.cpp); the only library call is cub's radix sort of the 64-bit Morton keys.
#include "svgf_internal.h"
#include "lbvh_core.h"

#include <cub/device/device_radix_sort.cuh>

namespace {
static_assert(sizeof(LbvhF4) == sizeof(float4), "LbvhF4 mirrors float4");

__global__ void lbvh_bounds_kernel(const LbvhF4 *tri_hot, int n, float *tri_b6) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) lbvh_tri_bounds(tri_hot, k, tri_b6 + 6 * k);
}
// one block: min/max of all triangle boxes (min and max are exact and order-independent, so the result is deterministic)
__global__ void __launch_bounds__(1024) lbvh_scene_kernel(const float *tri_b6, int n, float *scene6) {
    __shared__ float s[6][1024];
    float m[6] = {3e38f, 3e38f, 3e38f, -3e38f, -3e38f, -3e38f};
    for (int k = threadIdx.x; k < n; k += blockDim.x)
        for (int a = 0; a < 3; a++) { m[a] = lbvh_min(m[a], tri_b6[6 * k + a]); m[3 + a] = lbvh_max(m[3 + a], tri_b6[6 * k + 3 + a]); }
    for (int a = 0; a < 6; a++) s[a][threadIdx.x] = m[a];
    __syncthreads();
    for (int w = 512; w >= 1; w >>= 1) {
        if ((int)threadIdx.x < w)
            for (int a = 0; a < 3; a++) {
                s[a][threadIdx.x] = lbvh_min(s[a][threadIdx.x], s[a][threadIdx.x + w]);
                s[3 + a][threadIdx.x] = lbvh_max(s[3 + a][threadIdx.x], s[3 + a][threadIdx.x + w]);
            }
        __syncthreads();
    }
    if (threadIdx.x < 6) scene6[threadIdx.x] = s[threadIdx.x][0];
}
__global__ void lbvh_keys_kernel(const float *tri_b6, const float *scene6, int n, unsigned long long *keys) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) keys[k] = lbvh_key(tri_b6 + 6 * k, scene6, k);
}
__global__ void lbvh_internal_kernel(const unsigned long long *keys, int n, int *left, int *right, int *parent, int *axis) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n - 1) lbvh_internal(reinterpret_cast<const uint64_t *>(keys), n, i, left, right, parent, axis);
}
__global__ void lbvh_climb_kernel(const unsigned long long *keys, int n, const float *tri_b6, const int *left, const int *right, const int *parent,
                                  float *node_b6, int *size, int *flags) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    lbvh_climb(k, n, reinterpret_cast<const uint64_t *>(keys), tri_b6, left, right, parent, node_b6, size, [flags](int p) {
        __threadfence();                        // this subtree's boxes and sizes first ...
        const int before = atomicAdd(flags + p, 1);
        __threadfence();                        // ... and the sibling's are visible once its arrival is
        return before;
    });
}
__global__ void lbvh_emit_kernel(int n, const int *left, const int *right, const int *parent, const int *size, const int *axis,
                                 const float *node_b6, LbvhF4 *nodes) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < 2 * n - 1) lbvh_emit(v, n, left, right, parent, size, axis, node_b6, nodes);
}
__global__ void lbvh_reorder_kernel(const unsigned long long *keys, int n, const float4 *hot, const float4 *cold, float4 *new_hot, float4 *new_cold) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int slot = (int)(unsigned)keys[k];
    for (int q = 0; q < 3; q++) new_hot[3 * k + q] = hot[3 * slot + q];
    for (int q = 0; q < 4; q++) new_cold[4 * k + q] = cold[4 * slot + q];
}
// ---- refit (svgf_refit_bvh): new vertex data into the records, then the boxes of the SAME tree bottom-up ----
__global__ void refit_tris_kernel(const svgf_triangle *tris, int n, const int *slot_of_input, float4 *hot, float4 *cold) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const svgf_triangle &t = tris[i];
    const int k = slot_of_input[i];
    const float *v0 = t.verts[0].pos, *v1 = t.verts[1].pos, *v2 = t.verts[2].pos;
    // the packing of upload_scene (api.cu): hot {v0,id} {e1} {e2}, cold {n0,u0} {n1,v0} {n2,u1} {v1,u2,v2}; e = v - v0 in fp32
    hot[3 * k] = make_float4(v0[0], v0[1], v0[2], __int_as_float(t.id));
    hot[3 * k + 1] = make_float4(__fsub_rn(v1[0], v0[0]), __fsub_rn(v1[1], v0[1]), __fsub_rn(v1[2], v0[2]), 0.f);
    hot[3 * k + 2] = make_float4(__fsub_rn(v2[0], v0[0]), __fsub_rn(v2[1], v0[1]), __fsub_rn(v2[2], v0[2]), 0.f);
    const float *n0 = t.verts[0].normal, *n1 = t.verts[1].normal, *n2 = t.verts[2].normal;
    cold[4 * k] = make_float4(n0[0], n0[1], n0[2], t.verts[0].uv[0]);
    cold[4 * k + 1] = make_float4(n1[0], n1[1], n1[2], t.verts[0].uv[1]);
    cold[4 * k + 2] = make_float4(n2[0], n2[1], n2[2], t.verts[1].uv[0]);
    cold[4 * k + 3] = make_float4(t.verts[1].uv[1], t.verts[2].uv[0], t.verts[2].uv[1], 0.f);
}
// parent of every node of the pre-order array (left child = index + 1, right child = offset)
__global__ void refit_parents_kernel(const float4 *nodes, int nn, int *parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    if (i == 0) parent[0] = -1;
    const int meta = __float_as_int(nodes[2 * i].w), off = __float_as_int(nodes[2 * i + 1].w);
    if ((meta & 0xffff) == 0) { parent[i + 1] = i; parent[off] = i; }
}
// One thread per LEAF: the union of its triangles' boxes (lbvh_tri_bounds: the points the intersection test itself uses, one ulp
// of margin), then up the tree; the second child to arrive at a node forms the union of both children and goes on (min and max
// are exact, so the result does not depend on who arrives first).
__global__ void refit_boxes_kernel(float4 *nodes, int nn, const float4 *hot, const int *parent, int *flags) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nn) return;
    const int meta = __float_as_int(nodes[2 * v].w), count = meta & 0xffff;
    if (count == 0) return;
    const int first = __float_as_int(nodes[2 * v + 1].w);
    float b[6] = {3e38f, 3e38f, 3e38f, -3e38f, -3e38f, -3e38f};
    for (int q = 0; q < count; q++) {
        float t6[6];
        lbvh_tri_bounds(reinterpret_cast<const LbvhF4 *>(hot), first + q, t6);
        for (int a = 0; a < 3; a++) { b[a] = lbvh_min(b[a], t6[a]); b[3 + a] = lbvh_max(b[3 + a], t6[3 + a]); }
    }
    volatile float4 *vn = nodes;
    vn[2 * v].x = b[0]; vn[2 * v].y = b[1]; vn[2 * v].z = b[2];
    vn[2 * v + 1].x = b[3]; vn[2 * v + 1].y = b[4]; vn[2 * v + 1].z = b[5];
    while (true) {
        const int p = parent[v];
        if (p < 0) return;
        __threadfence();                                    // this subtree's boxes first ...
        if (atomicAdd(flags + p, 1) == 0) return;           // ... the sibling will come by and take them along
        __threadfence();
        const int lc = p + 1, rc = __float_as_int(vn[2 * p + 1].w);
        vn[2 * p].x = lbvh_min(vn[2 * lc].x, vn[2 * rc].x); vn[2 * p].y = lbvh_min(vn[2 * lc].y, vn[2 * rc].y); vn[2 * p].z = lbvh_min(vn[2 * lc].z, vn[2 * rc].z);
        vn[2 * p + 1].x = lbvh_max(vn[2 * lc + 1].x, vn[2 * rc + 1].x); vn[2 * p + 1].y = lbvh_max(vn[2 * lc + 1].y, vn[2 * rc + 1].y);
        vn[2 * p + 1].z = lbvh_max(vn[2 * lc + 1].z, vn[2 * rc + 1].z);
        v = p;
    }
}
// after a rebuild: input triangle i now lives where its old slot went (new slot k holds old slot `low 32 bits of keys[k]`)
__global__ void refit_remap_inv_kernel(const unsigned long long *keys, int n, int *new_slot_of_old) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) new_slot_of_old[(int)(unsigned)keys[k]] = k;
}
__global__ void refit_remap_kernel(int n, const int *new_slot_of_old, int *slot_of_input) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slot_of_input[i] = new_slot_of_old[slot_of_input[i]];
}
__global__ void refit_iota_kernel(int n, int *a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}
}  // namespace

#define LB(call)                                                                    \
    do {                                                                            \
        cudaError_t e_ = (call);                                                    \
        if (e_ != cudaSuccess) {                                                    \
            c->err = std::string(#call) + ": " + cudaGetErrorString(e_);            \
            rc = SVGF_ERR_CUDA; goto done;                                          \
        }                                                                           \
    } while (0)

extern "C" int svgf_rebuild_bvh(svgf_ctx *c) {
    if (!c) return SVGF_ERR_INVALID;
    const int n = c->scene.n_tris;
    if (n <= 0) return SVGF_OK;
    int rc = SVGF_OK;
    cudaStream_t st = c->stream;
    float *tri_b6 = nullptr, *scene6 = nullptr, *node_b6 = nullptr;
    unsigned long long *keys = nullptr, *keys_sorted = nullptr;
    int *ints = nullptr;            // left, right, axis, flags: n each; parent, size: 2n each
    float4 *nodes = nullptr, *new_hot = nullptr, *new_cold = nullptr;
    void *tmp = nullptr; size_t tmp_bytes = 0;
    const int nn = 2 * n - 1, T = 256, gb = (n + T - 1) / T;
    LB(cudaSetDevice(c->device));
    LB(cudaMalloc((void **)&tri_b6, sizeof(float) * 6 * n)); LB(cudaMalloc((void **)&scene6, sizeof(float) * 6));
    LB(cudaMalloc((void **)&node_b6, sizeof(float) * 6 * nn));
    LB(cudaMalloc((void **)&keys, sizeof(unsigned long long) * n)); LB(cudaMalloc((void **)&keys_sorted, sizeof(unsigned long long) * n));
    LB(cudaMalloc((void **)&ints, sizeof(int) * 8 * (size_t)n)); LB(cudaMemsetAsync(ints, 0, sizeof(int) * 8 * (size_t)n, st));
    LB(cudaMalloc((void **)&nodes, sizeof(float4) * 2 * nn));
    LB(cudaMalloc((void **)&new_hot, sizeof(float4) * 3 * n)); LB(cudaMalloc((void **)&new_cold, sizeof(float4) * 4 * n));
    {
        int *left = ints, *right = ints + n, *axis = ints + 2 * n, *flags = ints + 3 * n, *parent = ints + 4 * n, *size = ints + 6 * n;
        const LbvhF4 *hot = reinterpret_cast<const LbvhF4 *>(c->scene.tri_hot);
        lbvh_bounds_kernel<<<gb, T, 0, st>>>(hot, n, tri_b6);
        lbvh_scene_kernel<<<1, 1024, 0, st>>>(tri_b6, n, scene6);
        lbvh_keys_kernel<<<gb, T, 0, st>>>(tri_b6, scene6, n, keys);
        LB(cudaGetLastError());
        LB(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys, keys_sorted, n, 0, 64, st));
        LB(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
        LB(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, keys, keys_sorted, n, 0, 64, st));
        if (n > 1) lbvh_internal_kernel<<<(n - 1 + T - 1) / T, T, 0, st>>>(keys_sorted, n, left, right, parent, axis);
        lbvh_climb_kernel<<<gb, T, 0, st>>>(keys_sorted, n, tri_b6, left, right, parent, node_b6, size, flags);
        lbvh_emit_kernel<<<(nn + T - 1) / T, T, 0, st>>>(n, left, right, parent, size, axis, node_b6, reinterpret_cast<LbvhF4 *>(nodes));
        lbvh_reorder_kernel<<<gb, T, 0, st>>>(keys_sorted, n, c->scene.tri_hot, c->scene.tri_cold, new_hot, new_cold);
        if (c->slot_of_input) {         // svgf_refit_bvh has been used: follow the triangles to their new slots (ints[0..n) is free again)
            refit_remap_inv_kernel<<<gb, T, 0, st>>>(keys_sorted, n, ints);
            refit_remap_kernel<<<gb, T, 0, st>>>(n, ints, c->slot_of_input);
        }
        cudaFree(c->bvh_parent); c->bvh_parent = nullptr;       // another tree: parents are recomputed on the next refit
        LB(cudaGetLastError());
        LB(cudaStreamSynchronize(st));
    }
    // swap the new tree and triangle order in (frames queued earlier on the stream have finished: synchronised above)
    cudaFree(c->scene.bvh); cudaFree(c->scene.tri_hot); cudaFree(c->scene.tri_cold);
    c->scene.bvh = nodes; c->scene.tri_hot = new_hot; c->scene.tri_cold = new_cold; c->scene.n_nodes = nn;
    nodes = nullptr; new_hot = nullptr; new_cold = nullptr;
    c->temporal_done_valid = false;
done:
    cudaFree(tri_b6); cudaFree(scene6); cudaFree(node_b6); cudaFree(keys); cudaFree(keys_sorted); cudaFree(ints); cudaFree(tmp);
    cudaFree(nodes); cudaFree(new_hot); cudaFree(new_cold);
    return rc;
}

// SURVEY.md 8(f) N3, second half: geometry that MOVES keeps its tree. New vertex data for all triangles (in the order they were
// given to svgf_create) goes into the records, and the boxes of the tree in place -- the uploaded SAH tree or the one
// svgf_rebuild_bvh built -- are recomputed bottom-up; topology, leaf contents and therefore the traversal order stay. A tree whose
// triangles moved far apart traverses slowly (overlapping boxes) but stays correct: rebuild when that matters. Stream-ordered
// with the frames; does not synchronise.
extern "C" int svgf_refit_bvh(svgf_ctx *c, const svgf_triangle *triangles, int n_triangles) {
    if (!c || !triangles) return SVGF_ERR_INVALID;
    const int n = c->scene.n_tris, nn = c->scene.n_nodes;
    if (n_triangles != n) { c->err = "svgf_refit_bvh: the triangle count differs from the scene's"; return SVGF_ERR_INVALID; }
    if (n <= 0 || nn <= 0) return SVGF_OK;
    int rc = SVGF_OK;
    cudaStream_t st = c->stream;
    const int T = 256;
    LB(cudaSetDevice(c->device));
    if (!c->slot_of_input) {
        LB(cudaMalloc((void **)&c->slot_of_input, sizeof(int) * n));
        refit_iota_kernel<<<(n + T - 1) / T, T, 0, st>>>(n, c->slot_of_input);      // svgf_create keeps the caller's order
    }
    if (!c->refit_stage) LB(cudaMalloc((void **)&c->refit_stage, sizeof(svgf_triangle) * (size_t)n));
    if (!c->bvh_parent) {
        LB(cudaMalloc((void **)&c->bvh_parent, sizeof(int) * 2 * (size_t)nn));      // parents, then arrival flags
        refit_parents_kernel<<<(nn + T - 1) / T, T, 0, st>>>(c->scene.bvh, nn, c->bvh_parent);
    }
    LB(cudaMemcpyAsync(c->refit_stage, triangles, sizeof(svgf_triangle) * (size_t)n, cudaMemcpyHostToDevice, st));
    refit_tris_kernel<<<(n + T - 1) / T, T, 0, st>>>(static_cast<const svgf_triangle *>(c->refit_stage), n, c->slot_of_input, c->scene.tri_hot, c->scene.tri_cold);
    LB(cudaMemsetAsync(c->bvh_parent + nn, 0, sizeof(int) * (size_t)nn, st));
    refit_boxes_kernel<<<(nn + T - 1) / T, T, 0, st>>>(c->scene.bvh, nn, c->scene.tri_hot, c->bvh_parent, c->bvh_parent + nn);
    LB(cudaGetLastError());
    // an overlapped next frame (option frame_overlap) starts its path tracer on another stream after the previous frame's temporal
    // pass only: make it wait for everything queued on this stream so far, the refit included
    c->temporal_done_valid = false;
done:
    return rc;
}
