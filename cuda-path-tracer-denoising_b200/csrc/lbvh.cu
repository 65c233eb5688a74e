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
        LB(cudaGetLastError());
        LB(cudaStreamSynchronize(st));
    }
    // swap the new tree and triangle order in (frames queued earlier on the stream have finished: synchronised above)
    cudaFree(c->scene.bvh); cudaFree(c->scene.tri_hot); cudaFree(c->scene.tri_cold);
    c->scene.bvh = nodes; c->scene.tri_hot = new_hot; c->scene.tri_cold = new_cold; c->scene.n_nodes = nn;
    nodes = nullptr; new_hot = nullptr; new_cold = nullptr;
done:
    cudaFree(tri_b6); cudaFree(scene6); cudaFree(node_b6); cudaFree(keys); cudaFree(keys_sorted); cudaFree(ints); cudaFree(tmp);
    cudaFree(nodes); cudaFree(new_hot); cudaFree(new_cold);
    return rc;
}
