// tools/p2p_probe.cu -- measures how fast a kernel on GPU 0 can read GPU 1's memory in place, for the access shapes
// the sharded a-trous tile loader uses (16/8/4-byte cp.async and plain loads of 16..512 contiguous bytes per group of
// lanes). Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/p2p_probe.cu -o gpurun_out/p2p_probe
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// each group of `lanes_per_chunk` lanes reads one contiguous chunk of lanes_per_chunk*16 bytes; chunks are `stride16` float4 apart
__global__ void read_ld(const float4 *__restrict__ src, float4 *out, size_t n16, int lanes_per_chunk, size_t stride16) {
    size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = gridDim.x * (size_t)blockDim.x;
    float4 acc = make_float4(0, 0, 0, 0);
    for (size_t i = tid; i < n16; i += nth) {
        size_t chunk = i / lanes_per_chunk, lane = i % lanes_per_chunk;
        size_t idx = (chunk * stride16 + lane) % n16;
        float4 v = __ldg(&src[idx]);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (acc.x == 1234.5f) out[0] = acc;
}
__global__ void read_cpasync(const float4 *__restrict__ src, float4 *out, size_t n16, int lanes_per_chunk, size_t stride16) {
    extern __shared__ float4 sm[];
    size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = gridDim.x * (size_t)blockDim.x;
    unsigned s = (unsigned)__cvta_generic_to_shared(&sm[threadIdx.x]);
    for (size_t i = tid; i < n16; i += nth) {
        size_t chunk = i / lanes_per_chunk, lane = i % lanes_per_chunk;
        size_t idx = (chunk * stride16 + lane) % n16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(&src[idx]) : "memory");
    }
    asm volatile("cp.async.commit_group; cp.async.wait_group 0;" ::: "memory");
    if (sm[threadIdx.x].x == 1234.5f) out[0] = sm[threadIdx.x];
}

int main() {
    int nd = 0; CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
    int can = 0; CK(cudaDeviceCanAccessPeer(&can, 0, 1)); printf("canAccessPeer(0,1) = %d\n", can);
    const size_t bytes = 256u << 20, n16 = bytes / 16;
    float4 *remote, *local, *out;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&remote, bytes)); CK(cudaMemset(remote, 0, bytes));
    CK(cudaSetDevice(0)); CK(cudaMalloc(&local, bytes)); CK(cudaMemset(local, 0, bytes)); CK(cudaMalloc(&out, 64));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int lanes[] = {1, 2, 8, 32};
    const size_t strides[] = {1, 64};       // in units of the chunk's own size: 1 = dense, 64 = scattered
    for (int mode = 0; mode < 2; mode++)
    for (int where = 0; where < 2; where++)
    for (int li = 0; li < 4; li++)
    for (int si = 0; si < 2; si++) {
        const float4 *src = where ? remote : local;
        size_t stride16 = (size_t)lanes[li] * strides[si];
        float best = 1e9f;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0));
            if (mode == 0) read_ld<<<148 * 8, 256>>>(src, out, n16, lanes[li], stride16);
            else read_cpasync<<<148 * 8, 256, 256 * 16>>>(src, out, n16, lanes[li], stride16);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        printf("%-8s %-6s chunk %4d B %-9s : %8.1f GB/s\n", mode ? "cp.async" : "ld.nc", where ? "PEER" : "local", lanes[li] * 16,
               si ? "scattered" : "dense", bytes / best / 1e6);
    }
    return 0;
}
