// CUDA-on-host emulation shim. TEST INFRASTRUCTURE ONLY (oracle/): lets the reference's own
// src/denoise.cu and src/pathtrace.cu be compiled by g++ and executed on the CPU in this GPU-less
// container, so that the oracle restatement can be pinned against the reference's own code.
// Kernels become plain functions; <<<grid, block>>> launches are rewritten (oracle/ref/stage.sh) into
// EMU_LAUNCH, which walks blocks and threads in a fixed sequential order:
//   for blockIdx.y, for blockIdx.x, for threadIdx.y, for threadIdx.x
// (z is always 1 in the reference). "Device" memory is host malloc.
#pragma once
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <cstdio>
#include <algorithm>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__

struct uchar4 { unsigned char x, y, z, w; };
struct uint3 { unsigned int x, y, z; };
struct dim3 {
    unsigned int x, y, z;
    dim3(unsigned int x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1) : x(x_), y(y_), z(z_) {}
};

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };

template <typename T> inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)malloc(n ? n : 1); return cudaSuccess; }
inline cudaError_t cudaMemset(void *p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }

// CUDA exposes these overloads in the global namespace for device code.
inline float max(float a, float b) { return fmaxf(a, b); }
inline float min(float a, float b) { return fminf(a, b); }
inline int max(int a, int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }

namespace emu {
struct launch {
    dim3 g, b;
    launch(dim3 g_, dim3 b_) : g(g_), b(b_) {}
    template <typename F> void run(F f) {
        gridDim = g; blockDim = b;
        for (unsigned by = 0; by < g.y; by++)
        for (unsigned bx = 0; bx < g.x; bx++)
        for (unsigned ty = 0; ty < b.y; ty++)
        for (unsigned tx = 0; tx < b.x; tx++) {
            blockIdx.x = bx; blockIdx.y = by; blockIdx.z = 0;
            threadIdx.x = tx; threadIdx.y = ty; threadIdx.z = 0;
            f();
        }
    }
};
}
#define EMU_LAUNCH(kernel, gb, args) emu::launch gb .run([&]() { kernel args; })
