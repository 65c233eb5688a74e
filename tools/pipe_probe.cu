// pipe_probe.cu -- per-SM issue rates of the instructions the a-trous kernel lives on (MUFU.EX2 / MUFU.SQRT / MUFU.RSQ,
// FFMA vs FFMA2), measured with independent dependency chains. Build: nvcc -arch=sm_100a -O3 -o tools/pipe_probe tools/pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP> __device__ __forceinline__ float op1(float x) {
    float y;
    if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    else if (OP == 1) asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    else if (OP == 2) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    else if (OP == 3) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    else if (OP == 4) asm volatile("fma.rn.f32 %0, %1, %1, %1;" : "=f"(y) : "f"(x));
    else asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int OP, int CH> __global__ void __launch_bounds__(256) probe(float *out, int iters, float seed) {
    float v[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) v[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) v[i] = op1<OP>(v[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; i++) s += v[i];
    if (s == 12345.678f) out[0] = s;
}

template <int CH> __global__ void __launch_bounds__(256) probe_ffma2(float *out, int iters, float seed) {
    float2 v[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) v[i] = make_float2(seed + i * 0.001f, seed - i * 0.001f);
    const float2 a = make_float2(0.999f, 1.001f), b = make_float2(1e-3f, -1e-3f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) v[i] = __ffma2_rn(v[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; i++) s += v[i].x + v[i].y;
    if (s == 12345.678f) out[0] = s;
}

// mixed: per trip 3 MUFU + N FFMA-pipe instructions (independent), to see whether MUFU issue steals FMA issue slots
template <int NF, int CH> __global__ void __launch_bounds__(256) probe_mix(float *out, int iters, float seed) {
    float v[CH], f[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { v[i] = seed + i * 0.001f; f[i] = seed - i * 0.002f; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            v[i] = op1<1>(v[i]);
#pragma unroll
            for (int j = 0; j < NF; j++) f[i] = op1<4>(f[i]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; i++) s += v[i] + f[i];
    if (s == 12345.678f) out[0] = s;
}

template <class F> static float time_ms(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 8, iters = 4096;
    float *out; cudaMalloc(&out, 4);
    const double thr = (double)blocks * 256;
    printf("%s, %d SMs, max clock %.0f MHz (rates below assume it)\n", p.name, sms, clk_khz / 1e3);
#define RUN(name, kern, per_iter) { float ms = time_ms([&] { kern<<<blocks, 256>>>(out, iters, 1.5f); }); \
        double ops = thr * iters * (per_iter); printf("%-28s %8.3f ms  %7.1f thread-ops/clk/SM\n", name, ms, ops / (ms * 1e-3) / (clk_khz * 1e3) / sms); }
    RUN("MUFU.EX2", (probe<0, 8>), 8.0);
    RUN("MUFU.SQRT", (probe<1, 8>), 8.0);
    RUN("MUFU.RSQ", (probe<2, 8>), 8.0);
    RUN("MUFU.RCP", (probe<3, 8>), 8.0);
    RUN("MUFU.LG2", (probe<5, 8>), 8.0);
    RUN("FFMA", (probe<4, 8>), 8.0);
    RUN("FFMA2 (instr)", (probe_ffma2<8>), 8.0);
    RUN("1 SQRT + 4 FFMA (all instr)", (probe_mix<4, 8>), 8.0 * 5);
    RUN("1 SQRT + 7 FFMA (all instr)", (probe_mix<7, 8>), 8.0 * 8);
    RUN("1 SQRT + 12 FFMA (all instr)", (probe_mix<12, 8>), 8.0 * 13);
    return 0;
}
