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

// generalised mix: per trip NM MUFU of kind OP + NF scalar FFMA (or packed FFMA2 when PACKED) + NI integer adds, independent chains
template <int OP, int NM, int NF, int NI, bool PACKED, int CH> __global__ void __launch_bounds__(256) probe_mix2(float *out, int iters, float seed) {
    float v[CH][NM > 0 ? NM : 1]; float2 f[CH]; int q[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) {
#pragma unroll
        for (int m = 0; m < (NM > 0 ? NM : 1); m++) v[i][m] = seed + i * 0.001f + m * 0.01f;
        f[i] = make_float2(seed - i * 0.002f, seed + i * 0.003f); q[i] = i + (int)seed;
    }
    const float2 a = make_float2(0.999f, 1.001f), b = make_float2(1e-3f, -1e-3f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
#pragma unroll
            for (int m = 0; m < NM; m++) v[i][m] = op1<OP>(v[i][m]);
#pragma unroll
            for (int j = 0; j < NF; j++) {
                if (PACKED) f[i] = __ffma2_rn(f[i], a, b);
                else f[i].x = op1<4>(f[i].x);
            }
#pragma unroll
            for (int j = 0; j < NI; j++) asm volatile("add.s32 %0, %0, %1;" : "+r"(q[i]) : "r"(it));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; i++) { s += f[i].x + f[i].y + (float)q[i]; for (int m = 0; m < (NM > 0 ? NM : 1); m++) s += v[i][m]; }
    if (s == 12345.678f) out[0] = s;
}

// the a-trous "twin" (one tap against two centres, csrc/atrous.cu at_twin): 12 FADD2/FMUL2/FFMA2 for the four squared
// distances, 4 MUFU.SQRT, 3 packed, 2 MUFU.EX2, 8 packed -- on register-resident data, NT independent twins per trip
__device__ __forceinline__ float2 d2_(float2 tx, float2 ty, float2 tz, float2 cx, float2 cy, float2 cz) {
    const float2 dx = __fadd2_rn(tx, cx), dy = __fadd2_rn(ty, cy), dz = __fadd2_rn(tz, cz);
    return __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
}
template <int NT, int THREADS> __global__ void __launch_bounds__(THREADS) probe_twin(float *out, int iters, float seed) {
    float2 cx[2], cy[2], cz[2], lum, kl, aw[NT], aw2[NT], ar[NT], ag[NT], ab[NT], av[NT];
    for (int i = 0; i < 2; i++) { cx[i] = make_float2(seed + i, seed * 2 + i); cy[i] = make_float2(seed * 3 + i, seed - i); cz[i] = make_float2(seed * 0.5f + i, seed * 0.25f - i); }
    lum = make_float2(seed, seed * 1.1f); kl = make_float2(seed * 0.7f, seed * 0.9f);
    for (int t = 0; t < NT; t++) { aw[t] = aw2[t] = ar[t] = ag[t] = ab[t] = av[t] = make_float2(0.f, 0.f); }
    float4 cv = make_float4(seed, seed * 0.3f, seed * 0.6f, seed * 0.1f);
    float2 tx = make_float2(seed * 1.3f, seed * 0.8f), ty = make_float2(seed * 0.2f, seed * 1.8f), tz = make_float2(seed * 0.4f, seed * 1.4f);
    float tl = seed * 0.97f;
    const float2 h = make_float2(0.0625f, 0.25f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int t = 0; t < NT; t++) {
            // a different tap per twin (cheap perturbation so that nothing is loop-invariant)
            const float pert = (float)(it + t) * 1e-3f;
            const float2 ttx = make_float2(tx.x + pert, tx.y - pert);
            const float2 d0 = d2_(ttx, ty, tz, cx[0], cy[0], cz[0]), d1 = d2_(ttx, ty, tz, cx[1], cy[1], cz[1]);
            float dn0 = op1<1>(d0.x), dn1 = op1<1>(d1.x), dp0 = op1<1>(d0.y), dp1 = op1<1>(d1.y);
            const float2 dn = make_float2(dn0, dn1), dp = make_float2(dp0, dp1);
            const float2 dl = __fadd2_rn(make_float2(tl + pert, tl + pert), make_float2(-lum.x, -lum.y));
            const float2 e = __fadd2_rn(__ffma2_rn(make_float2(fabsf(dl.x), fabsf(dl.y)), kl, dn), dp);
            float w0 = op1<0>(-e.x), w1 = op1<0>(-e.y);
            const float2 w = __fmul2_rn(make_float2(w0, w1), h), w2 = __fmul2_rn(w, w);
            aw[t] = __fadd2_rn(aw[t], w); aw2[t] = __fadd2_rn(aw2[t], w2);
            ar[t] = __ffma2_rn(make_float2(cv.x, cv.x), w, ar[t]); ag[t] = __ffma2_rn(make_float2(cv.y, cv.y), w, ag[t]);
            ab[t] = __ffma2_rn(make_float2(cv.z, cv.z), w, ab[t]); av[t] = __ffma2_rn(make_float2(cv.w, cv.w), w2, av[t]);
        }
    }
    float s = 0.f;
    for (int t = 0; t < NT; t++) s += aw[t].x + aw[t].y + aw2[t].x + aw2[t].y + ar[t].x + ar[t].y + ag[t].x + ag[t].y + ab[t].x + ab[t].y + av[t].x + av[t].y;
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
    // cycles per trip and SM sub-partition (8 warps each): time * clock / (iters * CH chains * 8 warps... per SMSP: blocks*8 warps / sms / 4)
#define RUN2(name, OP, NM, NF, NI, PK) { float ms = time_ms([&] { probe_mix2<OP, NM, NF, NI, PK, 4><<<blocks, 256>>>(out, iters, 1.5f); }); \
        double warps_per_smsp = (double)blocks * 8 / sms / 4; double trips = (double)iters * 4 * warps_per_smsp; \
        printf("%-44s %8.3f ms  %6.2f cycles per trip per SMSP\n", name, ms, ms * 1e-3 * clk_khz * 1e3 / trips); }
    RUN2("1 SQRT", 1, 1, 0, 0, false);
    RUN2("1 EX2", 0, 1, 0, 0, false);
    RUN2("1 RSQ", 2, 1, 0, 0, false);
    RUN2("8 FFMA", 1, 0, 8, 0, false);
    RUN2("8 FFMA2", 1, 0, 8, 0, true);
    RUN2("8 IADD", 1, 0, 0, 8, false);
    RUN2("1 SQRT + 8 IADD", 1, 1, 0, 8, false);
    RUN2("1 SQRT + 16 IADD", 1, 1, 0, 16, false);
    RUN2("1 SQRT + 8 FFMA", 1, 1, 8, 0, false);
    RUN2("1 EX2 + 8 FFMA", 0, 1, 8, 0, false);
    RUN2("1 RSQ + 8 FFMA", 2, 1, 8, 0, false);
    RUN2("1 SQRT + 16 FFMA", 1, 1, 16, 0, false);
    RUN2("1 SQRT + 24 FFMA", 1, 1, 24, 0, false);
    RUN2("3 SQRT + 24 FFMA", 1, 3, 24, 0, false);
    RUN2("3 SQRT + 18 FFMA", 1, 3, 18, 0, false);
    RUN2("2 SQRT + 16 FFMA", 1, 2, 16, 0, false);
    RUN2("1 SQRT + 4 FFMA2", 1, 1, 4, 0, true);
    RUN2("1 SQRT + 8 FFMA2", 1, 1, 8, 0, true);
    RUN2("3 SQRT + 8 FFMA2 + 8 IADD", 1, 3, 8, 8, true);
    RUN2("1 SQRT + 4 FFMA + 4 IADD", 1, 1, 4, 4, false);
    RUN2("1 SQRT + 8 FFMA + 8 IADD", 1, 1, 8, 8, false);
    // twins: cycles per twin (= 2 pairs) per SMSP at different occupancies (warps per SMSP) and ILP (independent twins per trip)
#define RUN3(name, NT, THREADS, BPS) { const int nb = sms * (BPS); float ms = time_ms([&] { probe_twin<NT, THREADS><<<nb, THREADS>>>(out, iters, 1.5f); }); \
        double warps_per_smsp = (double)(BPS) * (THREADS) / 32 / 4; double twins = (double)iters * (NT) * warps_per_smsp; \
        printf("%-44s %8.3f ms  %6.2f cycles per twin per SMSP (%g warps/SMSP)\n", name, ms, ms * 1e-3 * clk_khz * 1e3 / twins, warps_per_smsp); }
    RUN3("twin x1, 3 warps/SMSP", 1, 128, 3);
    RUN3("twin x1, 5 warps/SMSP", 1, 128, 5);
    RUN3("twin x1, 8 warps/SMSP", 1, 256, 4);
    RUN3("twin x1, 12 warps/SMSP", 1, 256, 6);
    RUN3("twin x2, 3 warps/SMSP", 2, 128, 3);
    RUN3("twin x2, 5 warps/SMSP", 2, 128, 5);
    RUN3("twin x2, 8 warps/SMSP", 2, 256, 4);
    RUN3("twin x4, 3 warps/SMSP", 4, 128, 3);
    RUN3("twin x4, 5 warps/SMSP", 4, 128, 5);
    RUN3("twin x4, 8 warps/SMSP", 4, 256, 4);
    return 0;
}
