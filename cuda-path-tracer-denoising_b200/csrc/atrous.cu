// atrous.cu -- edge-avoiding a-trous wavelet filter level (ATrousFilter, src/denoise.cu:77-170), sm_100a.
//
// Numerics (all within the 1e-4 relative parity budget, see tests/test_atrous_parity.py):
//   * the three edge-stopping weights  expf(-dl/(sqrt(var)*sc + 1e-6)) * expf(-dn/(sn + 1e-6)) * expf(-dx/(sx + 1e-6))
//     (denoise.cu:143-145, each with an fp64 add and an fp64 divide) are evaluated as ONE
//       ex2.approx( -(dl*kl + dn*kn + dx*kx) ),  k = log2(e) / (sigma + 1e-6)
//     with the per-pixel/per-launch reciprocals k hoisted out of the 25-tap loop (kl is computed in fp64 once per
//     pixel; kn, kx once per launch on the host in fp64). `min(1, expf(-x))` is a no-op for x >= 0.
//   * luminance keeps the reference's fp64 evaluation (denoise.cu:121,138): when the variance is exactly 0 the
//     luminance weight is exp(-|lp-lq| / 1e-6) and a 1-ulp difference in lp would be visible, so lp/lq must be
//     the same floats the reference computes.
//   * variance is double-buffered ("Jacobi"): every read sees the previous level. The reference updates it in place
//     while neighbours read it (denoise.cu:111,153,161), which is a data race; the Jacobi form is what its result
//     converges to when all reads win the race and is the only deterministic, shardable definition.
#include "svgf_internal.h"

namespace {

__device__ __forceinline__ float lum_ref(float r, float g, float b) {
    return (float)(0.2126 * r + 0.7152 * g + 0.0722 * b);      // fp64, denoise.cu:121
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct AtrousK {
    const float4 *cv_in; float4 *cv_out;
    const float4 *nrm, *pos, *alb;
    float *denoised_out, *var_out;
    int W, H, row_begin, row_end, step;
    int is_last, blur_variance, addcolor;
    float sigma_c;          // luminance sigma (ui_sigmal)
    float kn, kx;           // log2(e) / (sigma + 1e-6), fp64 on the host
};

// v1: one thread per pixel, taps straight from L1/L2. Kept as the simple, obviously-correct variant; the tiled
// kernel below is what the frame path uses.
__global__ void __launch_bounds__(256)
atrous_direct_kernel(AtrousK k) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = k.row_begin + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= k.W || y >= k.row_end) return;
    const int W = k.W, H = k.H, step = k.step;
    const int p = x + y * W;
    const float4 cp = __ldg(&k.cv_in[p]);

    float var;
    if (k.blur_variance) {      // 3x3 Gaussian, renormalised at the border (denoise.cu:102-115)
        float sum = 0.0f, sumw = 0.0f;
#pragma unroll
        for (int dy = -1; dy <= 1; dy++)
#pragma unroll
            for (int dx = -1; dx <= 1; dx++) {
                const int lx = x + dx, ly = y + dy;
                if (lx >= 0 && ly >= 0 && lx < W && ly < H) {
                    const float g = (dx == 0 ? 0.5f : 0.25f) * (dy == 0 ? 0.5f : 0.25f);
                    sum += g * __ldg(&k.cv_in[lx + ly * W]).w;
                    sumw += g;
                }
            }
        var = fmaxf(sum / sumw, 0.0f);
    } else {
        var = fmaxf(cp.w, 0.0f);
    }
    const float kl = (float)(1.4426950408889634 / ((double)(sqrtf(var) * k.sigma_c) + 1e-6));
    const float lp = lum_ref(cp.x, cp.y, cp.z);
    const float4 np = __ldg(&k.nrm[p]), pp = __ldg(&k.pos[p]);

    float cr = 0.f, cg = 0.f, cb = 0.f, vs = 0.f, ws = 0.f, w2s = 0.f;
#pragma unroll
    for (int i = -2; i <= 2; i++) {
#pragma unroll
        for (int j = -2; j <= 2; j++) {
            const int xq = x + step * i, yq = y + step * j;
            if (xq >= 0 && xq < W && yq >= 0 && yq < H) {
                const int q = xq + yq * W;
                const float4 cq = __ldg(&k.cv_in[q]), nq = __ldg(&k.nrm[q]), pq = __ldg(&k.pos[q]);
                const float lq = lum_ref(cq.x, cq.y, cq.z);
                const float dnx = nq.x - np.x, dny = nq.y - np.y, dnz = nq.z - np.z;
                const float dpx = pq.x - pp.x, dpy = pq.y - pp.y, dpz = pq.z - pp.z;
                const float dn = sqrtf(dnx * dnx + dny * dny + dnz * dnz);
                const float dx = sqrtf(dpx * dpx + dpy * dpy + dpz * dpz);
                const float e = fabsf(lq - lp) * kl + dn * k.kn + dx * k.kx;
                const float hi = (i == 0 ? 0.375f : ((i == 1 || i == -1) ? 0.25f : 0.0625f));
                const float hj = (j == 0 ? 0.375f : ((j == 1 || j == -1) ? 0.25f : 0.0625f));
                const float w = (hi * hj) * ex2_approx(-e);
                ws += w; w2s += w * w;
                cr += cq.x * w; cg += cq.y * w; cb += cq.z * w;
                vs += cq.w * w * w;
            }
        }
    }
    float4 o;
    if (ws > 1e-5f) { o.x = cr / ws; o.y = cg / ws; o.z = cb / ws; o.w = vs / w2s; }
    else o = cp;
    if (k.is_last) {
        if (k.addcolor) { const float4 a = __ldg(&k.alb[p]); o.x *= a.x; o.y *= a.y; o.z *= a.z; }
        float *d = k.denoised_out + 3 * (size_t)p;
        d[0] = o.x; d[1] = o.y; d[2] = o.z;
        k.var_out[p] = o.w;
    }
    if (k.cv_out) k.cv_out[p] = o;
}

}  // namespace

cudaError_t launch_atrous(svgf_ctx *c, const AtrousArgs &a) {
    const int rows = c->shard.row_end - c->shard.row_begin;
    if (rows <= 0) return cudaSuccess;
    AtrousK k;
    k.cv_in = a.cv_in; k.cv_out = a.cv_out; k.nrm = a.nrm; k.pos = a.pos; k.alb = a.alb;
    k.denoised_out = a.denoised_out; k.var_out = a.var_out;
    k.W = c->W; k.H = c->H; k.row_begin = c->shard.row_begin; k.row_end = c->shard.row_end; k.step = 1 << a.level;
    k.is_last = a.is_last; k.blur_variance = a.blur_variance; k.addcolor = a.addcolor;
    k.sigma_c = a.sigma_c;
    const double log2e = 1.4426950408889634;
    k.kn = (float)(log2e / ((double)a.sigma_n + 1e-6));
    k.kx = (float)(log2e / ((double)a.sigma_x + 1e-6));
    dim3 b(32, 8), g((c->W + 31) / 32, (rows + 7) / 8);
    atrous_direct_kernel<<<g, b, 0, c->stream>>>(k);
    return cudaGetLastError();
}
