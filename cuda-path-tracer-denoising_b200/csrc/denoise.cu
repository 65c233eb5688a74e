// denoise.cu -- SVGF temporal pass, PBO pack and layout conversion kernels, sm_100a.
//
//   temporal_kernel     <- BackProjection + isReprjValid (src/denoise.cu:172-317), fused with the variance
//                          estimate and writing {colour, variance} as one float4; history is read from whichever
//                          buffers the previous frame left (pointer rotation instead of denoise.cu:366,396-398)
//   no_temporal_kernel  <- EstimateVariance + the input->history copy (denoise.cu:320-329, 369-370)
//   pack_pbo_kernel     <- sendTwoImagesToPBO (src/pathtrace.cu:46-78)
//   debug_view_kernel   <- DebugView<int|float> (denoise.cu:331-340, 373-378)
//   aos<->soa           layout conversion for the reference-layout entry point svgf_denoise() and svgf_fetch
// The a-trous filter itself lives in atrous.cu.
#include "svgf_internal.h"
#include "halo_sync.cuh"

#include <algorithm>
#include <cstring>

namespace {

// isReprjValid, denoise.cu:172-182, for a tap at float coordinates (px, py) of the previous frame.
// Returns the linear index of the tap through `q`.
// Which rank's planes hold row `row` of the previous frame. SINGLE (one strip = the whole frame): always mine, and the table lookups
// fold into one pointer per plane. The row of a tap is known from its coordinates; only frames of 2^24 pixels and more, where the
// reference's float index arithmetic stops being exact, divide the (rounded) index.
template <bool SINGLE> struct TapOwner {
    const RowOwner &ro; int me; int W; bool exact;
    __device__ __forceinline__ int operator()(int row, int q) const { return SINGLE ? me : owner_of(ro, exact ? row : q / W); }
};
template <bool SINGLE>
__device__ __forceinline__ bool reprj_valid(int W, int H, float px, float py, const float4 &ncur,
                                            const PeerPtr<float4> &nrm_prev, const TapOwner<SINGLE> &own, int &q) {
    // NaN coordinates pass the reference's bounds test and index garbage (undefined); rejected here.
    if (!(px >= 0.f) || !(px < (float)W) || !(py >= 0.f) || !(py < (float)H)) return false;
    q = (int)(px + py * (float)W);
    const float4 np = __ldg(&nrm_prev.p[own((int)py, q)][q]);
    const int gprev = __float_as_int(np.w), gcur = __float_as_int(ncur.w);
    if (gprev == -1 || gprev != gcur) return false;
    // glm::distance(n_prev, n_cur) = length(n_cur - n_prev) (detail/func_geometric.inl:108-111) > 1e-1f without the square root: the correctly rounded sqrtf is monotonic, and 0x3c23d70b
    // (0.010000001f) is the largest float whose root is still <= 0.1f, so the comparison of the squares decides identically
    // (NaN: false both ways). Saves an IEEE square root (~10 instructions) per tap in a kernel that is short of issue slots.
    const float dx = ncur.x - np.x, dy = ncur.y - np.y, dz = ncur.z - np.z;
    if (dx * dx + dy * dy + dz * dz > __uint_as_float(0x3c23d70bu)) return false;
    return true;
}

struct Mat4 { float m[16]; };

struct TemporalOut { float4 cv; float2 lv, mom; int hlen; };

// One pixel of BackProjection (denoise.cu:185-317) fused with the variance estimate. 108 algorithmic bytes per pixel
// (SURVEY.md 8(d)): reads image 12 + normal/geomId 16 + position 12 + own history length 4 + (reprojected, cache-shared)
// prev normal/geomId 16, colour history 12(16), moments 8, history length 4; writes {colour,variance} 16 + moments 8 +
// history length 4.
// (Requesting the four taps' history records together with their normals -- two dependent round trips to memory instead of
// three -- was measured on B200 and is SLOWER, 69.6 vs 62 us at C2: 64 registers instead of 47 cost a resident block per SM.)
template <bool SINGLE>
__device__ __forceinline__ TemporalOut temporal_pixel(int W, int H, int p, const float *__restrict__ image, const float4 *__restrict__ nrm_cur,
                                                      const PeerPtr<float4> &nrm_prev, const float4 *__restrict__ pos, const PeerPtr<float4> &hist_cv,
                                                      const PeerPtr<float2> &mom_hist, const PeerPtr<int> &hlen_tab, const RowOwner &ro, int me,
                                                      const Mat4 &vm, float color_alpha_min, float moment_alpha_min, float clip_rx, float clip_ry,
                                                      int hist_cap) {
    const TapOwner<SINGLE> own{ro, me, W, (long long)W * H < (1ll << 24)};
    const int N = hlen_tab.p[me][p];     // own pixel (denoise.cu:194)
    const float sr = image[3 * (size_t)p], sg = image[3 * (size_t)p + 1], sb = image[3 * (size_t)p + 2];
    const float4 ncur = nrm_cur[p];
    const float luminance = 0.2126 * sr + 0.7152 * sg + 0.0722 * sb;    // double, as denoise.cu:196
    TemporalOut o;
    if (N > 0 && __float_as_int(ncur.w) != -1) {
        const float4 pp = pos[p];
        // prev_viewmat * vec4(position, 1): (m0 v0 + m1 v1) + (m2 v2 + m3 v3), glm type_mat4x4.inl:617-628
        // (every operation below is spelled out with the rounding and the fusion nvcc gives the reference's expressions -- the
        // product m1 v1 rounded, m0 v0 fused onto it; m2 v2 fused onto m3 * 1 -- because which pixel the history comes from, and
        // with it the integer history length, hangs on the last bit of these coordinates)
        const float vx = __fadd_rn(__fmaf_rn(vm.m[0], pp.x, __fmul_rn(vm.m[4], pp.y)), __fmaf_rn(vm.m[8], pp.z, vm.m[12]));
        const float vy = __fadd_rn(__fmaf_rn(vm.m[1], pp.x, __fmul_rn(vm.m[5], pp.y)), __fmaf_rn(vm.m[9], pp.z, vm.m[13]));
        const float vz = __fadd_rn(__fmaf_rn(vm.m[2], pp.x, __fmul_rn(vm.m[6], pp.y)), __fmaf_rn(vm.m[10], pp.z, vm.m[14]));
        // clip_rx = clip_ry = 1 reproduces the reference, which leaves out the FOV/aspect term (denoise.cu:201-207: exact only
        // for FOVY 45 and square frames; x * 1.0f is exact, so the default path keeps its bits). SURVEY.md 8(f) N4 switch
        // "reprojection_fov_aspect": 1 / (tan(fovy) * aspect) and 1 / tan(fovy), the inverse of generateRayFromCamera's scaling.
        const float clipx = __fmul_rn(__fdiv_rn(vx, vz), clip_rx), clipy = __fmul_rn(__fdiv_rn(vy, vz), clip_ry);
        const float ndcx = __fmaf_rn(clipx, -0.5f, 0.5f), ndcy = __fmaf_rn(clipy, -0.5f, 0.5f);
        const float prevx = __fmaf_rn(ndcx, (float)W, -0.5f), prevy = __fmaf_rn(ndcy, (float)H, -0.5f);
        const float floorx = floorf(prevx), floory = floorf(prevy);
        const float fracx = __fsub_rn(prevx, floorx), fracy = __fsub_rn(prevy, floory);
        bool valid = (floorx >= 0 && floory >= 0 && floorx < W && floory < H);
        // glm::ivec2(floorx, floory) + offset: saturating cvt, wrapping integer add (as the reference's SASS)
        const int ifx = __float2int_rz(floorx), ify = __float2int_rz(floory);
        bool v[4]; int qi[4];
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const int lx = (int)((unsigned)ifx + (unsigned)(s & 1)), ly = (int)((unsigned)ify + (unsigned)(s >> 1));
            qi[s] = 0;
            v[s] = reprj_valid<SINGLE>(W, H, (float)lx, (float)ly, ncur, nrm_prev, own, qi[s]);
            qi[s] = lx + ly * W;
            valid = valid && v[s];
        }
        float pr = 0.f, pg = 0.f, pb = 0.f, pm1 = 0.f, pm2 = 0.f, phl = 0.f;
        if (valid) {
            // The operations are spelled out (one rounding each, fused exactly where nvcc fuses the reference's expressions:
            // `acc += w * h` is an FMA, the weights are plain products, `sumw += w` a plain add): left to the optimiser, another
            // register budget re-associated this block and moved (int)phl across an integer for a few pixels per frame.
            float sumw = 0.0f;
            const float ofx = __fsub_rn(1.0f, fracx), ofy = __fsub_rn(1.0f, fracy);
            const float w[4] = {__fmul_rn(ofx, ofy), __fmul_rn(fracx, ofy), __fmul_rn(ofx, fracy), __fmul_rn(fracx, fracy)};
#pragma unroll
            for (int s = 0; s < 4; s++) {
                if (v[s]) {
                    const int o = SINGLE ? me : owner_of(ro, (int)((unsigned)ify + (unsigned)(s >> 1)));      // a valid tap lies in the image: its row is ly
                    const float4 hc = __ldg(&hist_cv.p[o][qi[s]]); const float2 hm = __ldg(&mom_hist.p[o][qi[s]]);
                    pr = __fmaf_rn(w[s], hc.x, pr); pg = __fmaf_rn(w[s], hc.y, pg); pb = __fmaf_rn(w[s], hc.z, pb);
                    pm1 = __fmaf_rn(w[s], hm.x, pm1); pm2 = __fmaf_rn(w[s], hm.y, pm2);
                    phl = __fmaf_rn(w[s], (float)__ldg(&hlen_tab.p[o][qi[s]]), phl);
                    sumw = __fadd_rn(sumw, w[s]);
                }
            }
            if (sumw >= 0.01) {
                // (x / 1.0f is x: a camera at rest reprojects onto pixel centres, the weights are {1, 0, 0, 0}, and six IEEE
                // divides per pixel go away)
                if (sumw != 1.0f) { pr /= sumw; pg /= sumw; pb /= sumw; pm1 /= sumw; pm2 /= sumw; phl /= sumw; }
                valid = true;
            }
        }
        if (!valid) {
            float cnt = 0.0f;
            for (int yy = -1; yy <= 1; yy++) {
                for (int xx = -1; xx <= 1; xx++) {
                    const float lx = floorx + (float)xx, ly = floory + (float)yy;
                    int q = 0;
                    if (reprj_valid<SINGLE>(W, H, lx, ly, ncur, nrm_prev, own, q)) {
                        q = (int)(lx + (float)W * ly);
                        const int o = own((int)ly, q);
                        const float4 hc = __ldg(&hist_cv.p[o][q]); const float2 hm = __ldg(&mom_hist.p[o][q]);
                        pr = __fadd_rn(pr, hc.x); pg = __fadd_rn(pg, hc.y); pb = __fadd_rn(pb, hc.z); pm1 = __fadd_rn(pm1, hm.x); pm2 = __fadd_rn(pm2, hm.y);
                        phl = __fadd_rn(phl, (float)__ldg(&hlen_tab.p[o][q]));
                        cnt = __fadd_rn(cnt, 1.0f);
                    }
                }
            }
            if (cnt > 0.0f) {
                pr /= cnt; pg /= cnt; pb /= cnt; pm1 /= cnt; pm2 /= cnt; phl /= cnt;
                valid = true;
            }
        }
        if (valid) {
            const float color_alpha = fmaxf(1.0f / (float)(N + 1), color_alpha_min);
            const float moment_alpha = fmaxf(1.0f / (float)(N + 1), moment_alpha_min);
            const int hl = (int)phl + 1;            // unbounded in the reference (denoise.cu:290-294); optional cap (N4)
            o.hlen = hist_cap > 0 ? min(hl, hist_cap) : hl;
            const float first = moment_alpha * pm1 + (1.0f - moment_alpha) * luminance;
            const float second = moment_alpha * pm2 + (1.0f - moment_alpha) * luminance * luminance;
            o.mom = make_float2(first, second);
            const float variance = second - first * first;
            const float ar = sr * color_alpha + pr * (1.0f - color_alpha), ag = sg * color_alpha + pg * (1.0f - color_alpha),
                        ab = sb * color_alpha + pb * (1.0f - color_alpha);
            o.cv = make_float4(ar, ag, ab, variance > 0.0f ? variance : 0.0f);
            // luminance as the a-trous taps will read it (denoise.cu:121,138), and the variance once more for the 3x3 blur
            o.lv = make_float2((float)(0.2126 * ar + 0.7152 * ag + 0.0722 * ab), variance > 0.0f ? variance : 0.0f);
            return o;
        }
    }
    o.hlen = 1;
    o.mom = make_float2(luminance, luminance * luminance);
    o.cv = make_float4(sr, sg, sb, 100.0f);
    o.lv = make_float2(luminance, 100.0f);
    return o;
}

struct TemporalPush {           // sharded frames: the neighbours' copies of the accumulated planes (level 1 taps +-4 rows)
    HaloOut ho;
    float4 *cv[SVGF_MAX_RANKS - 1]; float2 *lv[SVGF_MAX_RANKS - 1];
};

// 8 blocks/SM (32 registers, full occupancy): the kernel sits out three dependent round trips to memory per pixel, and resident
// warps are what hides them (C2 / C5: 63.5 / 62.6 us at the compiler's own 48 registers, 56 / 53 us here,
// profiles/r2_ab_temporal_blocks_per_sm.txt). The hint changes how the optimiser fuses and re-associates the bilinear
// accumulation of temporal_pixel -- with the arithmetic left to it, the history lengths of moving-camera frames stopped matching
// the reference's (18 parity tests) -- which is why that block is written with explicit single-rounding intrinsics.
template <bool SINGLE>
__global__ void __launch_bounds__(256, 8)
temporal_kernel(int W, int H, int row_begin, int row_end, const float *__restrict__ image,
                const float4 *__restrict__ nrm_cur, const __grid_constant__ PeerPtr<float4> nrm_prev, const float4 *__restrict__ pos,
                const __grid_constant__ PeerPtr<float4> hist_cv, const __grid_constant__ PeerPtr<float2> mom_hist,
                const __grid_constant__ PeerPtr<int> hlen_tab, const __grid_constant__ RowOwner ro, int me,
                float4 *__restrict__ acc_cv, float2 *__restrict__ acc_lv, float2 *__restrict__ mom_acc, int *__restrict__ hlen_out, const __grid_constant__ Mat4 vm,
                float color_alpha_min, float moment_alpha_min, float clip_rx, float clip_ry, int hist_cap,
                const __grid_constant__ TemporalPush push) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = row_begin + blockIdx.y * blockDim.y + threadIdx.y;
    if (x < W && y < row_end) {
        const int p = x + y * W;
        const TemporalOut o = temporal_pixel<SINGLE>(W, H, p, image, nrm_cur, nrm_prev, pos, hist_cv, mom_hist, hlen_tab, ro, me, vm,
                                             color_alpha_min, moment_alpha_min, clip_rx, clip_ry, hist_cap);
        hlen_out[p] = o.hlen; mom_acc[p] = o.mom; acc_cv[p] = o.cv; acc_lv[p] = o.lv;
        for (unsigned m = halo_targets(push.ho.peers, y); m; m &= m - 1) {
            const int i = __ffs(m) - 1;
            push.cv[i][p] = o.cv; push.lv[i][p] = o.lv;
        }
    }
    {
        const int by0 = row_begin + blockIdx.y * blockDim.y;
        halo_block_done(push.ho, halo_rows_touch(push.ho.peers, by0, by0 + (int)blockDim.y - 1));
    }
}

// SURVEY.md 8(f) N4, "spatial_variance_estimate" (off by default). The reference's EstimateVariance is a stub (denoise.cu:320-329
// writes 10.0) and BackProjection gives a pixel without history the constant 100 (denoise.cu:315), so freshly disoccluded
// regions are filtered with a luminance weight that is switched off. With the option, a pixel whose history is shorter than
// 4 frames takes its variance from the luminance moments of its 7x7 neighbourhood on the same surface (same geomId, normals
// within ~10 degrees), boosted by 4 / history length -- the estimate of the SVGF paper (Schied et al. 2017, section 4.2).
// Reads moments / normals / history lengths, writes only the pixel's own variance: no ordering between threads is needed.
__global__ void __launch_bounds__(256)
spatial_variance_kernel(int W, int H, const int *__restrict__ hlen, const float2 *__restrict__ mom, const float4 *__restrict__ nrm,
                        float4 *__restrict__ cv, float2 *__restrict__ lv) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int p = x + y * W;
    const int h = hlen[p];
    const float4 nc = nrm[p];
    if (h >= 4 || __float_as_int(nc.w) == -1) return;
    float sw = 0.f, s1 = 0.f, s2 = 0.f;
    for (int dy = -3; dy <= 3; dy++)
        for (int dx = -3; dx <= 3; dx++) {
            const int qx = x + dx, qy = y + dy;
            if (qx < 0 || qy < 0 || qx >= W || qy >= H) continue;
            const int q = qx + qy * W;
            const float4 nq = nrm[q];
            if (__float_as_int(nq.w) != __float_as_int(nc.w)) continue;
            const float d = nc.x * nq.x + nc.y * nq.y + nc.z * nq.z;
            if (!(d > 0.985f)) continue;
            const float2 m = mom[q];
            sw += 1.f; s1 += m.x; s2 += m.y;
        }
    if (sw < 2.f) return;       // nothing but the pixel itself: keep the reference's value
    const float m1 = s1 / sw, m2 = s2 / sw;
    const float var = fmaxf(m2 - m1 * m1, 0.f) * (4.0f / (float)max(h, 1));
    cv[p].w = var; lv[p].y = var;
}

__global__ void __launch_bounds__(256)
no_temporal_kernel(int W, int row_begin, int row_end, const float *__restrict__ image, float4 *__restrict__ acc_cv,
                   float2 *__restrict__ acc_lv) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = row_begin + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= row_end) return;
    const size_t p = x + (size_t)y * W;
    const float r = image[3 * p], g = image[3 * p + 1], b = image[3 * p + 2];
    acc_cv[p] = make_float4(r, g, b, 10.0f);
    acc_lv[p] = make_float2((float)(0.2126 * r + 0.7152 * g + 0.0722 * b), 10.0f);
}

// clamp((int)(v * 255.0), 0, 255) with the reference's DOUBLE product (pathtrace.cu:60-62), evaluated in fp32:
// the float product can only truncate differently from the exact one when it rounded up onto an integer, which the
// FMA residual detects.
__device__ __forceinline__ unsigned int to_u8(float v) {
    const float p = v * 255.0f;
    int i = __float2int_rz(p);
    if ((float)i == p && fmaf(v, 255.0f, -p) < 0.0f) i -= 1;
    return (unsigned int)min(max(i, 0), 255);
}

__global__ void __launch_bounds__(256)
pack_pbo_kernel(int W, int row_begin, int row_end, uchar4 *__restrict__ pbo, const float *__restrict__ left,
                const float *__restrict__ right) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = row_begin + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= row_end) return;
    const size_t idx = x + (size_t)y * W;
    const size_t li = x + (size_t)y * W * 2;
    pbo[li] = make_uchar4(to_u8(left[3 * idx]), to_u8(left[3 * idx + 1]), to_u8(left[3 * idx + 2]), 0);
    pbo[li + W] = make_uchar4(to_u8(right[3 * idx]), to_u8(right[3 * idx + 1]), to_u8(right[3 * idx + 2]), 0);
}

__global__ void __launch_bounds__(256)
debug_view_kernel(size_t begin, size_t end, int option, const int *__restrict__ hlen, const float4 *__restrict__ cv,
                  float *__restrict__ out) {
    const size_t p = begin + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (p >= end) return;
    const float v = option == 1 ? (float)hlen[p] / 100.0f : cv[p].w / 0.1f;
    out[3 * p] = v; out[3 * p + 1] = v; out[3 * p + 2] = v;
}

__global__ void __launch_bounds__(256)
cv_to_outputs_kernel(size_t begin, size_t end, const float4 *__restrict__ cv, float *__restrict__ denoised,
                     float *__restrict__ var_out) {
    const size_t p = begin + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (p >= end) return;
    const float4 c = cv[p];
    denoised[3 * p] = c.x; denoised[3 * p + 1] = c.y; denoised[3 * p + 2] = c.z;
    var_out[p] = c.w;
}

__global__ void __launch_bounds__(256)
aos_to_soa_kernel(size_t n, const svgf_gbuffer_texel *__restrict__ g, float4 *__restrict__ nrm, float4 *__restrict__ pos,
                  float4 *__restrict__ alb, float4 *__restrict__ gnp, float2 *__restrict__ gzl, float kn, float kx) {
    const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float *t = reinterpret_cast<const float *>(g + p);
    nrm[p] = make_float4(t[0], t[1], t[2], t[12]);
    pos[p] = make_float4(t[3], t[4], t[5], 0.f);
    // the last a-trous level multiplies by albedo * ialbedo (denoise.cu:167); fold ialbedo in here
    alb[p] = make_float4(t[6] * t[9], t[7] * t[10], t[8] * t[11], 0.f);
    gnp[p] = make_float4(t[0] * kn, t[3] * kx, t[1] * kn, t[4] * kx);
    gzl[p] = make_float2(t[2] * kn, t[5] * kx);
}

__global__ void __launch_bounds__(256)
soa_to_aos_kernel(size_t n, const float4 *__restrict__ nrm, const float4 *__restrict__ pos, const float4 *__restrict__ alb,
                  svgf_gbuffer_texel *__restrict__ g) {
    const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float4 a = nrm[p], b = pos[p], c = alb[p];
    float *t = reinterpret_cast<float *>(g + p);
    t[0] = a.x; t[1] = a.y; t[2] = a.z; t[3] = b.x; t[4] = b.y; t[5] = b.z; t[6] = c.x; t[7] = c.y; t[8] = c.z;
    t[9] = 1.0f; t[10] = 1.0f; t[11] = 1.0f;        // ialbedo == 1, pathtrace.cu:323
    t[12] = a.w;
}

__global__ void __launch_bounds__(256)
copy_f3_kernel(size_t begin, size_t end, float *__restrict__ dst, const float *__restrict__ src) {
    const size_t i = begin + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < end) dst[i] = src[i];
}

// ---- cross-rank ordering: sequence flags pushed into the peers' memory, polled locally (halo_sync.cuh). These two kernels
// are the stand-alone forms, used where the producer/consumer kernel of a stage has no fused signal/wait (frame boundary, the
// rarely used paths); the hot stages raise their flags from the producer's last block and poll in the consumer's edge blocks.
struct FlagList { int n; unsigned *flag[SVGF_MAX_RANKS]; };
__global__ void signal_kernel(const __grid_constant__ FlagList l, unsigned seq) {
    const int j = threadIdx.x;
    if (j >= l.n) return;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned *>(l.flag[j]) = seq;
}
__global__ void wait_kernel(const __grid_constant__ FlagList l, unsigned seq, unsigned *err) {
    const int j = threadIdx.x;
    if (j >= l.n) return;
    const volatile unsigned *f = l.flag[j];
    const long long t0 = clock64();
    while ((int)(*f - seq) < 0) {
        if (clock64() - t0 > 4000000000LL) {        // ~2 s: a peer died; record it instead of hanging the GPU
            *reinterpret_cast<volatile unsigned *>(err) = 1u;
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

// ---- halo push: whole rows of a row-major plane are contiguous, so "my rows that rank r will tap" is one memcpy-shaped
// segment per (plane, peer). One launch moves all segments of a stage (blockIdx.y = segment) with 16-byte stores that
// travel over NVLink as full packets; the stage's sequence flag follows in stream order (signal_kernel).
struct PushSeg { const void *src; void *dst; unsigned long long bytes; };
enum { SVGF_PUSH_MAXSEG = 4 * SVGF_MAX_RANKS };
struct PushArgs { PushSeg seg[SVGF_PUSH_MAXSEG]; };

template <class V>
__global__ void __launch_bounds__(256)
halo_push_kernel(const __grid_constant__ PushArgs a, const __grid_constant__ HaloOut sig) {
    const PushSeg &s = a.seg[blockIdx.y];
    const V *__restrict__ src = static_cast<const V *>(s.src);
    V *__restrict__ dst = static_cast<V *>(s.dst);
    const size_t n = s.bytes / sizeof(V), stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
    halo_block_done(sig, true);         // sig.peers.n == 0: no flag follows from here
}

inline dim3 grid2d(int W, int rows, dim3 b) { return dim3((W + b.x - 1) / b.x, (rows + b.y - 1) / b.y); }

}  // namespace

void preload_denoise_kernels() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, temporal_kernel<true>); cudaFuncGetAttributes(&a, temporal_kernel<false>); cudaFuncGetAttributes(&a, no_temporal_kernel); cudaFuncGetAttributes(&a, spatial_variance_kernel);
    cudaFuncGetAttributes(&a, pack_pbo_kernel); cudaFuncGetAttributes(&a, debug_view_kernel); cudaFuncGetAttributes(&a, cv_to_outputs_kernel);
    cudaFuncGetAttributes(&a, aos_to_soa_kernel); cudaFuncGetAttributes(&a, soa_to_aos_kernel); cudaFuncGetAttributes(&a, copy_f3_kernel);
    cudaFuncGetAttributes(&a, signal_kernel); cudaFuncGetAttributes(&a, wait_kernel);
    cudaFuncGetAttributes(&a, halo_push_kernel<uint4>); cudaFuncGetAttributes(&a, halo_push_kernel<uint2>);
    (void)cudaGetLastError();
}

// Ranks whose strips lie within `reach` rows of this rank's strip, with the rows of MY strip each of them taps.
HaloPeers halo_peers(const svgf_ctx *c, int reach) {
    HaloPeers h; h.n = 0;
    const int b = c->shard.row_begin, e = c->shard.row_end;
    if (c->rows.world <= 1 || reach <= 0 || b >= e) return h;
    for (int r = 0; r < c->rows.world; r++) {
        if (r == c->shard.rank || c->rows.start[r] >= c->rows.start[r + 1]) continue;
        const int lo = std::max(b, c->rows.start[r] - reach), hi = std::min(e, c->rows.start[r + 1] + reach);
        if (lo >= hi) continue;
        h.rank[h.n] = r; h.lo[h.n] = lo; h.hi[h.n] = hi; h.n++;
    }
    return h;
}

// Producer side of `stage` for the peers in `reach`: where their copies of my flag live, the block counter, this frame's seq.
HaloOut halo_out(svgf_ctx *c, int stage, int reach, bool fused_signal) {
    HaloOut o; memset(&o, 0, sizeof(o));
    o.peers = halo_peers(c, reach);
    for (int i = 0; i < o.peers.n; i++) o.flag[i] = c->p_flags.p[o.peers.rank[i]] + c->shard.rank * SVGF_NUM_STAGES + stage;
    o.counter = c->done_count + stage; o.seq = c->seq; o.signal = fused_signal ? 1 : 0;
    return o;
}

// Consumer side: my local copies of the flags of `stage` of the peers in `reach`.
HaloIn halo_in(svgf_ctx *c, int stage, int reach, unsigned seq) {
    HaloIn w; memset(&w, 0, sizeof(w));
    const HaloPeers p = halo_peers(c, reach);
    w.n = p.n;
    for (int i = 0; i < p.n; i++) w.flag[i] = c->flags + p.rank[i] * SVGF_NUM_STAGES + stage;
    w.seq = seq; w.err = c->comm_err_dev;
    return w;
}

cudaError_t launch_temporal(svgf_ctx *c, const float *image, const float4 *nrm_cur, const PeerPtr<float4> &nrm_prev,
                            const float4 *pos, const PeerPtr<float4> &hist_cv, const PeerPtr<float2> &mom_hist,
                            const PeerPtr<int> &hlen_in, float4 *acc_cv, float2 *acc_lv, float2 *mom_acc, int *hlen_out,
                            const float *prev_viewmat, float color_alpha, float moment_alpha, float clip_rx, float clip_ry,
                            const HaloOut &ho, const PeerPtr<float4> &acc_cv_peers, const PeerPtr<float2> &acc_lv_peers) {
    const int rows = c->shard.row_end - c->shard.row_begin;
    if (rows <= 0) return cudaSuccess;
    Mat4 vm; for (int i = 0; i < 16; i++) vm.m[i] = prev_viewmat[i];
    TemporalPush push; memset(&push, 0, sizeof(push));
    push.ho = ho;
    for (int i = 0; i < ho.peers.n; i++) { push.cv[i] = acc_cv_peers.p[ho.peers.rank[i]]; push.lv[i] = acc_lv_peers.p[ho.peers.rank[i]]; }
    dim3 b(32, 8);
    // one strip = the whole frame (every table entry is this context's own plane, entry 0 as good as any: a strip set with
    // svgf_set_shard on an unconnected context may carry any rank number): no owner lookups (SVGF_TEMPORAL_SINGLE=0: A/B)
    static const bool single_ok = !(getenv("SVGF_TEMPORAL_SINGLE") && atoi(getenv("SVGF_TEMPORAL_SINGLE")) == 0);
    auto kern = (c->rows.world == 1 && single_ok) ? temporal_kernel<true> : temporal_kernel<false>;
    kern<<<grid2d(c->W, rows, b), b, 0, c->stream>>>(c->W, c->H, c->shard.row_begin, c->shard.row_end, image, nrm_cur,
                                                      nrm_prev, pos, hist_cv, mom_hist, hlen_in, c->rows, c->rows.world == 1 ? 0 : c->shard.rank, acc_cv, acc_lv, mom_acc,
                                                      hlen_out, vm, color_alpha, moment_alpha, clip_rx, clip_ry, c->opt_history_cap, push);
    return cudaGetLastError();
}

cudaError_t launch_spatial_variance(svgf_ctx *c, const int *hlen, const float2 *mom, const float4 *nrm, float4 *cv, float2 *lv) {
    dim3 b(32, 8);
    spatial_variance_kernel<<<grid2d(c->W, c->H, b), b, 0, c->stream>>>(c->W, c->H, hlen, mom, nrm, cv, lv);
    return cudaGetLastError();
}

cudaError_t launch_no_temporal(svgf_ctx *c, const float *image, float4 *acc_cv, float2 *acc_lv) {
    const int rows = c->shard.row_end - c->shard.row_begin;
    if (rows <= 0) return cudaSuccess;
    dim3 b(32, 8);
    no_temporal_kernel<<<grid2d(c->W, rows, b), b, 0, c->stream>>>(c->W, c->shard.row_begin, c->shard.row_end, image, acc_cv, acc_lv);
    return cudaGetLastError();
}

cudaError_t launch_pack_pbo(svgf_ctx *c, unsigned char *pbo, const float *left, const float *right) {
    const int rows = c->shard.row_end - c->shard.row_begin;
    if (rows <= 0) return cudaSuccess;
    dim3 b(32, 8);
    pack_pbo_kernel<<<grid2d(c->W, rows, b), b, 0, c->stream>>>(c->W, c->shard.row_begin, c->shard.row_end,
                                                                 reinterpret_cast<uchar4 *>(pbo), left, right);
    return cudaGetLastError();
}

// Copies this rank's rows that lie within `halo_rows` of another rank's strip into that rank's copy of the plane(s): the
// stand-alone form of the dual stores, for producers that do not push by themselves.
// `sig` (optional): the stage flag the kernel's last block raises in the peers' memory once all rows are on their way.
cudaError_t launch_halo_push(svgf_ctx *c, int halo_rows, const HaloPlane *planes, int nplanes, const HaloOut *sig) {
    const HaloPeers hp = halo_peers(c, halo_rows);
    if (hp.n == 0) return cudaSuccess;
    HaloOut so; memset(&so, 0, sizeof(so));
    if (sig) so = *sig;
    PushArgs a; int nseg = 0; unsigned long long longest = 0; bool v16 = true;
    for (int i = 0; i < hp.n; i++) {
        const int r = hp.rank[i];
        for (int p = 0; p < nplanes; p++) {
            if (!planes[p].peer[r] || planes[p].peer[r] == planes[p].local) continue;     // not connected (yet): own plane
            if (nseg == SVGF_PUSH_MAXSEG) return cudaErrorInvalidValue;
            const size_t off = (size_t)hp.lo[i] * c->W * planes[p].esz, bytes = (size_t)(hp.hi[i] - hp.lo[i]) * c->W * planes[p].esz;
            a.seg[nseg].src = static_cast<const char *>(planes[p].local) + off;
            a.seg[nseg].dst = static_cast<char *>(planes[p].peer[r]) + off;
            a.seg[nseg].bytes = bytes;
            v16 = v16 && (off % 16 == 0) && (bytes % 16 == 0);
            longest = std::max<unsigned long long>(longest, bytes);
            nseg++;
        }
    }
    if (nseg == 0) {        // nothing to copy (planes not connected): the flag alone
        if (so.peers.n > 0 && so.signal) {
            FlagList l; l.n = 0;
            for (int i = 0; i < so.peers.n; i++) l.flag[l.n++] = so.flag[i];
            signal_kernel<<<1, 32, 0, c->stream>>>(l, so.seq);
        }
        return cudaGetLastError();
    }
    const unsigned long long per_block = 256ull * (v16 ? 16 : 8) * 4;        // ~4 vectors per thread
    const unsigned gx = (unsigned)std::min<unsigned long long>((longest + per_block - 1) / per_block, 1024ull);
    dim3 g(gx ? gx : 1, nseg);
    if (v16) halo_push_kernel<uint4><<<g, 256, 0, c->stream>>>(a, so);
    else halo_push_kernel<uint2><<<g, 256, 0, c->stream>>>(a, so);
    return cudaGetLastError();
}

// reach < 0: every connected rank (frame boundary: the temporal pass reads history in place from any strip)
cudaError_t launch_signal(svgf_ctx *c, int stage, int reach) {
    if (c->rows.world <= 1) return cudaSuccess;
    FlagList l; l.n = 0;
    if (reach < 0) { for (int r = 0; r < c->rows.world; r++) l.flag[l.n++] = c->p_flags.p[r] + c->shard.rank * SVGF_NUM_STAGES + stage; }
    else { const HaloOut o = halo_out(c, stage, reach, false); for (int i = 0; i < o.peers.n; i++) l.flag[l.n++] = o.flag[i]; }
    if (l.n == 0) return cudaSuccess;
    signal_kernel<<<1, 32, 0, c->stream>>>(l, c->seq);
    return cudaGetLastError();
}
cudaError_t launch_wait(svgf_ctx *c, int stage, unsigned seq, int reach) {
    if (c->rows.world <= 1) return cudaSuccess;
    FlagList l; l.n = 0;
    if (reach < 0) { for (int r = 0; r < c->rows.world; r++) l.flag[l.n++] = c->flags + r * SVGF_NUM_STAGES + stage; }
    else { const HaloIn w = halo_in(c, stage, reach, seq); for (int i = 0; i < w.n; i++) l.flag[l.n++] = const_cast<unsigned *>(w.flag[i]); }
    if (l.n == 0) return cudaSuccess;
    wait_kernel<<<1, 32, 0, c->stream>>>(l, seq, c->comm_err_dev);
    return cudaGetLastError();
}

static inline void strip_range(const svgf_ctx *c, size_t &b, size_t &e) {
    b = (size_t)c->shard.row_begin * c->W; e = (size_t)c->shard.row_end * c->W;
}

cudaError_t launch_debug_view(svgf_ctx *c, int option, const int *hlen, const float4 *cv, float *denoised) {
    size_t b, e; strip_range(c, b, e);
    if (e <= b) return cudaSuccess;
    debug_view_kernel<<<(unsigned)((e - b + 255) / 256), 256, 0, c->stream>>>(b, e, option, hlen, cv, denoised);
    return cudaGetLastError();
}

cudaError_t launch_cv_to_outputs(svgf_ctx *c, const float4 *cv, float *denoised, float *var_out) {
    size_t b, e; strip_range(c, b, e);
    if (e <= b) return cudaSuccess;
    cv_to_outputs_kernel<<<(unsigned)((e - b + 255) / 256), 256, 0, c->stream>>>(b, e, cv, denoised, var_out);
    return cudaGetLastError();
}

cudaError_t launch_aos_to_soa(svgf_ctx *c, const svgf_gbuffer_texel *g, float4 *nrm, float4 *pos, float4 *alb, float kn, float kx) {
    aos_to_soa_kernel<<<(unsigned)((c->px + 255) / 256), 256, 0, c->stream>>>(c->px, g, nrm, pos, alb, c->gnp, c->gzl, kn, kx);
    return cudaGetLastError();
}

cudaError_t launch_soa_to_aos(svgf_ctx *c, const float4 *nrm, const float4 *pos, const float4 *alb, svgf_gbuffer_texel *g) {
    soa_to_aos_kernel<<<(unsigned)((c->px + 255) / 256), 256, 0, c->stream>>>(c->px, nrm, pos, alb, g);
    return cudaGetLastError();
}

cudaError_t launch_copy_f3(svgf_ctx *c, float *dst, const float *src) {
    size_t b, e; strip_range(c, b, e);
    if (e <= b) return cudaSuccess;
    copy_f3_kernel<<<(unsigned)((3 * (e - b) + 255) / 256), 256, 0, c->stream>>>(3 * b, 3 * e, dst, src);
    return cudaGetLastError();
}
