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
#include "halo_sync.cuh"
#include "atrous_pair_core.h"
#include "atrous_tile_core.h"
#include "atrous_slide_core.h"

#include <cuda.h>            // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda/barrier>
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace {

__device__ __forceinline__ float lum_ref(float r, float g, float b) {
    return (float)(0.2126 * r + 0.7152 * g + 0.0722 * b);      // fp64, denoise.cu:121
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct AtrousK {
    const float4 *cv_in; float4 *cv_out;
    const float2 *lv_in; float2 *lv_out;        // {luminance (fp64 formula), variance} per pixel
    const float4 *nrm, *pos, *alb;
    const float4 *gnp; const float2 *gzl;       // pre-scaled, interleaved G-buffer view (svgf_internal.h)
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
                const float dn = fmaxf(sqrtf(dnx * dnx + dny * dny + dnz * dnz), 0.0f);     // NaN -> weight 1, see dist_of()
                const float dx = fmaxf(sqrtf(dpx * dpx + dpy * dpy + dpz * dpz), 0.0f);
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
    if (k.cv_out) { k.cv_out[p] = o; k.lv_out[p] = make_float2(lum_ref(o.x, o.y, o.z), o.w); }
}

// ---------------------------------------------------------------------------------------------------------------
// Lattice-tiled kernel.
//
// With step s the 5x5 dilated stencil only couples pixels of one residue class (x mod s, y mod s): on that
// sub-lattice it is a DENSE 5x5 stencil. A block therefore owns a tile of LX x LY lattice points (x C = 2 adjacent
// columns, so every global access is a full 32-byte sector) of one class, stages tile + 2-point apron once into
// shared memory -- the same code and the same apron cost (1.4x) for step 2 and step 128 -- and every thread
// computes a 2 x 4 patch of lattice points from a 6 x 8 window of taps held one at a time in registers, so each tap
// is read from shared memory once per thread and reused for up to 8 centres (4.2 pair evaluations per 48-byte read).
//   per tap in shared memory: {r,g,b,var} {kn nx, kx px, kn ny, kx py} {kn nz, kx pz} {lum, var}. Normals/positions are
//   pre-scaled by log2(e)/(sigma+1e-6), so the edge-stopping exponent is |lq-lp|*kl + |n'q-n'p| + |p'q-p'p|
//   (2 sqrt.approx + 1 ex2.approx per pair). Out-of-image taps carry lum = 3e38: their exponent overflows and
//   ex2(-inf) = 0 removes them without a branch.
// Staging is done by the TMA: the strided lattice of one residue class is a plain 5-D box of the row-major plane --
//   dims {floats of an s-pixel cell, column parity, column pair, row within cell, lattice row} -- so one thread issues
//   8 cp.async.bulk.tensor copies (4 planes x 2 column parities) per block and nobody computes a single tap address.
//   The two parities land in separate halves so that a quarter-warp reads 8 consecutive float4 (lane = (ap, c));
//   out-of-range coordinates are zero-filled by the hardware and only border tiles run a fix-up pass (lum = 3e38).
//   Tiles that need rows owned by another GPU (sharded frames) fall back to per-thread cp.async from the owner's memory.
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

struct AtrousT {
    AtrousK k;
    RowOwner ro;                // multi-GPU: which rank owns a row, and that rank's base pointers
    PeerPtr<const float4> p_cv; PeerPtr<const float2> p_lv; PeerPtr<const float4> p_gnp; PeerPtr<const float2> p_gzl;
    int me;
    int b_first;        // first lattice row index covered by the grid (row_begin / step)
    int ncg;            // column groups per class row: step / C
    int use_tma;
#ifdef SVGF_ATROUS_PROBES
    int probe;          // timing probes (make PROBES=1 + SVGF_ATROUS_PROBE; not compiled into the product): 1 = skip the tile load, 2 = skip the arithmetic
#endif
    const float *kl;    // per-pixel luminance-weight scale from atrous_kl_kernel
    // sharded frames: rows of this level's output that the neighbours tap at the NEXT level go into their planes as well, and
    // the last block to finish raises this level's flag in their memory (halo_sync.cuh)
    HaloOut ho;
    float4 *cv_peer[SVGF_MAX_RANKS - 1]; float2 *lv_peer[SVGF_MAX_RANKS - 1];
    alignas(64) CUtensorMap tm_cv, tm_np, tm_zl, tm_lv;
};

// Pre-pass of a level: per-pixel luminance-weight scale  kl = log2(e) / (sqrt(max(blur3x3(variance), 0)) * sigma_l + 1e-6)
// (denoise.cu:100-118,143). The 3x3 Gaussian lives in PIXEL space, i.e. across residue classes, so it is done here where
// it is coalesced instead of per lattice point inside the tiled kernel. 8 B read (L1-shared) + 4 B written per pixel.
// Programmatic dependent launch (at_launch_kernel): let the next kernel of the stream start launching, then wait until the
// previous one has completed and its stores are visible. Both are no-ops for a kernel launched the ordinary way.
__device__ __forceinline__ void at_pdl_sync() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// COHERENT: loads that must see what OTHER blocks of the same launch stored (the fused stage kernel): L2 only, never L1.
template <bool COHERENT> __device__ __forceinline__ float4 at_ld4(const float4 *p) { return COHERENT ? __ldcg(p) : __ldg(p); }
template <bool COHERENT> __device__ __forceinline__ float at_ld1(const float *p) { return COHERENT ? __ldcg(p) : __ldg(p); }

// one thread = 4 consecutive pixels x0 .. x0 + 3 of row y: 3 rows x (left neighbour, 4 pixels, right neighbour)
template <bool COHERENT>
__device__ __forceinline__ void at_kl_job(const PeerPtr<const float2> &lv, const RowOwner &ro, float *__restrict__ kl, int W, int H,
                                          int blur_variance, float sigma_c, int x0, int y) {
    float v[3][6];
    bool rok[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const int ly = y + r - 1;
        rok[r] = ly >= 0 && ly < H && (blur_variance || r == 1);
#pragma unroll
        for (int i = 0; i < 6; i++) v[r][i] = 0.f;
        if (rok[r]) {
            const float2 *row = lv.p[owner_of(ro, ly)] + (size_t)ly * W;
            if (((W & 1) == 0) && x0 + 3 < W) {
                const float4 m0 = at_ld4<COHERENT>(reinterpret_cast<const float4 *>(row + x0)), m1 = at_ld4<COHERENT>(reinterpret_cast<const float4 *>(row + x0 + 2));
                v[r][1] = m0.y; v[r][2] = m0.w; v[r][3] = m1.y; v[r][4] = m1.w;
            } else for (int i = 0; i < 4; i++) if (x0 + i < W) v[r][1 + i] = at_ld1<COHERENT>(&row[x0 + i].y);
            if (x0 > 0) v[r][0] = at_ld1<COHERENT>(&row[x0 - 1].y);
            if (x0 + 4 < W) v[r][5] = at_ld1<COHERENT>(&row[x0 + 4].y);
        }
    }
    float out[4];
    // interior threads (all 18 taps inside the image): the weights sum to exactly 1, so the renormalising divide is an
    // identity and the bounds tests fold away; same products, same order, same bits
    const bool interior = blur_variance && rok[0] && rok[2] && x0 > 0 && x0 + 4 < W;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int x = x0 + i;
        float var;
        if (interior) {
            float sum = 0.0f;
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int dx = -1; dx <= 1; dx++) sum += ((dx == 0 ? 0.5f : 0.25f) * (r == 1 ? 0.5f : 0.25f)) * v[r][1 + i + dx];
            var = sum;
        } else if (blur_variance) {        // same tap order as the reference: rows outer, columns inner (denoise.cu:105-114)
            float sum = 0.0f, sumw = 0.0f;
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int dx = -1; dx <= 1; dx++) {
                    const int lx = x + dx;
                    if (rok[r] && lx >= 0 && lx < W) {
                        const float g = (dx == 0 ? 0.5f : 0.25f) * (r == 1 ? 0.5f : 0.25f);
                        sum += g * v[r][1 + i + dx];
                        sumw += g;
                    }
                }
            var = sum / sumw;
        } else {
            var = v[1][1 + i];
        }
        var = fmaxf(var, 0.0f);
        // fp32: the reference's fp64 add/divide here (denoise.cu:143) only has to be matched to ~1e-7 relative
        // (sqrt.approx and an approximate divide: 2-3 ulp on kl, i.e. < 1e-6 relative on any exponent that matters)
        out[i] = __fdividef(1.4426950408889634f, fmaf(sqrt_approx(var), sigma_c, 1e-6f));
    }
    float *dst = kl + (size_t)y * W + x0;
    if (((W & 3) == 0) && x0 + 3 < W) *reinterpret_cast<float4 *>(dst) = make_float4(out[0], out[1], out[2], out[3]);
    else for (int i = 0; i < 4; i++) if (x0 + i < W) dst[i] = out[i];
}

__global__ void __launch_bounds__(256)
atrous_kl_kernel(const __grid_constant__ PeerPtr<const float2> lv, const __grid_constant__ RowOwner ro, float *__restrict__ kl,
                 int W, int H, int row_begin, int row_end, int blur_variance, float sigma_c, const __grid_constant__ HaloIn wait) {
    // Sharded frames: the rows beyond the strip were stored into this GPU's planes by the neighbours' producers. Only the
    // first and the last block row read such rows here (+-1 row); they wait for the neighbours' flags -- for every neighbour in
    // reach of the level, so that the tile kernel, which starts after this grid has drained, finds its whole apron in place.
    // Interior blocks start at once; the waiting blocks are a few dozen, so they cannot keep a producer off the SMs.
    at_pdl_sync();
    if (wait.n > 0 && (blockIdx.y == 0 || blockIdx.y == gridDim.y - 1)) {
        if (threadIdx.x == 0 && threadIdx.y == 0) halo_wait(wait);
        __syncthreads();
    }
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = row_begin + blockIdx.y * blockDim.y + threadIdx.y;
    if (x0 >= W || y >= row_end) return;
    at_kl_job<false>(lv, ro, kl, W, H, blur_variance, sigma_c, x0, y);
}

namespace cde = cuda::device::experimental;

using barrier_t = cuda::barrier<cuda::thread_scope_block>;

// Stage tile + apron of one residue class into shared memory (shape SH: AtShape or PairShape). Returns true when the TMA path
// was taken (the data has landed), false when cp.async copies are still in flight (caller: cp.async.wait_group + barrier).
struct AtNoMid { __device__ __forceinline__ void operator()() const {} };
// `mid`: work to do while the copies are in flight (after this thread's arrival at the barrier, before its wait).
template <class SH, bool INIT_BAR = true, class MID = AtNoMid>
__device__ __forceinline__ bool at_stage_tile(const AtrousT &t, float4 *s_cv, float4 *s_np, float2 *s_zl, float2 *s_lv, barrier_t &bar,
                                              int X0, int a0, int b0, int yc, int tid, MID mid = MID()) {
    constexpr int AT_SW = SH::SW, AT_SH = SH::SH, AT_THREADS = SH::THREADS, TILE = SH::TILE, HALF = SH::HALF;
    const AtrousK &k = t.k;
    const int W = k.W, H = k.H, step = k.step;
    // rows a live centre of this strip can reach (strip +- 2 steps); a coarse tile spans far more rows than the strip
    const int y_lo = max(0, k.row_begin - 2 * step), y_hi = min(H, k.row_end + 2 * step);
    // does the tile need rows owned by another rank? (first/last needed row of this tile)
    const int yt0 = max(y_lo, yc + max(b0, 0) * step), yt1 = min(y_hi - 1, yc + (b0 + AT_SH - 1) * step);
    const bool remote = t.ro.world > 1 && yt0 <= yt1 && (yt0 < t.ro.start[t.me] || yt1 >= t.ro.start[t.me + 1]);
    const bool tma = t.use_tma && !remote;

#ifdef SVGF_ATROUS_PROBES
    if (t.probe == 1) return tma;       // timing probe: arithmetic on whatever shared memory holds
#endif
    if (tma) {
        if (INIT_BAR) {
            if (tid == 0) {
                init(&bar, AT_THREADS);
                cde::fence_proxy_async_shared_cta();
            }
            __syncthreads();
        } else if (tid == 0) {
            // persistent block: the barrier lives on; shared memory was last touched through the generic proxy (reads, border fix-up)
            cde::fence_proxy_async_shared_cta();
        }
        barrier_t::arrival_token token;
        if (tid == 0) {
#pragma unroll
            for (int par = 0; par < 2; par++) {     // a0 is even: window column ta has parity ta & 1 and pair index a0/2 + ta/2
                cde::cp_async_bulk_tensor_5d_global_to_shared(s_cv + par * SH::HALFP, &t.tm_cv, 4 * X0, par, a0 >> 1, yc, b0, bar);
                cde::cp_async_bulk_tensor_5d_global_to_shared(s_np + par * SH::HALFP, &t.tm_np, 4 * X0, par, a0 >> 1, yc, b0, bar);
                cde::cp_async_bulk_tensor_5d_global_to_shared(s_zl + par * SH::HALFP, &t.tm_zl, 2 * X0, par, a0 >> 1, yc, b0, bar);
                cde::cp_async_bulk_tensor_5d_global_to_shared(s_lv + par * SH::HALFP, &t.tm_lv, 2 * X0, par, a0 >> 1, yc, b0, bar);
            }
            token = cuda::device::barrier_arrive_tx(bar, 1, TILE * 48);
        } else {
            token = bar.arrive();
        }
        mid();
        bar.wait(std::move(token));
        // border tiles: the hardware zero-filled what lies outside the tensor; taps outside the IMAGE (columns >= W inside
        // the rounded-up lattice, padded rows >= H) get zeros too, and every invalid tap gets lum = 3e38
        const bool border = a0 < 0 || X0 + (a0 + AT_SW - 1) * step + AT_C > W || b0 < 0 || yc + (b0 + AT_SH - 1) * step >= H;
        if (border) {
            for (int n = tid; n < TILE; n += AT_THREADS) {
                const int c = n % AT_C, ta = (n / AT_C) % AT_SW, tb = n / (AT_C * AT_SW);
                const int x = X0 + (a0 + ta) * step + c, y = yc + (b0 + tb) * step;
                if (!(a0 + ta >= 0 && b0 + tb >= 0 && x < W && y < H)) {
                    const int si = SH::idx(c, tb, ta);
                    s_cv[si] = make_float4(0.f, 0.f, 0.f, 0.f); s_np[si] = make_float4(0.f, 0.f, 0.f, 0.f);
                    s_zl[si] = make_float2(0.f, 0.f); s_lv[si] = make_float2(3e38f, 0.f);
                }
            }
        }
    } else {
        // per-thread cp.async (LDGSTS): rows of other strips come straight from their owner's memory over NVLink.
        // Each of the first ST_STRIDE * 2 * SW threads owns one (lattice column, sub-column) and walks down the rows.
        constexpr int ST_STRIDE = AT_THREADS / (AT_SW * AT_C);
        if (tid < ST_STRIDE * AT_SW * AT_C) {
            const int slot = tid % (AT_SW * AT_C), r0 = tid / (AT_SW * AT_C);
            const int c = slot % AT_C, ta = slot / AT_C;
            const int x = X0 + (a0 + ta) * step + c;
            const bool x_ok = a0 + ta >= 0 && x < W;
#pragma unroll 4
            for (int tb = r0; tb < AT_SH; tb += ST_STRIDE) {
                const int y = yc + (b0 + tb) * step;
                const int si = SH::idx(c, tb, ta);
                if (x_ok && b0 + tb >= 0 && y >= y_lo && y < y_hi) {
                    const int q = x + y * W, o = owner_of(t.ro, y);
                    cp_async16(&s_cv[si], &t.p_cv.p[o][q]);
                    cp_async16(&s_np[si], &t.p_gnp.p[o][q]);
                    cp_async8(&s_zl[si], &t.p_gzl.p[o][q]);
                    cp_async8(&s_lv[si], &t.p_lv.p[o][q]);
                } else {
                    s_cv[si] = make_float4(0.f, 0.f, 0.f, 0.f); s_np[si] = make_float4(0.f, 0.f, 0.f, 0.f);
                    s_zl[si] = make_float2(0.f, 0.f); s_lv[si] = make_float2(3e38f, 0.f);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    return tma;
}

// Normalise the sums of a 2 x TY patch and write them: {colour, variance} + {luminance, variance} for the next level, and on the
// last level the final colour (x albedo) in the reference's vec3 layout + the variance plane.
template <int AT_TY>
__device__ __forceinline__ void at_write_outputs(const AtrousT &t, const AtAcc2 (&A)[AT_TY], int X0, int a0, int b0, int yc, int ap, int bq, int c) {
    const AtrousK &k = t.k;
    const int W = k.W, step = k.step;
    // ---- outputs: all 8 results (incl. the fp64 luminance of the new colour, a long dependent chain) are computed as
    // straight-line code first, then stored under predicates, so the 8 chains overlap ----
    float4 o[AT_TX][AT_TY]; float ol[AT_TX][AT_TY]; int op[AT_TX][AT_TY];
#pragma unroll
    for (int ca = 0; ca < AT_TX; ca++)
#pragma unroll
        for (int cb = 0; cb < AT_TY; cb++) {
            const int x = X0 + (a0 + 2 * ap + ca + 2) * step + c, y = yc + (b0 + AT_TY * bq + cb + 2) * step;
            op[ca][cb] = (x < W && y >= k.row_begin && y < k.row_end) ? x + y * W : -1;
            // weights_sum >= 9/64 always (the centre tap), so the reference's `else` branch (denoise.cu:162-164) is dead
            const AtAcc2 &a = A[cb];
            const float rw = __frcp_rn(ca ? a.w.y : a.w.x);
            o[ca][cb] = make_float4((ca ? a.r.y : a.r.x) * rw, (ca ? a.g.y : a.g.x) * rw, (ca ? a.b.y : a.b.x) * rw,
                                    __fdividef(ca ? a.v.y : a.v.x, ca ? a.w2.y : a.w2.x));
        }
    if (k.is_last && k.addcolor) {
#pragma unroll
        for (int ca = 0; ca < AT_TX; ca++)
#pragma unroll
            for (int cb = 0; cb < AT_TY; cb++) {
                const float4 al = __ldg(&k.alb[max(op[ca][cb], 0)]);
                o[ca][cb].x *= al.x; o[ca][cb].y *= al.y; o[ca][cb].z *= al.z;
            }
    }
    if (k.cv_out) {
#pragma unroll
        for (int ca = 0; ca < AT_TX; ca++)
#pragma unroll
            for (int cb = 0; cb < AT_TY; cb++) ol[ca][cb] = lum_ref(o[ca][cb].x, o[ca][cb].y, o[ca][cb].z);
    }
#pragma unroll
    for (int ca = 0; ca < AT_TX; ca++)
#pragma unroll
        for (int cb = 0; cb < AT_TY; cb++) {
            const int p = op[ca][cb];
            if (p < 0) continue;
            if (k.is_last) {
                float *d = k.denoised_out + 3 * (size_t)p;
                d[0] = o[ca][cb].x; d[1] = o[ca][cb].y; d[2] = o[ca][cb].z;
                k.var_out[p] = o[ca][cb].w;
            }
            if (k.cv_out) {
                const float2 lvo = make_float2(ol[ca][cb], o[ca][cb].w);
                k.cv_out[p] = o[ca][cb]; k.lv_out[p] = lvo;
                for (unsigned m = halo_targets(t.ho.peers, p / W); m; m &= m - 1) {       // rows a neighbour taps at the next level
                    const int i = __ffs(m) - 1;
                    t.cv_peer[i][p] = o[ca][cb]; t.lv_peer[i][p] = lvo;
                }
            }
        }
}

template <int LX, int LY, int AT_TY, int MINB, bool NS = true>
__global__ void __launch_bounds__((AtShape<LX, LY, AT_TY>::THREADS), MINB)
atrous_tiled_kernel(const __grid_constant__ AtrousT t) {
    using SH = AtShape<LX, LY, AT_TY>;
    static_assert(SH::OK, "tile shape");
    constexpr int AT_LX = LX, AT_SW = SH::SW, AT_SH = SH::SH, AT_THREADS = SH::THREADS, TILE = SH::TILE, HALF = SH::HALF;
    extern __shared__ __align__(128) unsigned char at_smem_raw[];
    float4 *s_cv = reinterpret_cast<float4 *>(at_smem_raw), *s_np = s_cv + SH::TILEP;
    float2 *s_zl = reinterpret_cast<float2 *>(s_np + SH::TILEP), *s_lv = s_zl + SH::TILEP;
    barrier_t &bar = *reinterpret_cast<barrier_t *>(at_smem_raw + SH::TILEP * 48);
    const AtrousK &k = t.k;
    const int W = k.W, H = k.H, step = k.step;
    const int cg = blockIdx.x % t.ncg, tile_x = blockIdx.x / t.ncg;
    const int yc = blockIdx.y % step, tile_y = blockIdx.y / step;
    const int X0 = cg * AT_C;
    const int a0 = tile_x * AT_LX - 2, b0 = t.b_first + tile_y * LY - 2;
    const int tid = threadIdx.x;
    at_pdl_sync();

    // does this block store edge rows into a neighbour's planes? (block-uniform; only such blocks fence at system scope)
    const bool pushes = k.cv_out != nullptr && halo_rows_touch(t.ho.peers, yc + (b0 + 2) * step, yc + (b0 + 1 + LY) * step);
    const bool tma = at_stage_tile<SH>(t, s_cv, s_np, s_zl, s_lv, bar, X0, a0, b0, yc, tid);

    const int c = tid & 1, ap = (tid >> 1) % (AT_LX / 2), bq = tid / AT_LX;
    // centres' kl from the pre-pass plane, issued before the barrier
    float c_kl[AT_TX][AT_TY];
    bool live = false;
#pragma unroll
    for (int ca = 0; ca < AT_TX; ca++)
#pragma unroll
        for (int cb = 0; cb < AT_TY; cb++) {
            const int x = X0 + (a0 + 2 * ap + ca + 2) * step + c, y = yc + (b0 + AT_TY * bq + cb + 2) * step;
            const bool ok = x < W && y >= k.row_begin && y < k.row_end;
            live |= ok;
            c_kl[ca][cb] = ok ? __ldg(&t.kl[x + y * W]) : 0.f;
        }
    if (!tma) asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
#ifdef SVGF_ATROUS_PROBES
    if (t.probe == 2) {     // timing probe: tile load + kl only; one store keeps the loads alive
        if (c_kl[0][0] == 123.456f) k.var_out[0] = s_cv[tid].x;
        return;
    }
#endif
    if (live) {
        AtAcc2 A[AT_TY];
        at_thread_compute<SH, NS>(c, ap, bq, s_cv, s_np, s_zl, s_lv, c_kl, A);
        at_write_outputs<AT_TY>(t, A, X0, a0, b0, yc, ap, bq, c);
    }
    // the block's centres lie in pixel rows yc + (b0 + 2 .. b0 + 2 + LY - 1) * step
    halo_block_done(t.ho, pushes);
}

// ---------------------------------------------------------------------------------------------------------------
// The whole a-trous stage of a frame as ONE launch (SVGF_ATROUS_FUSED=1; bit-identical to the level-by-level launches,
// tests/test_gpu_atrous.py). MEASURED SLOWER ON B200 AND THEREFORE NOT THE DEFAULT: C2 650 us against 479 us, 4K 2203 against
// 1582, a 270-row strip of a 4K frame 405 against 295 (profiles/r2_ab_atrous_stage_one_launch.jsonl). The idea: launched level
// by level, a level costs 26 us + 30.5 us/Mpixel (89 us at 1080p, 279 us at 4K, 55 us for the strip): the kl pre-pass (11 us),
// two kernel boundaries, and half a wave of tiles at the end (5.5 waves of 12 us blocks). What it costs instead, from clock64
// sums of thread 0 of every block (tools/stage_timers.py, share of a block's life at C2): tile compute 40 %, tile load 24 %,
// WAITING FOR DEPENDENCIES 23 % -- at every level's start all blocks sit out the first bands' K items, and a band's tiles
// need ALL tiles of the band D further up to have finished their share of its K work, which 1000 queue positions of
// look-ahead (1.4 waves) only just cover -- claim/decode 5.5 %, K items 5 %. A tile costs 12.8 us here and 12.3 us in its own
// launch, so even with every wait removed the stage would come out at ~460 us: the per-level kernels lose little more than the
// kl pre-pass to their launch structure, and a persistent block pays for it again in claim, dependency probe and signal
// (~2 us per item, 40 items per block and frame). Kept as a measured negative result. The levels' work items -- "K": the luminance-weight scale of 8 pixel rows, "T": one
// lattice tile, the very code of atrous_tiled_kernel -- stand in one queue in dependency order, persistent blocks take them in
// order, and an item starts as soon as the items it reads from have finished: a T item of level L needs the K items of its own
// band of rows and the T items of level L - 1 whose rows its taps reach, a K item the T items of level L - 1 around its rows.
// Levels overlap: while the last tiles of level L drain, tiles of level L + 1 further up the frame are already running.
//   * order: per level, group g = {K items of band g} then {T items of band g - D}: a band's kl is queued D bands -- at least a
//     thousand items, more than the blocks in flight -- ahead of its tiles, so the tiles find it finished (with D = 1 every tile
//     sat out the 20 us its band's K items took: the stage ran 3x SLOWER than level by level). Every dependency points to an
//     EARLIER queue position, and blocks fetch in order, so waiting cannot deadlock;
//   * completion counters per (level, band) in global memory: writers __syncthreads + __threadfence + atomicAdd, readers poll
//     (one thread), fence, __syncthreads; data written by other blocks of the same launch is read through L2 only (TMA, ld.cg);
//   * sharded frames: items that read rows beyond the strip, or store edge rows into a neighbour's planes, poll the
//     neighbours' stage flags first (halo_sync.cuh); the block that completes a level's last tile raises this rank's flag.
// Results are bit-identical to the level-by-level launches (same per-thread code, same operands).
enum { ST_MAXBANDS = 192, ST_KJOBS = 512, ST_CTR_PER_LEVEL = 2 * ST_MAXBANDS + 8 };
struct StageLevel {
    int begin;          // queue position of the level's first item
    int bands, ahead;   // bands of the level; D = min(bands, lookahead): how many bands the K items run ahead of the T items
    int items;          // D * nK + bands * nT: {K items of bands 0..D-1} {T items band by band}; a T item of band b also does
                        // its share (one job per thread) of the K work of band b + D while its tile is in flight
    int nK, nT;         // K / T items per group (a K item = ST_KJOBS jobs of 4 pixels each, 4 per thread)
    int gx;             // tiles per lattice row of tiles x column groups (the per-level kernel's gridDim.x)
    int ly;             // lattice rows per tile
    int band0, band_h;  // pixel row of band 0, pixel rows per band (ly * step)
    int shape;          // 0: 16 x 16 tiles, 1: 32 x 8 tiles
    int total_T;        // T items of the level
};
struct StageArgs {
    AtrousT lvl[SVGF_MAX_LEVELS];
    StageLevel sl[SVGF_MAX_LEVELS];
    HaloIn wait[SVGF_MAX_LEVELS];
    int nlevels, total_items;
    unsigned *ctr;      // [0] queue head, then per level {tdone[ST_MAXBANDS], kdone[ST_MAXBANDS], level_done, ...}
    unsigned *err;
};

__device__ __forceinline__ void st_wait_counter(const unsigned *ctr, unsigned target, unsigned *err) {
    const volatile unsigned *p = ctr;
    const long long t0 = clock64();
    while (*p < target) {
        if (clock64() - t0 > 4000000000LL) { if (err) *reinterpret_cast<volatile unsigned *>(err) = 2u; break; }      // a bug, not a hang
        __nanosleep(64);
    }
}
// Bands of level P whose pixel rows meet [ya, yb] (clipped to the strip) must be complete. BLOCK = false: only look (the counters
// are read all together: independent loads, one round trip to L2); BLOCK = true: wait for every one of them.
template <bool BLOCK>
__device__ __forceinline__ bool st_rows_done(const StageArgs &S, int P, int ya, int yb) {
    const StageLevel &sp = S.sl[P];
    const AtrousK &kp = S.lvl[P].k;
    ya = max(ya, kp.row_begin); yb = min(yb, kp.row_end - 1);
    if (ya > yb) return true;
    const unsigned *tdone = S.ctr + 8 + P * ST_CTR_PER_LEVEL;
    const int j0 = (ya - sp.band0) / sp.band_h, j1 = (yb - sp.band0) / sp.band_h;
    bool ok = true;
    for (int j = j0; j <= j1; j++) {
        if (BLOCK) st_wait_counter(tdone + j, (unsigned)sp.nT, S.err);
        else ok &= *reinterpret_cast<const volatile unsigned *>(tdone + j) >= (unsigned)sp.nT;
    }
    return ok;
}
template <bool BLOCK>
__device__ __forceinline__ bool st_halo_done(const HaloIn &w) {
    if (BLOCK) { halo_wait(w); return true; }
    bool ok = true;
    for (int i = 0; i < w.n; i++) ok &= (int)(*reinterpret_cast<const volatile unsigned *>(w.flag[i]) - w.seq) >= 0;
    return ok;
}

// Tile geometry of a T item, shared by the dependency check (scheduler thread) and the tile code (all threads).
template <class SH>
__device__ __forceinline__ void st_tile_rows(const AtrousT &t, int by, int &yc, int &b0, int &y_first, int &y_last, bool &pushes) {
    constexpr int LY = SH::SH - 4;
    const int step = t.k.step;
    yc = by % step; b0 = t.b_first + (by / step) * LY - 2;
    y_first = yc + b0 * step; y_last = yc + (b0 + SH::SH - 1) * step;
    pushes = t.k.cv_out != nullptr && halo_rows_touch(t.ho.peers, yc + (b0 + 2) * step, yc + (b0 + 1 + LY) * step);
}

struct StItem { int kind, L, band, r; };        // kind: -1 end of queue, 1 K item, 2 T item
__device__ __forceinline__ StItem st_decode(const StageArgs &S, int item) {
    StItem d; d.kind = -1; d.L = 0; d.band = 0; d.r = 0;
    if (item >= S.total_items) return d;
    int L = 0;
    while (L + 1 < S.nlevels && item >= S.sl[L + 1].begin) L++;
    const StageLevel &sl = S.sl[L];
    const int pre = sl.ahead * sl.nK;
    int i = item - sl.begin;
    d.L = L;
    if (i < pre) { d.kind = 1; d.band = i / sl.nK; d.r = i % sl.nK; }
    else { i -= pre; d.kind = 2; d.band = i / sl.nT; d.r = i % sl.nT; }
    return d;
}
// What the item reads has been produced? (BLOCK: wait until it has.)
template <bool BLOCK>
__device__ __forceinline__ bool st_deps(const StageArgs &S, const StItem &d) {
    if (d.kind < 0) return true;
    const StageLevel &sl = S.sl[d.L];
    const AtrousT &t = S.lvl[d.L];
    const AtrousK &k = t.k;
    bool ok = true;
    if (d.kind == 1) {
        const int b_lo = max(k.row_begin, sl.band0 + d.band * sl.band_h), b_hi = min(k.row_end, sl.band0 + (d.band + 1) * sl.band_h);
        const int jobs_x = (k.W + 3) >> 2, jobs = jobs_x * max(b_hi - b_lo, 0);
        const int j0 = d.r * ST_KJOBS, j1 = min(jobs, j0 + ST_KJOBS);
        if (j0 < j1) {
            const int r0 = b_lo + j0 / jobs_x, r1 = b_lo + (j1 - 1) / jobs_x + 1;
            if (d.L > 0) ok &= st_rows_done<BLOCK>(S, d.L - 1, r0 - 1, r1);
            if (S.wait[d.L].n > 0 && (r0 - 1 < k.row_begin || r1 >= k.row_end)) ok &= st_halo_done<BLOCK>(S.wait[d.L]);
        }
    } else {
        const int by = d.band * k.step + d.r / sl.gx;
        int yc, b0, y_first, y_last; bool pushes;
        if (sl.shape == 0) st_tile_rows<AtShape<16, 16, 2>>(t, by, yc, b0, y_first, y_last, pushes);
        else st_tile_rows<AtShape<32, 8, 2>>(t, by, yc, b0, y_first, y_last, pushes);
        const unsigned *kdone = S.ctr + 8 + d.L * ST_CTR_PER_LEVEL + ST_MAXBANDS + d.band;
        const unsigned ktarget = (unsigned)(d.band < sl.ahead ? sl.nK : sl.nT);     // who did the band's K work: K items, or the tiles D bands up
        if (BLOCK) st_wait_counter(kdone, ktarget, S.err);
        else ok &= *reinterpret_cast<const volatile unsigned *>(kdone) >= ktarget;
        // the K work this tile carries: rows of band + D (and one row either side of them)
        const bool carries = d.band + sl.ahead < sl.bands;
        const int ka = sl.band0 + (d.band + sl.ahead) * sl.band_h - 1, kb = sl.band0 + (d.band + sl.ahead + 1) * sl.band_h;
        if (d.L > 0) {
            ok &= st_rows_done<BLOCK>(S, d.L - 1, y_first, y_last);
            if (carries) ok &= st_rows_done<BLOCK>(S, d.L - 1, ka, kb);
        }
        if (S.wait[d.L].n > 0 && (pushes || y_first < k.row_begin || y_last >= k.row_end || (carries && (ka < k.row_begin || kb >= k.row_end))))
            ok &= st_halo_done<BLOCK>(S.wait[d.L]);
    }
    if (ok) __threadfence();        // acquire side: the data behind the counters is read after this
    return ok;
}
// The item's outputs are complete (all threads' stores precede a __syncthreads the caller has passed): publish.
__device__ __forceinline__ void st_signal(const StageArgs &S, const StItem &d) {
    if (d.kind < 0) return;
    __threadfence();
    unsigned *base = S.ctr + 8 + d.L * ST_CTR_PER_LEVEL;
    if (d.kind == 1) { atomicAdd(base + ST_MAXBANDS + d.band, 1u); return; }
    atomicAdd(base + d.band, 1u);
    if (d.band + S.sl[d.L].ahead < S.sl[d.L].bands) atomicAdd(base + ST_MAXBANDS + d.band + S.sl[d.L].ahead, 1u);      // its share of that band's K work
    const AtrousT &t = S.lvl[d.L];
    if (atomicAdd(base + 2 * ST_MAXBANDS, 1u) == (unsigned)S.sl[d.L].total_T - 1u && t.ho.peers.n > 0 && t.ho.signal) {
        __threadfence_system();
        for (int i = 0; i < t.ho.peers.n; i++) *reinterpret_cast<volatile unsigned *>(t.ho.flag[i]) = t.ho.seq;
    }
}

#ifdef SVGF_STAGE_TIMERS      // diagnostic build (tools/build_stage_timers.sh): where a persistent block's time goes, clock64 sums of thread 0
#define ST_T(i) if (tid == 0) { const long long now_ = clock64(); s_tm[i] += now_ - t_last; t_last = now_; }
#else
#define ST_T(i)
#endif

// One tile (the per-level kernel's block (bx, by)) by the 128 threads of a persistent block; its dependencies are complete.
template <class SH, bool NS>
__device__ __forceinline__ void st_tile_item(const StageArgs &S, int L, int bx, int by, int band, int ti, unsigned char *smem, barrier_t &bar, int tid
#ifdef SVGF_STAGE_TIMERS
                                             , long long *s_tm, long long &t_last
#endif
                                             ) {
    constexpr int AT_LX = SH::SW - 4, AT_TY = SH::TY, TILE = SH::TILE;
    const AtrousT &t = S.lvl[L];
    const AtrousK &k = t.k;
    float4 *s_cv = reinterpret_cast<float4 *>(smem), *s_np = s_cv + SH::TILEP;
    float2 *s_zl = reinterpret_cast<float2 *>(s_np + SH::TILEP), *s_lv = s_zl + SH::TILEP;
    const int W = k.W, step = k.step;
    const int cg = bx % t.ncg, tile_x = bx / t.ncg;
    const int X0 = cg * AT_C, a0 = tile_x * AT_LX - 2;
    int yc, b0, y_first, y_last; bool pushes;
    st_tile_rows<SH>(t, by, yc, b0, y_first, y_last, pushes);
    if (tid == 0) asm volatile("fence.proxy.async.global;" ::: "memory");     // other blocks' stores (generic proxy) before this thread's TMA reads
    // While the tile is in flight: this tile's share of the K work of the band D bands further down (about one job per thread).
    const StageLevel &sl = S.sl[L];
    const int kband = band + sl.ahead;
    auto k_share = [&]() {
        if (kband >= sl.bands) return;
        const int b_lo = max(k.row_begin, sl.band0 + kband * sl.band_h), b_hi = min(k.row_end, sl.band0 + (kband + 1) * sl.band_h);
        const int jobs_x = (k.W + 3) >> 2, jobs = jobs_x * max(b_hi - b_lo, 0);
        const int per = (jobs + sl.nT - 1) / sl.nT, j0 = ti * per, j1 = min(jobs, j0 + per);
        for (int j = j0 + tid; j < j1; j += 128)
            at_kl_job<true>(t.p_lv, t.ro, const_cast<float *>(t.kl), k.W, k.H, k.blur_variance, k.sigma_c, (j % jobs_x) * 4, b_lo + j / jobs_x);
    };
    at_stage_tile<SH, false>(t, s_cv, s_np, s_zl, s_lv, bar, X0, a0, b0, yc, tid, k_share);

    const int c = tid & 1, ap = (tid >> 1) % (AT_LX / 2), bq = tid / AT_LX;
    float c_kl[AT_TX][AT_TY];
    bool live = false;
#pragma unroll
    for (int ca = 0; ca < AT_TX; ca++)
#pragma unroll
        for (int cb = 0; cb < AT_TY; cb++) {
            const int x = X0 + (a0 + 2 * ap + ca + 2) * step + c, y = yc + (b0 + AT_TY * bq + cb + 2) * step;
            const bool ok = x < W && y >= k.row_begin && y < k.row_end;
            live |= ok;
            c_kl[ca][cb] = ok ? __ldcg(&t.kl[x + y * W]) : 0.f;
        }
    __syncthreads();
    ST_T(3)
    if (live) {
        AtAcc2 A[AT_TY];
        at_thread_compute<SH, NS>(c, ap, bq, s_cv, s_np, s_zl, s_lv, c_kl, A);
        at_write_outputs<AT_TY>(t, A, X0, a0, b0, yc, ap, bq, c);
    }
    if (pushes) __threadfence_system();
    ST_T(4)
}

using StShapeA = AtShape<16, 16, 2>;
using StShapeB = AtShape<32, 8, 2>;
static_assert(StShapeA::THREADS == 128 && StShapeB::THREADS == 128, "both tile shapes of the stage kernel run on 128 threads");
constexpr int ST_SMEM = (StShapeA::SMEM > StShapeB::SMEM ? StShapeA::SMEM : StShapeB::SMEM) + 112;

template <bool NS>
__global__ void __launch_bounds__(128, 5)
atrous_stage_kernel(const __grid_constant__ StageArgs S) {
    extern __shared__ __align__(128) unsigned char at_smem_raw[];
    barrier_t &bar = *reinterpret_cast<barrier_t *>(at_smem_raw + ST_SMEM - 112);
    int *s_slots = reinterpret_cast<int *>(at_smem_raw + ST_SMEM - 96);      // two slots {kind, level, band, r}: what thread 0 decoded
    const int tid = threadIdx.x;
#ifdef SVGF_STAGE_TIMERS
    long long *s_tm = reinterpret_cast<long long *>(at_smem_raw + ST_SMEM - 64);     // 8 accumulators
    long long t_last = clock64();
    if (tid == 0) for (int i = 0; i < 8; i++) s_tm[i] = 0;
#endif
    // Thread 0 claims items (always one ahead, so the atomic's round trip to L2 hides behind an item's work), decodes them and
    // waits for what they read; the other threads wait at the barrier. Items are claimed in queue order and only ever wait for
    // earlier ones, so the early claim cannot deadlock.
    int next = 0;
    if (tid == 0) {
        init(&bar, 128);
        cde::fence_proxy_async_shared_cta();
        next = (int)atomicAdd(S.ctr, 1u);
    }
    for (int it = 0;; it++) {
        int *slot = s_slots + (it & 1) * 4;
        if (tid == 0) {
            const StItem d = st_decode(S, next);
            if (d.kind >= 0) next = (int)atomicAdd(S.ctr, 1u);
            ST_T(0)
            if (!st_deps<false>(S, d)) st_deps<true>(S, d);
            ST_T(1)
            slot[0] = d.kind; slot[1] = d.L; slot[2] = d.band; slot[3] = d.r;
        }
        __syncthreads();
        const int kind = slot[0], L = slot[1], band = slot[2], r = slot[3];
        if (kind < 0) break;
        const StageLevel &sl = S.sl[L];
        const AtrousT &t = S.lvl[L];
        const AtrousK &k = t.k;
        if (kind == 1) {
            // ---- K item: jobs [r * ST_KJOBS, (r + 1) * ST_KJOBS) of the band (a job = 4 pixels of a row) ----
            const int b_lo = max(k.row_begin, sl.band0 + band * sl.band_h), b_hi = min(k.row_end, sl.band0 + (band + 1) * sl.band_h);
            const int jobs_x = (k.W + 3) >> 2, jobs = jobs_x * max(b_hi - b_lo, 0);
            const int j0 = r * ST_KJOBS, j1 = min(jobs, j0 + ST_KJOBS);
#pragma unroll 2
            for (int j = j0 + tid; j < j1; j += 128)
                at_kl_job<true>(t.p_lv, t.ro, const_cast<float *>(t.kl), k.W, k.H, k.blur_variance, k.sigma_c, (j % jobs_x) * 4, b_lo + j / jobs_x);
            __syncthreads();
            ST_T(2)
        } else {
            // ---- T item: tile (bx, by) of the band ----
            const int bx = r % sl.gx, by = band * k.step + r / sl.gx;
#ifdef SVGF_STAGE_TIMERS
            if (sl.shape == 0) st_tile_item<StShapeA, NS>(S, L, bx, by, band, r, at_smem_raw, bar, tid, s_tm, t_last);
            else st_tile_item<StShapeB, NS>(S, L, bx, by, band, r, at_smem_raw, bar, tid, s_tm, t_last);
#else
            if (sl.shape == 0) st_tile_item<StShapeA, NS>(S, L, bx, by, band, r, at_smem_raw, bar, tid);
            else st_tile_item<StShapeB, NS>(S, L, bx, by, band, r, at_smem_raw, bar, tid);
#endif
            __syncthreads();        // all stores issued, all shared-memory reads done: the next item may overwrite the tile
            ST_T(5)
        }
        // publishing the finished item (fence + atomics, ~1.5 us) is another thread's job, so that it runs next to thread 0's
        // claim / dependency check of the next item instead of in front of it
        if (tid == 32) {
            StItem d; d.kind = kind; d.L = L; d.band = band; d.r = r;
            st_signal(S, d);
        }
        ST_T(6)
    }
#ifdef SVGF_STAGE_TIMERS
    if (tid == 0) for (int i = 0; i < 8; i++) atomicAdd(reinterpret_cast<unsigned long long *>(S.ctr + 8 + SVGF_MAX_LEVELS * ST_CTR_PER_LEVEL) + i, (unsigned long long)s_tm[i]);
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// Symmetric ("pair") variant of the tile kernel (SVGF_ATROUS_VARIANT=4): same tiles, same staging, same outputs; the pair
// arithmetic is split in two phases over shared memory so that the two square roots are taken once per UNORDERED pair
// (csrc/atrous_pair_core.h, which the CPU suite also runs through a host emulation, tests/test_atrous_emu.py).
template <int LX, int LY, int PR, int MINB>
__global__ void __launch_bounds__((PairShape<LX, LY, PR>::THREADS), MINB)
atrous_pair_kernel(const __grid_constant__ AtrousT t) {
    using SH = PairShape<LX, LY, PR>;
    static_assert(SH::OK, "tile shape");
    constexpr int TILE = SH::TILE;
    extern __shared__ __align__(128) unsigned char at_smem_raw[];
    float4 *s_cv = reinterpret_cast<float4 *>(at_smem_raw), *s_np = s_cv + TILE;
    float2 *s_zl = reinterpret_cast<float2 *>(s_np + TILE), *s_lv = s_zl + TILE;
    float *s_g = reinterpret_cast<float *>(at_smem_raw + TILE * 48);
    barrier_t &bar = *reinterpret_cast<barrier_t *>(at_smem_raw + TILE * 48 + SH::NOFF * SH::GN * 4);
    const AtrousK &k = t.k;
    const int W = k.W, step = k.step;
    const int cg = blockIdx.x % t.ncg, tile_x = blockIdx.x / t.ncg;
    const int yc = blockIdx.y % step, tile_y = blockIdx.y / step;
    const int X0 = cg * AT_C;
    const int a0 = tile_x * LX - 2, b0 = t.b_first + tile_y * LY - 2;
    const int tid = threadIdx.x;

    const bool tma = at_stage_tile<SH>(t, s_cv, s_np, s_zl, s_lv, bar, X0, a0, b0, yc, tid);

    // Phase-2 patch of this thread (2 x PR centres). The two half-warps of a warp take patches 4 lattice rows apart (80 floats of
    // g = 16 banks), so that their 4-byte reads of g do not meet in a bank; plain order when the tile height does not allow it.
    const int c = tid & 1, ap = (tid >> 1) % (LX / 2);
    int bq = tid / LX;
    constexpr int D = 4 / PR;
    if (LX == 16 && (LY / PR) % (2 * D) == 0) { const int hw = (tid >> 4) & 1, wp = tid >> 5; bq = (wp % D) + D * hw + 2 * D * (wp / D); }
    float c_kl[2][PR];
    bool live = false;
#pragma unroll
    for (int ca = 0; ca < 2; ca++)
#pragma unroll
        for (int cb = 0; cb < PR; cb++) {
            const int x = X0 + (a0 + 2 * ap + ca + 2) * step + c, y = yc + (b0 + PR * bq + cb + 2) * step;
            const bool ok = x < W && y >= k.row_begin && y < k.row_end;
            live |= ok;
            c_kl[ca][cb] = ok ? __ldg(&t.kl[x + y * W]) : 0.f;
        }
    if (!tma) asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ---- phase 1: g = |dn'| + |dp'| - log2 h for the 12 forward offsets of every staged point that can pair with a centre ----
    for (int n = tid; n < SH::ITEMS; n += SH::THREADS) pair_phase1_item<SH>(n, s_np, s_zl, s_g);
    __syncthreads();
    if (live) {
        // ---- phase 2: one ex2 per (centre, tap) ----
        PairAcc A[PR];
        pair_phase2_thread<SH>(c, ap, bq, s_cv, s_lv, s_g, c_kl, A);
        at_write_outputs<PR>(t, A, X0, a0, b0, yc, ap, bq, c);
    }
    halo_block_done(t.ho, true);
}

// ---------------------------------------------------------------------------------------------------------------
// Sliding kernel (atrous_variant 5): the symmetric formulation with the pair distances in registers, csrc/atrous_slide_core.h.
// Every WARP is its own pipeline: it owns a strip of 16 lattice columns x 2 sub-columns of one residue class and walks down a
// band of lattice rows. Rows are staged two at a time into a private ring of SL_GROUPS groups by TMA -- two rows of a residue
// class are the box {2 pixels, 16 cells, 1, 2} of the 4-D view {s-pixel cell | cell | row in band | band} of the row-major plane
// (4 copies per group, issued by one lane, completion on the group's mbarrier) -- five to six rows ahead of their use, so no
// lane waits for a tile and nobody synchronises across warps. Items (class, strip, band) are dealt to the warps round-robin;
// launch_atrous_slide sizes the bands so that the items fill the resident warps once.
constexpr int SL_GROUPS = 4, SL_WARPS = 4;
constexpr int SL_GROUP_BYTES = 2 * SL_ROW * 48;         // cv r0 r1 (2 x 512) | np r0 r1 (2 x 512) | zl r0 r1 (2 x 256) | lv r0 r1 (2 x 256)
constexpr int SL_RING_BYTES = SL_GROUPS * SL_GROUP_BYTES;
constexpr int SL_SMEM = SL_WARPS * (SL_RING_BYTES + SL_GROUPS * 8);

struct AtrousS {
    AtrousK k;
    SlGrid g;
    const float *kl;
    HaloOut ho;
    float4 *cv_peer[SVGF_MAX_RANKS - 1]; float2 *lv_peer[SVGF_MAX_RANKS - 1];
    alignas(64) CUtensorMap tm_cv, tm_np, tm_zl, tm_lv;
};

__device__ __forceinline__ unsigned sl_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sl_mbar_init(unsigned long long *bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sl_smem(bar)) : "memory");
}
__device__ __forceinline__ void sl_mbar_expect(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sl_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sl_mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned done, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(sl_smem(bar)), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 22)) __trap();       // rows that never land are a bug: fail the launch, do not hang the GPU
    } while (!done);
}
__device__ __forceinline__ void sl_tma_rows(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(sl_smem(dst)), "l"(tm), "r"(sl_smem(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

struct SlWarp {                     // what a warp knows about its current item
    unsigned char *ring; unsigned long long *bars;
    int lane, x, xv, need_fix;      // this lane's pixel column, is it inside the image, does the strip touch the image's edge
    int X0, yc, a0, b0;             // b0 = lattice row of staged row 0 of the item
    int ngroups;                    // groups of two staged rows of the item
    unsigned g0;                    // groups staged by this warp before the item (slot and parity bookkeeping)
};

// staged row r of the item: second half of a group for odd r
__device__ __forceinline__ SlRow sl_row(const SlWarp &w, int r) {
    unsigned char *p = w.ring + ((w.g0 + (unsigned)(r >> 1)) % SL_GROUPS) * SL_GROUP_BYTES;
    const int q = r & 1;
    SlRow o;
    o.cv = reinterpret_cast<const float4 *>(p + q * 512); o.np = reinterpret_cast<const float4 *>(p + 1024 + q * 512);
    o.zl = reinterpret_cast<const float2 *>(p + 2048 + q * 256); o.lv = reinterpret_cast<const float2 *>(p + 2560 + q * 256);
    return o;
}

// Start the load of group gi of the item (staged rows 2 gi, 2 gi + 1 = lattice rows b0 + 2 gi, + 1) into its slot.
template <bool TMA>
__device__ __forceinline__ void sl_issue(const AtrousS &t, const SlWarp &w, int gi) {
    const unsigned g = w.g0 + (unsigned)gi;
    unsigned char *p = w.ring + (g % SL_GROUPS) * SL_GROUP_BYTES;
    const int b = w.b0 + 2 * gi;
    if (TMA) {
        if (w.lane == 0) {
            unsigned long long *bar = w.bars + (g % SL_GROUPS);
            sl_mbar_expect(bar, SL_GROUP_BYTES);
            sl_tma_rows(p, &t.tm_cv, 4 * w.X0, w.a0, w.yc, b, bar);
            sl_tma_rows(p + 1024, &t.tm_np, 4 * w.X0, w.a0, w.yc, b, bar);
            sl_tma_rows(p + 2048, &t.tm_zl, 2 * w.X0, w.a0, w.yc, b, bar);
            sl_tma_rows(p + 2560, &t.tm_lv, 2 * w.X0, w.a0, w.yc, b, bar);
        }
    } else {        // odd widths (8-byte planes need a 16-byte pitch for TMA): every lane fetches its own entries
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int y = w.yc + (b + q) * t.k.step;
            float4 cv = make_float4(0.f, 0.f, 0.f, 0.f), np = cv; float2 zl = make_float2(0.f, 0.f), lv = make_float2(3e38f, 0.f);
            if (w.xv && b + q >= 0 && y < t.k.H) {
                const size_t i = (size_t)w.x + (size_t)y * t.k.W;
                cv = __ldg(&t.k.cv_in[i]); np = __ldg(&t.k.gnp[i]); zl = __ldg(&t.k.gzl[i]); lv = __ldg(&t.k.lv_in[i]);
            }
            reinterpret_cast<float4 *>(p + q * 512)[w.lane] = cv; reinterpret_cast<float4 *>(p + 1024 + q * 512)[w.lane] = np;
            reinterpret_cast<float2 *>(p + 2048 + q * 256)[w.lane] = zl; reinterpret_cast<float2 *>(p + 2560 + q * 256)[w.lane] = lv;
        }
    }
}

// Group gi has landed (and columns outside the image carry lum = 3e38, which zeroes every weight of such a tap).
template <bool TMA>
__device__ __forceinline__ void sl_wait(const SlWarp &w, int gi) {
    const unsigned g = w.g0 + (unsigned)gi;
    if (TMA) {
        sl_mbar_wait(w.bars + (g % SL_GROUPS), (g / SL_GROUPS) & 1u);
        if (w.need_fix) {           // cells left of the image are zero-filled, cells right of it alias the next row
            if (!w.xv) {
                float2 *lv = reinterpret_cast<float2 *>(w.ring + (g % SL_GROUPS) * SL_GROUP_BYTES + 2560);
                lv[w.lane].x = 3e38f; lv[SL_ROW + w.lane].x = 3e38f;
            }
            __syncwarp();
        }
    } else {
        __syncwarp();
    }
}

// Sharded frames: rows a neighbour taps at the next level also go into its planes (halo_sync.cuh). Out of line: the single-GPU
// path should not carry this code in its five step bodies.
__device__ __noinline__ void sl_halo_store(const AtrousS &t, bool ok, int yo, int p, float4 c, float2 lvo) {
    if (!ok) return;
    for (unsigned m = halo_targets(t.ho.peers, yo); m; m &= m - 1) {
        const int i = __ffs(m) - 1;
        t.cv_peer[i][p] = c; t.lv_peer[i][p] = lvo;
    }
}

// One step: tap row rt of the item (pixel row y_t) against the five centres in flight.
template <int PHI, bool TMA, bool LAST>
__device__ __forceinline__ void sl_step(const AtrousS &t, const SlWarp &w, SlLane &L, int rt, int rows, int y_t, const int (&e)[5], float kl_enter,
                                        bool interior) {
    constexpr int KE = sl_set(PHI, 2), KX = sl_set(PHI, -2);
    const AtrousK &k = t.k;
    // the centre two rows below the tap row enters flight; its group is new on even steps
    if (!(rt & 1)) sl_wait<TMA>(w, (rt >> 1) + 1);
    sl_enter<KE>(L, sl_row(w, rt + 2), w.lane, kl_enter);
    if (y_t >= 0 && y_t < k.H) {        // tap rows outside the image have no pairs (warp-uniform)
        const SlRow row = sl_row(w, rt);
        // what the lanes owning the neighbouring columns hold for this lane's centres above the tap row (registers written one
        // and two steps ago: the shuffles depend on nothing in this step)
        float r1[5], r2[5], s[2], bk[2];
        r1[2] = sl_offer_r1<PHI>(L, 0); r2[2] = sl_offer_r2<PHI>(L, 0);
        r1[0] = __shfl_up_sync(0xffffffffu, sl_offer_r1<PHI>(L, -2), 4); r2[0] = __shfl_up_sync(0xffffffffu, sl_offer_r2<PHI>(L, -2), 4);
        r1[1] = __shfl_up_sync(0xffffffffu, sl_offer_r1<PHI>(L, -1), 2); r2[1] = __shfl_up_sync(0xffffffffu, sl_offer_r2<PHI>(L, -1), 2);
        r1[3] = __shfl_down_sync(0xffffffffu, sl_offer_r1<PHI>(L, 1), 2); r2[3] = __shfl_down_sync(0xffffffffu, sl_offer_r2<PHI>(L, 1), 2);
        r1[4] = __shfl_down_sync(0xffffffffu, sl_offer_r1<PHI>(L, 2), 4); r2[4] = __shfl_down_sync(0xffffffffu, sl_offer_r2<PHI>(L, 2), 4);
        {   // same-row pairs: to the right computed, to the left received
            const SlTap t3 = sl_load_tap(row, e[3]), t4 = sl_load_tap(row, e[4]);
            sl_same_row<PHI>(L, t3, t4, s);
            bk[0] = __shfl_up_sync(0xffffffffu, s[0], 2);
            bk[1] = __shfl_up_sync(0xffffffffu, s[1], 4);
            sl_tap<PHI, 3>(L, t3, r1[3], r2[3], s[0]);
            sl_tap<PHI, 4>(L, t4, r1[4], r2[4], s[1]);
        }
        sl_tap<PHI, 2>(L, sl_load_tap(row, e[2]), r1[2], r2[2], pair_nlog2h(0, 0));
        sl_tap<PHI, 1>(L, sl_load_tap(row, e[1]), r1[1], r2[1], bk[0]);
        sl_tap<PHI, 0>(L, sl_load_tap(row, e[0]), r1[0], r2[0], bk[1]);
    }
    // the centre two rows above the tap row is complete: straight-line arithmetic, predicated stores
    {
        const int yo = y_t - 2 * k.step;
        const bool ok = rt >= 4 && interior && w.x < k.W && yo >= k.row_begin && yo < k.row_end;
        const SlAccS o = sl_exit<KX>(L);
        // sum w >= h(0,0) always (the centre tap), so the reference's `else` branch (denoise.cu:162-164) is dead
        const float rw = __frcp_rn(o.w);
        float4 c = make_float4(o.r * rw, o.g * rw, o.b * rw, __fdividef(o.v, o.w2));
        const int p = ok ? w.x + yo * k.W : 0;
        if (LAST) {
            if (k.addcolor) { const float4 al = __ldg(&k.alb[p]); c.x *= al.x; c.y *= al.y; c.z *= al.z; }
            if (ok) {
                float *d = k.denoised_out + 3 * (size_t)p;
                d[0] = c.x; d[1] = c.y; d[2] = c.z;
                k.var_out[p] = c.w;
            }
        }
        if (!LAST || k.cv_out) {
            const float2 lvo = make_float2(lum_ref(c.x, c.y, c.z), c.w);
            if (ok) {
                k.cv_out[p] = c; k.lv_out[p] = lvo;
            }
            if (t.ho.peers.n) sl_halo_store(t, ok, yo, p, c, lvo);       // sharded frames only (warp-uniform)
        }
    }
    // both rows of the tap row's group are done after an odd step: its slot takes the group SL_GROUPS further down
    if (rt & 1) {
        __syncwarp();
        if ((rt >> 1) + SL_GROUPS < w.ngroups) sl_issue<TMA>(t, w, (rt >> 1) + SL_GROUPS);
    }
}

template <bool TMA, bool LAST>
__global__ void __launch_bounds__(SL_WARPS * 32, 3)
atrous_slide_kernel(const __grid_constant__ AtrousS t) {
    extern __shared__ __align__(1024) unsigned char sl_smem_raw[];
    const AtrousK &k = t.k;
    SlWarp w;
    // (the shuffle tells the compiler that the warp index -- and with it every item parameter, ring address and TMA coordinate
    // derived from it -- is uniform across the warp, so they live in uniform registers instead of being broadcast lane by lane)
    const int wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    w.lane = threadIdx.x & 31;
    w.ring = sl_smem_raw + wid * SL_RING_BYTES;
    w.bars = reinterpret_cast<unsigned long long *>(sl_smem_raw + SL_WARPS * SL_RING_BYTES) + wid * SL_GROUPS;
    if (TMA) {
        if (w.lane == 0) {
            for (int i = 0; i < SL_GROUPS; i++) sl_mbar_init(w.bars + i);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
    }
    w.g0 = 0;
    int e[5];
#pragma unroll
    for (int ti = 0; ti < 5; ti++) e[ti] = min(max(w.lane + 2 * (ti - 2), 0), SL_ROW - 1);       // edge lanes: any entry, their sums are dropped
    const int items = t.g.items(), a = w.lane >> 1, c = w.lane & 1;
    const bool interior = a >= SL_EDGE && a < SL_COLS - SL_EDGE;
    for (int n = blockIdx.x * SL_WARPS + wid; n < items; n += gridDim.x * SL_WARPS) {
        const SlItem it = t.g.item(n);
        const int rows = it.b_hi - it.b_lo;             // centre rows of the band; staged rows: rows + 6, steps: rows + 4
        if (rows <= 0) continue;
        w.X0 = it.X0; w.yc = it.yc; w.a0 = it.a0; w.b0 = it.b_lo - 2;
        w.x = it.X0 + (it.a0 + a) * k.step + c;
        w.xv = (it.a0 + a >= 0) && (w.x < k.W);
        w.need_fix = (it.a0 < 0) || (it.X0 + (it.a0 + SL_COLS - 1) * k.step + 1 >= k.W);
        w.ngroups = (rows + 7) >> 1;
        for (int gi = 0; gi < SL_GROUPS && gi < w.ngroups; gi++) sl_issue<TMA>(t, w, gi);
        SlLane L;
        memset(&L, 0, sizeof(L));
        // luminance-weight scale of the centres (pre-pass), requested two steps before they enter flight
        int y_t = w.yc + w.b0 * k.step;                 // pixel row of tap row 0
        auto kl_at = [&](int y) -> float { return (w.xv && y >= 0 && y < k.H) ? __ldg(&t.kl[w.x + (size_t)y * k.W]) : 0.f; };
        float q0 = kl_at(y_t + 2 * k.step), q1 = kl_at(y_t + 3 * k.step);
        sl_wait<TMA>(w, 0);
        sl_enter<0>(L, sl_row(w, 0), w.lane, kl_at(y_t));
        sl_enter<1>(L, sl_row(w, 1), w.lane, kl_at(y_t + k.step));
        const int steps = rows + 4;
        for (int rt = 0; rt < steps; rt += 5) {
            // five steps = one turn of the register rotation. The centre entering at step r lies in staged row r + 2; its kl is
            // requested two steps earlier and waits in q0/q1.
            float kq;
            kq = q0; q0 = kl_at(y_t + 4 * k.step); sl_step<0, TMA, LAST>(t, w, L, rt, rows, y_t, e, kq, interior); y_t += k.step;
            if (rt + 1 < steps) { kq = q1; q1 = kl_at(y_t + 4 * k.step); sl_step<1, TMA, LAST>(t, w, L, rt + 1, rows, y_t, e, kq, interior); y_t += k.step; }
            if (rt + 2 < steps) { kq = q0; q0 = kl_at(y_t + 4 * k.step); sl_step<2, TMA, LAST>(t, w, L, rt + 2, rows, y_t, e, kq, interior); y_t += k.step; }
            if (rt + 3 < steps) { kq = q1; q1 = kl_at(y_t + 4 * k.step); sl_step<3, TMA, LAST>(t, w, L, rt + 3, rows, y_t, e, kq, interior); y_t += k.step; }
            if (rt + 4 < steps) { kq = q0; q0 = kl_at(y_t + 4 * k.step); sl_step<4, TMA, LAST>(t, w, L, rt + 4, rows, y_t, e, kq, interior); y_t += k.step; }
            kq = q0; q0 = q1; q1 = kq;
        }
        w.g0 += (unsigned)w.ngroups;
    }
    halo_block_done(t.ho, true);
}

}  // namespace

// log2(e) / (sigma + 1e-6) in fp64 (the reference adds and divides in double, denoise.cu:144-145)
void atrous_scales(float sigma_n, float sigma_x, float *kn, float *kx) {
    const double log2e = 1.4426950408889634;
    *kn = (float)(log2e / ((double)sigma_n + 1e-6));
    *kx = (float)(log2e / ((double)sigma_x + 1e-6));
}

// Programmatic dependent launch (the default; SVGF_PDL=0 for A/B): the kl pre-pass and the tile kernel of the a-trous chain are
// launched with cudaLaunchAttributeProgrammaticStreamSerialization, trigger their dependents as their first instruction
// (griddepcontrol.launch_dependents) and wait for their prerequisite grid right after (griddepcontrol.wait, before the first
// read of anything another kernel wrote). The next kernel's blocks are thus resident, past their launch latency and set-up, when
// the previous grid's last block retires -- instead of the ~2-3 us between two dependent launches of a stream.
static bool g_pdl = !getenv("SVGF_PDL") || atoi(getenv("SVGF_PDL")) != 0;
template <class... KArgs, class... Args>
static cudaError_t at_launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- tile shapes ------------------------------------------------------------------------------------------------------
// One entry per instantiation of the tiled kernel. `warp_rows` = lattice rows one warp covers (work is issued per warp, and a
// warp whose patches all lie outside the strip exits right after the tile has landed).
struct AtShapeInfo {
    int lx, ly, ty, threads, smem, warp_rows;
    const void *fn;
    void (*launch)(dim3 grid, cudaStream_t st, const AtrousT &t);
};
template <int LX, int LY, int TY, int MINB> static void at_launch(dim3 grid, cudaStream_t st, const AtrousT &t) {
    using SH = AtShape<LX, LY, TY>;
    (void)at_launch_kernel(atrous_tiled_kernel<LX, LY, TY, MINB, true>, grid, dim3(SH::THREADS), SH::SMEM, st, t);
}
template <int LX, int LY, int TY, int MINB> static AtShapeInfo at_info() {
    using SH = AtShape<LX, LY, TY>;
    return AtShapeInfo{LX, LY, TY, SH::THREADS, SH::SMEM, (32 / LX > 0 ? 32 / LX : 1) * TY,
                       (const void *)atrous_tiled_kernel<LX, LY, TY, MINB>, at_launch<LX, LY, TY, MINB>};
}
// the two default shapes once more without the NaN guard of the distances (frames whose G-buffer cannot hold a NaN)
template <int LX, int LY, int TY, int MINB> static void at_launch_nonan(dim3 grid, cudaStream_t st, const AtrousT &t) {
    using SH = AtShape<LX, LY, TY>;
    (void)at_launch_kernel(atrous_tiled_kernel<LX, LY, TY, MINB, false>, grid, dim3(SH::THREADS), SH::SMEM, st, t);
}
enum { AT_NSHAPES = 15 };
static const AtShapeInfo g_at_shapes[AT_NSHAPES] = {
    at_info<16, 32, 4, 3>(),    // 0: 128 threads, 69 KB, 3 blocks/SM
    at_info<32, 16, 4, 3>(),    // 1: 128 threads, 69 KB, 3 blocks/SM
    at_info<16, 16, 2, 5>(),    // 2: 128 threads, 38 KB, 5 blocks/SM (20 warps)
    at_info<16, 32, 2, 2>(),    // 3: 256 threads, 69 KB, 2-3 blocks/SM (16-24 warps)
    at_info<32, 16, 2, 2>(),    // 4: 256 threads, 69 KB
    at_info<16, 24, 4, 4>(),    // 5:  96 threads, 54 KB, 4 blocks/SM
    at_info<32, 12, 4, 4>(),    // 6:  96 threads, 55 KB, 4 blocks/SM (34 lattice rows = 3 tiles)
    at_info<16, 32, 2, 3>(),    // 7: 256 threads, 69 KB, 3 blocks/SM (24 warps, 80 registers)
    at_info<32, 8, 2, 5>(),     // 8: 128 threads, 41 KB, 5 blocks/SM (lattices of a few rows: strips of a sharded frame)
    at_info<16, 12, 2, 6>(),    // 9:  96 threads, 31 KB, 6-7 blocks/SM
    at_info<32, 12, 2, 3>(),    // 10: 192 threads, 55 KB, 3 blocks/SM
    at_info<16, 16, 2, 4>(),    // 11: shape 2 at 4 blocks/SM (128 registers: more pairs in flight per warp, fewer warps) -- A/B
    at_info<16, 16, 2, 3>(),    // 12: shape 2 at 3 blocks/SM (168 registers) -- A/B
    at_info<16, 18, 3, 5>(),    // 13: 2 x 3 patches, 96 threads, 43 KB, 5 blocks/SM (15 warps, 128 registers); 540 and 270 lattice rows divide by 18 -- A/B
    at_info<16, 16, 1, 5>(),    // 14: 2 x 1 patches, 256 threads, 38 KB, 5 blocks/SM (40 warps, <= 51 registers) -- A/B
};

// ---- TMA descriptors ------------------------------------------------------------------------------------------------
// The lattice of residue class (., yc) with sub-columns [X0, X0+2) of a row-major plane with `fpp` floats per pixel is
// the 5-D tensor  {j: fpp*s floats of one s-pixel cell | parity of the cell | cell pair | row inside the s-row band | band}
// with byte strides {-, fpp*4*s, fpp*4*2s, fpp*4*W, fpp*4*s*W}; a tile is the box {fpp*2, 1, SW/2, 1, SH}.
// Extents round up, so addresses may run past the image: columns >= W alias the next row (fixed up in the kernel),
// rows >= H fall into SVGF_PAD_ROWS rows of zeroed padding behind every plane.
typedef CUresult (*encode_fn_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
enum { TM_PLANES = 2 * SVGF_NCV + 4, TM_LEVELS = SVGF_MAX_LEVELS + 1, TM_SHAPES = AT_NSHAPES };      // cv x4, lv x4, {gnp, gzl} x 2 sets
static inline CUtensorMap *tmap_at(svgf_ctx *c, int plane, int level, int shape) {
    return static_cast<CUtensorMap *>(c->tmaps) + ((plane * TM_LEVELS + level) * TM_SHAPES + shape);
}

static int atrous_build_slide_maps(svgf_ctx *c, encode_fn_t encode);

int atrous_build_tensor_maps(svgf_ctx *c) {
    c->tma_ok = 0;
    if (c->W & 1) return 0;         // 8 B/pixel planes need a 16-byte row pitch; odd widths use the cp.async loader
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    encode_fn_t encode = reinterpret_cast<encode_fn_t>(fn);
    if (!c->tmaps) c->tmaps = aligned_alloc(64, sizeof(CUtensorMap) * TM_PLANES * TM_LEVELS * TM_SHAPES);
    if (!c->tmaps) return 0;
    // L2 promotion: a tile row is one 32-byte sector per (lattice point, plane); the sectors next to it belong to the column
    // groups that the neighbouring blocks (adjacent blockIdx.x) load at about the same time. SVGF_TMA_L2PROMO=0..3 (A/B runs).
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_NONE;
    if (const char *v = getenv("SVGF_TMA_L2PROMO")) {
        const int q = atoi(v);
        promo = q == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : q == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : q == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : promo;
    }
    void *planes[TM_PLANES]; int fpp[TM_PLANES];
    for (int i = 0; i < SVGF_NCV; i++) { planes[i] = c->cv[i]; fpp[i] = 4; planes[SVGF_NCV + i] = c->lv[i]; fpp[SVGF_NCV + i] = 2; }
    for (int g = 0; g < 2; g++) {
        planes[2 * SVGF_NCV + 2 * g] = c->gnp_set[g]; fpp[2 * SVGF_NCV + 2 * g] = 4;
        planes[2 * SVGF_NCV + 2 * g + 1] = c->gzl_set[g]; fpp[2 * SVGF_NCV + 2 * g + 1] = 2;
    }
    for (int pl = 0; pl < TM_PLANES; pl++)
        for (int level = 1; level <= SVGF_MAX_LEVELS; level++)
            for (int shp = 0; shp < TM_SHAPES; shp++) {
                const int sw = g_at_shapes[shp].lx + 4, sh = g_at_shapes[shp].ly + 4;
                const cuuint64_t s = 1ull << level, W = c->W, H = c->H, f = fpp[pl];
                const cuuint64_t dims[5] = {f * s, 2, (W + 2 * s - 1) / (2 * s), s, (H + s - 1) / s};
                const cuuint64_t strides[4] = {f * 4 * s, f * 4 * 2 * s, f * 4 * W, f * 4 * s * W};
                const cuuint32_t box[5] = {(cuuint32_t)(f * AT_C), 1, (cuuint32_t)(sw / 2), 1, (cuuint32_t)sh};
                const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
                CUresult r = encode(tmap_at(c, pl, level, shp), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, planes[pl], dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) return 0;
            }
    c->tma_ok = atrous_build_slide_maps(c, encode) ? 1 : 0;
    return c->tma_ok;
}

// Tensor maps of the sliding kernel: one per (plane, level), 4-D {floats of an s-pixel cell | cell | row in band | band}; a staged
// pair of rows is the box {2 pixels, SL_COLS cells, 1, 2}. Same rounding-up of the extents (and the same padding behind the planes) as above.
static inline CUtensorMap *tmap_slide(svgf_ctx *c, int plane, int level) {
    return static_cast<CUtensorMap *>(c->tmaps_slide) + (plane * TM_LEVELS + level);
}
static int atrous_build_slide_maps(svgf_ctx *c, encode_fn_t encode) {
    if (!c->tmaps_slide) c->tmaps_slide = aligned_alloc(64, sizeof(CUtensorMap) * TM_PLANES * TM_LEVELS);
    if (!c->tmaps_slide) return 0;
    void *planes[TM_PLANES]; int fpp[TM_PLANES];
    for (int i = 0; i < SVGF_NCV; i++) { planes[i] = c->cv[i]; fpp[i] = 4; planes[SVGF_NCV + i] = c->lv[i]; fpp[SVGF_NCV + i] = 2; }
    planes[2 * SVGF_NCV] = c->gnp_set[0]; fpp[2 * SVGF_NCV] = 4; planes[2 * SVGF_NCV + 1] = c->gzl_set[0]; fpp[2 * SVGF_NCV + 1] = 2;
    for (int pl = 0; pl < 2 * SVGF_NCV + 2; pl++)       // (this A/B variant runs on G-buffer set 0 only: no cross-frame overlap)
        for (int level = 1; level <= SVGF_MAX_LEVELS; level++) {
            const cuuint64_t s = 1ull << level, W = c->W, H = c->H, f = fpp[pl];
            const cuuint64_t dims[4] = {f * s, (W + s - 1) / s, s, (H + s - 1) / s};
            const cuuint64_t strides[3] = {f * 4 * s, f * 4 * W, f * 4 * s * W};
            const cuuint32_t box[4] = {(cuuint32_t)(f * 2), (cuuint32_t)SL_COLS, 1, 2};
            const cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult r = encode(tmap_slide(c, pl, level), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, planes[pl], dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return 0;
        }
    return 1;
}

// Which tile shape for a lattice of lat_w x lat_rows points per residue class. Measured on B200 (tools/ab_atrous.py, us per
// level, cornell): every block stages its whole tile, live or not, and a warp with no live patch exits, so what counts is
// how many blocks are mostly dead and how many warps per SM stay busy.
//   1080p, whole frame (lattice rows 540/270/135/68/34):  shape 2: 88 89 93 101 132 | shape 9: 91 93 94  99 118 | shape 0: 93 98 105 126 173
//   4K, strip of 270 rows  (lattice rows 135/68/34/17/9):  shape 2: 54 54 60  71  82 | shape 9: 55 54 56  62  73 | shape 0: 59 61  68  84  91
// 2 x 2 patches (20 warps/SM) beat 2 x 4 patches (12 warps/SM) at every level; 16 x 16 tiles win while the lattice is tall,
// 16 x 12 tiles (6-7 blocks/SM, 3 tiles for 34 rows) once it is short.
static int at_pick_shape(const svgf_ctx *c, int level, int lat_w, int lat_rows) {
    if (c->atrous_shape >= 0 && c->atrous_shape < AT_NSHAPES) return c->atrous_shape;
    if (c->atrous_shape_level[level] >= 0 && c->atrous_shape_level[level] < AT_NSHAPES) return c->atrous_shape_level[level];
    (void)lat_w;
    return lat_rows >= 100 ? 2 : 9;
}

// SVGF_ATROUS_VARIANT=4: the symmetric pair kernel, 16 x 16 tiles (16 x 12 for short lattices); its tile boxes are those of
// shapes 2 and 9, so it shares their tensor maps.
static cudaError_t launch_atrous_pair(svgf_ctx *c, AtrousT &t, const AtrousArgs &a, int lat_w, int lat_rows) {
    const bool tall = lat_rows >= 100 || c->atrous_shape == 2;
    const int shape = (tall && c->atrous_shape != 9) ? 2 : 9;
    t.use_tma = c->tma_ok && a.src_slot >= 0;
    if (t.use_tma) {
        t.tm_cv = *tmap_at(c, a.src_slot, a.level, shape); t.tm_lv = *tmap_at(c, SVGF_NCV + a.src_slot, a.level, shape);
        t.tm_np = *tmap_at(c, 2 * SVGF_NCV, a.level, shape); t.tm_zl = *tmap_at(c, 2 * SVGF_NCV + 1, a.level, shape);
    } else {
        memset(&t.tm_cv, 0, sizeof(CUtensorMap)); memset(&t.tm_lv, 0, sizeof(CUtensorMap));
        memset(&t.tm_np, 0, sizeof(CUtensorMap)); memset(&t.tm_zl, 0, sizeof(CUtensorMap));
    }
    struct PairInst { const void *fn; int threads, smem; void (*launch)(dim3, cudaStream_t, const AtrousT &); };
    static const PairInst inst[4] = {       // [tall ? 0 : 1][patch rows 2 ? 0 : 1]
        {(const void *)atrous_pair_kernel<16, 16, 2, 3>, PairShape<16, 16, 2>::THREADS, PairShape<16, 16, 2>::SMEM,
         [](dim3 g, cudaStream_t st, const AtrousT &tt) { atrous_pair_kernel<16, 16, 2, 3><<<g, PairShape<16, 16, 2>::THREADS, PairShape<16, 16, 2>::SMEM, st>>>(tt); }},
        {(const void *)atrous_pair_kernel<16, 16, 1, 3>, PairShape<16, 16, 1>::THREADS, PairShape<16, 16, 1>::SMEM,
         [](dim3 g, cudaStream_t st, const AtrousT &tt) { atrous_pair_kernel<16, 16, 1, 3><<<g, PairShape<16, 16, 1>::THREADS, PairShape<16, 16, 1>::SMEM, st>>>(tt); }},
        {(const void *)atrous_pair_kernel<16, 12, 2, 3>, PairShape<16, 12, 2>::THREADS, PairShape<16, 12, 2>::SMEM,
         [](dim3 g, cudaStream_t st, const AtrousT &tt) { atrous_pair_kernel<16, 12, 2, 3><<<g, PairShape<16, 12, 2>::THREADS, PairShape<16, 12, 2>::SMEM, st>>>(tt); }},
        {(const void *)atrous_pair_kernel<16, 12, 1, 3>, PairShape<16, 12, 1>::THREADS, PairShape<16, 12, 1>::SMEM,
         [](dim3 g, cudaStream_t st, const AtrousT &tt) { atrous_pair_kernel<16, 12, 1, 3><<<g, PairShape<16, 12, 1>::THREADS, PairShape<16, 12, 1>::SMEM, st>>>(tt); }},
    };
    if (!c->atrous_pair_attr_set) {
        for (const PairInst &pi : inst) {
            cudaError_t e = cudaFuncSetAttribute(pi.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, pi.smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(pi.fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) return e;
        }
        c->atrous_pair_attr_set = true;
    }
    const int step = t.k.step, ly = shape == 2 ? 16 : 12;
    const dim3 g(((lat_w + 15) / 16) * t.ncg, ((lat_rows + ly - 1) / ly) * step);
    inst[(shape == 2 ? 0 : 2) + (c->atrous_pair_rows == 1 ? 1 : 0)].launch(g, c->stream, t);
    return cudaGetLastError();
}

// SVGF_ATROUS_VARIANT=5: the sliding kernel. Bands are sized so that the items (class x strip x band) fill the resident warps
// about once: a warp's item is a long serial walk, so what counts is that no warp slot idles while others still hold two items.
static cudaError_t launch_atrous_slide(svgf_ctx *c, const AtrousK &k, const AtrousArgs &a) {
    AtrousS t;
    memset(&t, 0, sizeof(t));
    t.k = k; t.kl = c->kl;
    t.ho = a.ho;
    for (int i = 0; i < SVGF_MAX_RANKS - 1; i++) {
        const bool on = i < a.ho.peers.n && a.cv_out && a.dst_slot >= 0;
        t.cv_peer[i] = on ? c->p_cv[a.dst_slot].p[a.ho.peers.rank[i]] : nullptr;
        t.lv_peer[i] = on ? c->p_lv[a.dst_slot].p[a.ho.peers.rank[i]] : nullptr;
    }
    const bool use_tma = c->tma_ok && a.src_slot >= 0;
    if (use_tma) {
        t.tm_cv = *tmap_slide(c, a.src_slot, a.level); t.tm_lv = *tmap_slide(c, SVGF_NCV + a.src_slot, a.level);
        t.tm_np = *tmap_slide(c, 2 * SVGF_NCV, a.level); t.tm_zl = *tmap_slide(c, 2 * SVGF_NCV + 1, a.level);
    }
    const void *fns[4] = {(const void *)atrous_slide_kernel<false, false>, (const void *)atrous_slide_kernel<false, true>,
                          (const void *)atrous_slide_kernel<true, false>, (const void *)atrous_slide_kernel<true, true>};
    if (!c->atrous_slide_attr_set) {
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < 4 && e == cudaSuccess; i++) {
            e = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, SL_SMEM);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(fns[i], cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        }
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->atrous_slide_blocks_per_sm, atrous_slide_kernel<true, false>, SL_WARPS * 32, SL_SMEM);
        int dev = 0, sms = 148;
        cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        c->atrous_slide_sms = sms;
        if (e != cudaSuccess) return e;
        c->atrous_slide_attr_set = true;
    }
    const int step = k.step;
    SlGrid &g = t.g;
    g.step = step; g.ncg = step / 2;
    const int lat_w = (c->W + step - 1) / step;
    g.strips = (lat_w + SL_USE - 1) / SL_USE;
    g.b_first = k.row_begin / step; g.b_end = (k.row_end - 1) / step + 1;
    const int lat_rows = g.b_end - g.b_first;
    const int slots = c->atrous_slide_sms * std::max(1, c->atrous_slide_blocks_per_sm) * SL_WARPS;
    const int columns = g.ncg * step * g.strips;            // items per band
    int bands = c->atrous_slide_bands > 0 ? c->atrous_slide_bands : std::max(1, slots / std::max(1, columns));
    if (columns > slots && c->atrous_slide_bands <= 0) bands = 2;      // several waves anyway: shorter items even the tail out
    bands = std::min(bands, std::max(1, lat_rows / 4));     // a band pays for 4 rows of run-in
    g.band_rows = (lat_rows + bands - 1) / bands;
    g.bands = (lat_rows + g.band_rows - 1) / g.band_rows;
    const int items = g.items();
    const int blocks = std::min((items + SL_WARPS - 1) / SL_WARPS, slots / SL_WARPS);
    const dim3 grid(std::max(blocks, 1)), block(SL_WARPS * 32);
    if (use_tma) { if (k.is_last) atrous_slide_kernel<true, true><<<grid, block, SL_SMEM, c->stream>>>(t); else atrous_slide_kernel<true, false><<<grid, block, SL_SMEM, c->stream>>>(t); }
    else { if (k.is_last) atrous_slide_kernel<false, true><<<grid, block, SL_SMEM, c->stream>>>(t); else atrous_slide_kernel<false, false><<<grid, block, SL_SMEM, c->stream>>>(t); }
    return cudaGetLastError();
}

void preload_atrous_kernels() {
    cudaFuncAttributes a;
    for (int s = 0; s < AT_NSHAPES; s++) cudaFuncGetAttributes(&a, g_at_shapes[s].fn);
    cudaFuncGetAttributes(&a, atrous_tiled_kernel<16, 16, 2, 5, false>); cudaFuncGetAttributes(&a, atrous_tiled_kernel<16, 12, 2, 6, false>);
    cudaFuncGetAttributes(&a, atrous_kl_kernel); cudaFuncGetAttributes(&a, atrous_direct_kernel); cudaFuncGetAttributes(&a, atrous_slide_kernel<true, false>); cudaFuncGetAttributes(&a, atrous_slide_kernel<true, true>);
    cudaFuncGetAttributes(&a, atrous_slide_kernel<false, false>); cudaFuncGetAttributes(&a, atrous_slide_kernel<false, true>);
    cudaFuncGetAttributes(&a, atrous_pair_kernel<16, 16, 2, 3>); cudaFuncGetAttributes(&a, atrous_pair_kernel<16, 16, 1, 3>);
    cudaFuncGetAttributes(&a, atrous_pair_kernel<16, 12, 2, 3>); cudaFuncGetAttributes(&a, atrous_pair_kernel<16, 12, 1, 3>);
    (void)cudaGetLastError();
}

// The kernel-side description of one level (everything but the tile shape's tensor maps).
static void at_fill_level(svgf_ctx *c, const AtrousArgs &a, AtrousT &t) {
    AtrousK k;
    k.cv_in = a.cv_in; k.cv_out = a.cv_out; k.lv_in = a.lv_in; k.lv_out = a.lv_out;
    k.nrm = a.nrm; k.pos = a.pos; k.alb = a.alb;
    k.gnp = a.gnp; k.gzl = a.gzl;
    k.denoised_out = a.denoised_out; k.var_out = a.var_out;
    k.W = c->W; k.H = c->H; k.row_begin = c->shard.row_begin; k.row_end = c->shard.row_end; k.step = 1 << a.level;
    k.is_last = a.is_last; k.blur_variance = a.blur_variance; k.addcolor = a.addcolor;
    k.sigma_c = a.sigma_c;
    atrous_scales(a.sigma_n, a.sigma_x, &k.kn, &k.kx);
    t.k = k; t.kl = c->kl; t.ro = c->rows; t.me = c->shard.rank;
    t.ho = a.ho;
    for (int i = 0; i < SVGF_MAX_RANKS - 1; i++) {
        const bool on = i < a.ho.peers.n && a.cv_out && a.dst_slot >= 0;
        t.cv_peer[i] = on ? c->p_cv[a.dst_slot].p[a.ho.peers.rank[i]] : nullptr;
        t.lv_peer[i] = on ? c->p_lv[a.dst_slot].p[a.ho.peers.rank[i]] : nullptr;
    }
    // push mode: the neighbours' rows this level taps were copied into this rank's planes by their producers (api.cu)
    const bool local_only = c->halo_push || c->rows.world <= 1;
    if (local_only) { t.ro.world = 1; t.me = 0; }
    for (int r = 0; r < SVGF_MAX_RANKS; r++) {
        const bool peer = !local_only && r < c->rows.world && a.src_slot >= 0;
        t.p_cv.p[r] = peer ? c->p_cv[a.src_slot].p[r] : a.cv_in; t.p_lv.p[r] = peer ? c->p_lv[a.src_slot].p[r] : a.lv_in;
        t.p_gnp.p[r] = peer ? c->p_gnp.p[r] : a.gnp; t.p_gzl.p[r] = peer ? c->p_gzl.p[r] : a.gzl;
    }
    t.b_first = k.row_begin / k.step;
    t.ncg = k.step / AT_C;
    t.use_tma = 0;
#ifdef SVGF_ATROUS_PROBES
    t.probe = c->atrous_probe;
#endif
    memset(&t.tm_cv, 0, sizeof(CUtensorMap)); memset(&t.tm_lv, 0, sizeof(CUtensorMap));
    memset(&t.tm_np, 0, sizeof(CUtensorMap)); memset(&t.tm_zl, 0, sizeof(CUtensorMap));
}
static void at_set_maps(svgf_ctx *c, const AtrousArgs &a, AtrousT &t, int shape) {
    t.use_tma = 1;
    t.tm_cv = *tmap_at(c, a.src_slot, a.level, shape); t.tm_lv = *tmap_at(c, SVGF_NCV + a.src_slot, a.level, shape);
    t.tm_np = *tmap_at(c, 2 * SVGF_NCV + 2 * a.gset, a.level, shape); t.tm_zl = *tmap_at(c, 2 * SVGF_NCV + 2 * a.gset + 1, a.level, shape);
}

// Can the whole stage go out as one launch (atrous_stage_kernel)? TMA staging for every level, no in-place peer reads (single
// GPU, or push mode), a band table that fits.
bool atrous_stage_possible(const svgf_ctx *c, const AtrousArgs *a, int n) {
    if (c->atrous_variant != 2 || !c->atrous_fused || !c->tma_ok || n < 1 || n > SVGF_MAX_LEVELS) return false;
    if (c->atrous_shape >= 0) return false;                             // a forced tile shape: A/B runs of the per-level kernel
    if (!(c->halo_push || c->rows.world <= 1)) return false;
    const int rows = c->shard.row_end - c->shard.row_begin;
    if (rows <= 0) return false;
    for (int i = 0; i < n; i++) {
        if (a[i].src_slot < 0 || a[i].level != i + 1) return false;
        if (rows / (8 << a[i].level) + 3 > ST_MAXBANDS) return false;
    }
    return true;
}

cudaError_t launch_atrous_stage(svgf_ctx *c, const AtrousArgs *a, int n) {
    static StageArgs S;             // ~10 KB: built in place, passed by value (kernel parameters up to 32 KB, CUDA 12.1+)
    memset(&S, 0, sizeof(S));
    const int rows = c->shard.row_end - c->shard.row_begin;
    int pos = 0;
    for (int i = 0; i < n; i++) {
        AtrousT &t = S.lvl[i];
        at_fill_level(c, a[i], t);
        const int step = t.k.step;
        const int lat_w = (c->W + step - 1) / step, lat_rows = (t.k.row_end - 1) / step - t.b_first + 1;
        StageLevel &sl = S.sl[i];
        sl.shape = lat_rows >= 100 ? 0 : 1;
        at_set_maps(c, a[i], t, sl.shape == 0 ? 2 : 8);             // g_at_shapes[2] = 16 x 16, [8] = 32 x 8
        const int lx = sl.shape == 0 ? 16 : 32;
        sl.ly = sl.shape == 0 ? 16 : 8;
        sl.gx = ((lat_w + lx - 1) / lx) * t.ncg;
        sl.bands = (lat_rows + sl.ly - 1) / sl.ly;
        sl.band0 = t.b_first * step; sl.band_h = sl.ly * step;
        sl.nK = (((c->W + 3) / 4) * sl.band_h + ST_KJOBS - 1) / ST_KJOBS; sl.nT = sl.gx * step;
        sl.ahead = std::min(sl.bands, 1 + (1000 + sl.nT - 1) / sl.nT);      // a band's K work is queued >= 1000 items (> the blocks in flight) before its tiles
        sl.total_T = sl.bands * sl.nT;
        sl.items = sl.ahead * sl.nK + sl.bands * sl.nT;
        sl.begin = pos;
        pos += sl.items;
        S.wait[i] = a[i].wait;
        (void)rows;
    }
    S.nlevels = n; S.total_items = pos;
    if (!c->stage_ctr) {
        cudaError_t e = cudaMalloc((void **)&c->stage_ctr, sizeof(unsigned) * (8 + SVGF_MAX_LEVELS * ST_CTR_PER_LEVEL + 16));
        if (e == cudaSuccess) e = cudaMemset(c->stage_ctr, 0, sizeof(unsigned) * (8 + SVGF_MAX_LEVELS * ST_CTR_PER_LEVEL + 16));
        if (e != cudaSuccess) return e;
    }
    S.ctr = c->stage_ctr; S.err = c->comm_err_dev;
    cudaError_t e = cudaMemsetAsync(c->stage_ctr, 0, sizeof(unsigned) * (8 + SVGF_MAX_LEVELS * ST_CTR_PER_LEVEL), c->stream);       // (the 16 words behind: SVGF_STAGE_TIMERS sums)
    if (e != cudaSuccess) return e;
    if (!c->stage_attr_set) {
        for (int ns = 0; ns < 2; ns++) {
            const void *fn = ns ? (const void *)atrous_stage_kernel<true> : (const void *)atrous_stage_kernel<false>;
            e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) return e;
        }
        int dev = 0, sms = 0;
        cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        c->stage_blocks = 5 * (sms > 0 ? sms : 148);
        c->stage_attr_set = true;
    }
    const int grid = std::min(c->stage_blocks, S.total_items);
    if (c->gbuf_nan_possible) atrous_stage_kernel<true><<<grid, 128, ST_SMEM, c->stream>>>(S);
    else atrous_stage_kernel<false><<<grid, 128, ST_SMEM, c->stream>>>(S);
    return cudaGetLastError();
}

cudaError_t launch_atrous(svgf_ctx *c, const AtrousArgs &a) {
    const int rows = c->shard.row_end - c->shard.row_begin;
    if (rows <= 0) return cudaSuccess;
    AtrousT t;
    at_fill_level(c, a, t);
    const AtrousK &k = t.k;
    if (c->atrous_variant == 1 && c->rows.world == 1) {
        dim3 b(32, 8), g((c->W + 31) / 32, (rows + 7) / 8);
        atrous_direct_kernel<<<g, b, 0, c->stream>>>(k);
        return cudaGetLastError();
    }
    PeerPtr<const float2> pv;
    for (int r = 0; r < SVGF_MAX_RANKS; r++) pv.p[r] = t.p_lv.p[r];
    // Pre-pass: kl per pixel. (Computing it per centre inside the tile kernel instead -- 9 gathers from the {lum, var} plane --
    // was measured on B200 and is SLOWER, C2 level 1: 115 vs 98 us: at coarse levels every lane's gather is its own 128-byte
    // line, ~400 L1 wavefronts per warp on the pipe the tile's shared-memory reads also use. Parity was green; removed.)
    {
        dim3 b(32, 8), g(((c->W + 3) / 4 + 31) / 32, (rows + 7) / 8);
        (void)at_launch_kernel(atrous_kl_kernel, g, b, 0, c->stream, pv, t.ro, c->kl, c->W, c->H, k.row_begin, k.row_end, k.blur_variance, k.sigma_c, a.wait);
    }
    const int step = k.step;
    const int lat_w = (c->W + step - 1) / step;                                 // lattice columns per class
    const int lat_rows = (k.row_end - 1) / step - t.b_first + 1;                // lattice rows touching the strip
    if (c->atrous_variant == 4) return launch_atrous_pair(c, t, a, lat_w, lat_rows);
    if (c->atrous_variant == 5) return launch_atrous_slide(c, k, a);
    const int shape = at_pick_shape(c, a.level, lat_w, lat_rows);
    const AtShapeInfo &si = g_at_shapes[shape];
    if (c->tma_ok && a.src_slot >= 0 && c->atrous_variant != 3) at_set_maps(c, a, t, shape);
    if (!c->atrous_attr_set) {      // per context (= per device)
        for (int s = 0; s < AT_NSHAPES; s++) {
            cudaError_t e = cudaFuncSetAttribute(g_at_shapes[s].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, g_at_shapes[s].smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(g_at_shapes[s].fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) return e;
        }
        const void *nn[2] = {(const void *)atrous_tiled_kernel<16, 16, 2, 5, false>, (const void *)atrous_tiled_kernel<16, 12, 2, 6, false>};
        const int nn_smem[2] = {g_at_shapes[2].smem, g_at_shapes[9].smem};
        for (int s = 0; s < 2; s++) {
            cudaError_t e = cudaFuncSetAttribute(nn[s], cudaFuncAttributeMaxDynamicSharedMemorySize, nn_smem[s]);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(nn[s], cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) return e;
        }
        c->atrous_attr_set = true;
    }
    dim3 g(((lat_w + si.lx - 1) / si.lx) * t.ncg, ((lat_rows + si.ly - 1) / si.ly) * step);
    if (!c->gbuf_nan_possible && shape == 2) at_launch_nonan<16, 16, 2, 5>(g, c->stream, t);
    else if (!c->gbuf_nan_possible && shape == 9) at_launch_nonan<16, 12, 2, 6>(g, c->stream, t);
    else si.launch(g, c->stream, t);
    return cudaGetLastError();
}
