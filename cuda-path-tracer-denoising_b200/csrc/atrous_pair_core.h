// atrous_pair_core.h -- the two phases of the symmetric ("pair") a-trous tile kernel, written as plain functions of
// (work-item index, shared-memory arrays) so that the SAME source runs inside atrous_pair_kernel (csrc/atrous.cu) and,
// item by item, in the host emulation the CPU tests use to check its indexing against the oracle (tests/emu/pair_emu.cpp;
// test infrastructure -- the product never executes this code on the host).
//
// Idea. The edge-stopping exponent of a (centre p, tap q) pair is
//     e_pq = |l_q - l_p| * kl_p  +  |n'_q - n'_p| + |p'_q - p'_p|  -  log2 h_pq
// and everything but kl_p is symmetric in (p, q). The one-phase kernel evaluates the two square roots for every ORDERED
// pair; here phase 1 evaluates  g = |dn'| + |dp'| - log2 h  once per UNORDERED pair (the 12 "forward" offsets of every staged
// lattice point: (1,0) (2,0) and (-2..2, 1) (-2..2, 2)) into shared memory, and phase 2 spends one ex2 per ordered pair on
// ex2(-(|dl| * kl_p + g)). Per ordered pair: 2 MUFU instead of 3, ~17 instead of ~25 FMA-pipe cycles.
#pragma once

#ifdef __CUDACC__
#define PAIR_FN __device__ __forceinline__
#else
#include <cmath>
#define PAIR_FN inline
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
#endif

PAIR_FN float pair_sqrt(float x) {
#ifdef __CUDA_ARCH__
    float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#else
    return sqrtf(x);
#endif
}
PAIR_FN float pair_ex2(float x) {
#ifdef __CUDA_ARCH__
    float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#else
    return exp2f(x);
#endif
}

template <int LX_, int LY_, int PR_ = 2> struct PairShape {
    static constexpr int LX = LX_, LY = LY_, C = 2, PR = PR_;  // PR: centre rows per phase-2 thread (2 x PR patch)
    static constexpr int SW = LX + 4, SH = LY + 4;              // staged lattice points (tile + 2-point apron)
    static constexpr int GROWS = LY + 2;                        // rows whose forward pairs can reach a centre
    static constexpr int TILE = SW * SH * C, HALF = SH * (SW / 2) * C, HALFP = HALF, TILEP = TILE;      // (no padding needed for these shapes: see OK)
    static constexpr int GN = GROWS * SW * C, GHALF = GROWS * (SW / 2) * C;
    static constexpr int NOFF = 12;
    static constexpr int THREADS = (LX / 2) * (LY / PR) * C;    // phase 2: one 2 x PR patch of centres per thread
    static constexpr int ITEMS = GROWS * (SW / 2) * C;          // phase 1: one (row, column pair, sub-column) per item
    static constexpr int SMEM = TILE * 48 + NOFF * GN * 4 + 16; // planes + g + mbarrier
    static constexpr bool OK = HALF * 8 % 128 == 0 && TILE * 8 % 128 == 0 && THREADS % 32 == 0 && (TILE * 48) % 16 == 0 && LY % PR == 0;
    // [column parity][lattice row][column pair][c] -- the order a TMA box arrives in (same as AtShape::idx)
    PAIR_FN static int idx(int c, int tb, int ta) { return (ta & 1) * HALF + (tb * (SW / 2) + (ta >> 1)) * C + c; }
    PAIR_FN static int gidx(int o, int c, int tb, int ta) { return o * GN + (ta & 1) * GHALF + (tb * (SW / 2) + (ta >> 1)) * C + c; }
};

// forward offsets: j > 0, or j == 0 and i > 0
PAIR_FN constexpr bool pair_forward(int i, int j) { return j > 0 || (j == 0 && i > 0); }
PAIR_FN constexpr int pair_oidx(int i, int j) { return j == 0 ? i - 1 : (j == 1 ? 2 + (i + 2) : 7 + (i + 2)); }
// -log2 of the a-trous tap weight h = h1(i) * h1(j), h1 = {3/8, 1/4, 1/16} for |.| = 0, 1, 2 (denoise.cu:84-92)
PAIR_FN constexpr float pair_nlg1(int a) { return (a == 0) ? 1.4150374992788437f : ((a == 1 || a == -1) ? 2.0f : 4.0f); }
PAIR_FN constexpr float pair_nlog2h(int i, int j) { return pair_nlg1(i) + pair_nlg1(j); }

// ---- phase 1: item n = (sub-column c, column pair pc, staged row tb): the forward pairs of the two points (2pc, tb), (2pc+1, tb)
template <class SH>
PAIR_FN void pair_phase1_item(int n, const float4 *s_np, const float2 *s_zl, float *s_g) {
    const int c = n & 1, pc = (n >> 1) % (SH::SW / 2), tb = (n >> 1) / (SH::SW / 2);
    float2 px[2], py[2], pz[2];         // negated {kn n, kx p} of the two points
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int si = SH::idx(c, tb, 2 * pc + q);
        const float4 np = s_np[si]; const float2 zl = s_zl[si];
        px[q] = make_float2(-np.x, -np.y); py[q] = make_float2(-np.z, -np.w); pz[q] = make_float2(-zl.x, -zl.y);
    }
#pragma unroll
    for (int j = 0; j <= 2; j++) {
#pragma unroll
        for (int w = (j == 0 ? 3 : 0); w < 6; w++) {    // window column w = staged column 2pc - 2 + w
            // Column 2pc - 2 + w may leave the stage at the tile's left/right edge (pc = 0, SW/2 - 1). Such a pair has no centre and
            // its g is never read; the index is NOT clamped: it stays affine in pc (base + constant, parity = w & 1) and lands
            // at most two entries outside the plane, i.e. in the neighbouring plane of the same shared-memory block (the planes
            // are laid out cv | np | zl | lv | g), so the read is harmless.
            const int si = (w & 1) * SH::HALF + ((tb + j) * (SH::SW / 2) + (pc - 1 + (w >> 1))) * SH::C + c;
            const float4 np = s_np[si]; const float2 zl = s_zl[si];
            const float2 tx = make_float2(np.x, np.y), ty = make_float2(np.z, np.w);
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int i = w - 2 - q;
                if (i < -2 || i > 2 || !pair_forward(i, j)) continue;       // compile-time
                const float2 dx = __fadd2_rn(tx, px[q]), dy = __fadd2_rn(ty, py[q]), dz = __fadd2_rn(zl, pz[q]);
                const float2 d2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
                // NaN distance -> 0, i.e. weight 1 like the reference's min(1, expf(-NaN)) (see dist_of() in atrous.cu)
                const float g = (fmaxf(pair_sqrt(d2.x), 0.0f) + fmaxf(pair_sqrt(d2.y), 0.0f)) + pair_nlog2h(i, j);
                s_g[SH::gidx(pair_oidx(i, j), c, tb, 2 * pc + q)] = g;
            }
        }
    }
}

struct PairAcc { float2 w, w2, r, g, b, v; };      // {centre column 0, centre column 1} of one patch row

// ---- phase 2: thread (c, ap, bq) owns the centres (2ap+2+ca, PR*bq+2+cb), ca in {0, 1}, cb < PR; kl[ca][cb] from the pre-pass
template <class SH>
PAIR_FN void pair_phase2_thread(int c, int ap, int bq, const float4 *s_cv, const float2 *s_lv, const float *s_g,
                                const float (&kl)[2][SH::PR], PairAcc (&A)[SH::PR]) {
    constexpr float H00 = 0.375f * 0.375f;
    constexpr int PR = SH::PR;
    float2 CL[PR], KL[PR];
#pragma unroll
    for (int cb = 0; cb < PR; cb++) {
        const int s0 = SH::idx(c, PR * bq + 2 + cb, 2 * ap + 2), s1 = SH::idx(c, PR * bq + 2 + cb, 2 * ap + 3);
        CL[cb] = make_float2(-s_lv[s0].x, -s_lv[s1].x);
        KL[cb] = make_float2(kl[0][cb], kl[1][cb]);
        const float4 c0 = s_cv[s0], c1 = s_cv[s1];
        // the centre tap: |differences| = 0, weight h(0,0) exactly
        A[cb].w = make_float2(H00, H00); A[cb].w2 = make_float2(H00 * H00, H00 * H00);
        A[cb].r = make_float2(c0.x * H00, c1.x * H00); A[cb].g = make_float2(c0.y * H00, c1.y * H00);
        A[cb].b = make_float2(c0.z * H00, c1.z * H00); A[cb].v = make_float2(c0.w * (H00 * H00), c1.w * (H00 * H00));
    }
#pragma unroll
    for (int u = 0; u < PR + 4; u++) {
#pragma unroll
        for (int tt = 0; tt < 6; tt++) {
            const int trow = PR * bq + u, tcol = 2 * ap + tt;
            const int si = SH::idx(c, trow, tcol);
            const float4 cv = s_cv[si];
            const float lum = s_lv[si].x;
#pragma unroll
            for (int cb = 0; cb < PR; cb++) {
                const int j = u - 2 - cb;
                if (j < -2 || j > 2) continue;                                  // compile-time
                const int i0 = tt - 2, i1 = tt - 3;
                const bool ok0 = i0 >= -2 && i0 <= 2 && !(i0 == 0 && j == 0), ok1 = i1 >= -2 && i1 <= 2 && !(i1 == 0 && j == 0);
                if (!ok0 && !ok1) continue;
                const int crow = PR * bq + 2 + cb;
                float g0 = 0.f, g1 = 0.f;
                if (ok0) g0 = pair_forward(i0, j) ? s_g[SH::gidx(pair_oidx(i0, j), c, crow, 2 * ap + 2)] : s_g[SH::gidx(pair_oidx(-i0, -j), c, trow, tcol)];
                if (ok1) g1 = pair_forward(i1, j) ? s_g[SH::gidx(pair_oidx(i1, j), c, crow, 2 * ap + 3)] : s_g[SH::gidx(pair_oidx(-i1, -j), c, trow, tcol)];
                PairAcc &a = A[cb];
                if (ok0 && ok1) {
                    const float2 dl = __fadd2_rn(make_float2(lum, lum), CL[cb]);
                    const float2 e = __ffma2_rn(make_float2(fabsf(dl.x), fabsf(dl.y)), KL[cb], make_float2(g0, g1));
                    const float2 w = make_float2(pair_ex2(-e.x), pair_ex2(-e.y)), w2 = __fmul2_rn(w, w);
                    a.w = __fadd2_rn(a.w, w); a.w2 = __fadd2_rn(a.w2, w2);
                    a.r = __ffma2_rn(make_float2(cv.x, cv.x), w, a.r); a.g = __ffma2_rn(make_float2(cv.y, cv.y), w, a.g);
                    a.b = __ffma2_rn(make_float2(cv.z, cv.z), w, a.b); a.v = __ffma2_rn(make_float2(cv.w, cv.w), w2, a.v);
                } else if (ok0) {
                    const float w = pair_ex2(-fmaf(fabsf(lum + CL[cb].x), KL[cb].x, g0)), w2 = w * w;
                    a.w.x += w; a.w2.x += w2; a.r.x = fmaf(cv.x, w, a.r.x); a.g.x = fmaf(cv.y, w, a.g.x); a.b.x = fmaf(cv.z, w, a.b.x); a.v.x = fmaf(cv.w, w2, a.v.x);
                } else {
                    const float w = pair_ex2(-fmaf(fabsf(lum + CL[cb].y), KL[cb].y, g1)), w2 = w * w;
                    a.w.y += w; a.w2.y += w2; a.r.y = fmaf(cv.x, w, a.r.y); a.g.y = fmaf(cv.y, w, a.g.y); a.b.y = fmaf(cv.z, w, a.b.y); a.v.y = fmaf(cv.w, w2, a.v.y);
                }
            }
        }
    }
}
