// atrous_tile_core.h -- the per-thread part of the lattice-tiled a-trous kernel (csrc/atrous.cu: atrous_tiled_kernel): tile
// shapes, the packed pair arithmetic and the walk over the tap window, as plain functions of (thread coordinates, shared-memory
// arrays). Like atrous_pair_core.h it compiles for the device (the kernel) and for the host, where tests/emu/tile_emu.cpp runs
// it thread by thread on tiles staged the way TMA stages them, so that the CPU suite checks the production kernel's arithmetic
// and indexing against the oracle for every tile shape (tests/test_atrous_emu.py). Test infrastructure only: the product
// never executes this code on the host.
#pragma once
#include "atrous_pair_core.h"       // float2/float4 and the packed-fp32 intrinsics on the host, pair_sqrt / pair_ex2, PairAcc

// Distance for an edge-stopping weight. The reference clamps the normal and position weights with min(1.0f, expf(-d/s))
// (denoise.cu:144-145; CUDA's min = fminf, which drops a NaN operand): for d >= 0 that is the identity, but a NaN distance --
// a mesh without vertex normals interpolates normalize(0) = NaN (sceneStructs.h:168-172) -- yields weight 1, not NaN. fmaxf
// drops the NaN here (one FMNMX on the ALU pipe, which has slack), so such a pair is filtered as if the two normals agreed.
// NS = false: the caller guarantees that no normal or position in the frame is NaN (svgf_ctx::gbuf_nan_possible, decided from
// the scene at upload); the two FMNMX per pair go away.
template <bool NS = true> PAIR_FN float dist_of(float d2) { return NS ? fmaxf(pair_sqrt(d2), 0.0f) : pair_sqrt(d2); }

constexpr int AT_C = 2, AT_TX = 2;
// Tile shape <LX, LY, TY>: LX x LY lattice points (x 2 columns) per block, every thread a 2 x TY patch of them.
//   TY = 4: 128 threads per 512 points, 168 registers, 69 KB -> 3 blocks (12 warps) per SM; 4.2 pair evaluations per tap read.
//   TY = 2: twice the threads per point at half the registers (more warps to hide the MUFU/LDS latencies, 2.8 pairs per tap
//           read), and 16 x 16 tiles of 38 KB so that 5 blocks per SM overlap their tile loads with each other's arithmetic.
// 16x32 suits fine levels, 32x16 / 32x12 lattices that are short (coarse levels: 34 lattice rows at step 32 for 1080 rows;
// narrow strips of a sharded frame). launch_atrous() picks per level; SVGF_ATROUS_SHAPE=<id> forces one (A/B runs).
template <int LX, int LY, int TY_> struct AtShape {
    static constexpr int TY = TY_;
    static constexpr int SW = LX + 4, SH = LY + 4;                  // staged lattice points (tile + 2-point apron)
    static constexpr int THREADS = (LX / AT_TX) * (LY / TY) * AT_C;
    static constexpr int TILE = SW * SH * AT_C;                     // taps
    static constexpr int HALF = SH * (SW / 2) * AT_C;               // taps of one column parity = one TMA box
    // TMA destinations (every plane, and the second parity half of every plane) must be 128-byte aligned: the parity halves sit
    // HALFP entries apart (HALF rounded up to 16 entries = 128 bytes of the 8-byte planes), a plane holds TILEP entries
    static constexpr int HALFP = (HALF + 15) / 16 * 16, TILEP = 2 * HALFP;
    static constexpr int SMEM = TILEP * 48 + 16;                    // + mbarrier
    static constexpr bool OK = THREADS % 32 == 0 && LX % 2 == 0 && LY % TY == 0;
    // [column parity][lattice row][column pair][c]  -- the order a TMA box arrives in
    PAIR_FN static int idx(int c, int tb, int ta) { return (ta & 1) * HALFP + (tb * (SW / 2) + (ta >> 1)) * AT_C + c; }
};


// Edge-stopping weights and accumulation, written for Blackwell's packed fp32 pipe (FADD2/FMUL2/FFMA2, sm_100+).
//   * distances: normal and position differences travel as float2 {n, p} lanes, so the two squared distances of a pair cost
//     6 packed instructions instead of 12;
//   * everything after the square roots is packed ACROSS THE TWO CENTRES of a patch row (ca = 0, 1), which meet the same tap
//     in tap columns 1..4: {e0, e1} = |{lq, lq} - {l0, l1}| * {kl0, kl1} + {dn0, dn1} + {dp0, dp1}, w = ex2(-e) * {h0, h1}, and
//     the six sums of both centres advance with six packed instructions. The tap's colour enters as a broadcast operand
//     (`R.F32` in SASS) and |.| is an operand modifier, so no instruction is spent on forming pairs.
// Why this matters (tools/pipe_probe.cu on B200, cycles per trip per SM sub-partition): 3 MUFU + 18 scalar FFMA take 36.5
// cycles although neither pipe needs more than 24 -- MUFU and scalar FMA issue get in each other's way -- while
// 3 MUFU + 8 FFMA2 + 8 integer adds take 25.0. Per pair the kernel needs 3 MUFU (24 XU cycles) and ~23 FMA-pipe cycles either
// way; issued as ~12 packed instructions instead of 8 packed + 7 scalar they overlap.
struct AtCentre2 {          // the two centres of one patch row; G-buffer terms negated so that tap + centre = difference
    float2 nx_px[AT_TX], ny_py[AT_TX], nz_pz[AT_TX];    // {-kn*n, -kx*p} per component, per centre
    float2 lum, kl;                                     // {centre 0, centre 1}
};
using AtAcc2 = PairAcc;                                 // {centre 0, centre 1}: sum w, sum w^2, sum w*rgb, sum w^2*var
struct AtTap { float4 cv; float2 nx_px, ny_py, nz_pz; float lum; };

PAIR_FN float2 at_dist2(const AtTap &T, float2 cx, float2 cy, float2 cz) {       // {|dn|^2, |dp|^2}
    const float2 dx = __fadd2_rn(T.nx_px, cx), dy = __fadd2_rn(T.ny_py, cy), dz = __fadd2_rn(T.nz_pz, cz);
    return __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
}

// one tap against both centres of a patch row; h = {h of centre 0, h of centre 1}
template <bool NS = true>
PAIR_FN void at_twin(const AtTap &T, const AtCentre2 &C, AtAcc2 &A, float2 h) {
    const float2 d0 = at_dist2(T, C.nx_px[0], C.ny_py[0], C.nz_pz[0]), d1 = at_dist2(T, C.nx_px[1], C.ny_py[1], C.nz_pz[1]);
    const float2 dn = make_float2(dist_of<NS>(d0.x), dist_of<NS>(d1.x)), dp = make_float2(dist_of<NS>(d0.y), dist_of<NS>(d1.y));
    const float2 dl = __fadd2_rn(make_float2(T.lum, T.lum), make_float2(-C.lum.x, -C.lum.y));
    const float2 e = __fadd2_rn(__ffma2_rn(make_float2(fabsf(dl.x), fabsf(dl.y)), C.kl, dn), dp);
    const float2 w = __fmul2_rn(make_float2(pair_ex2(-e.x), pair_ex2(-e.y)), h);
    const float2 w2 = __fmul2_rn(w, w);
    A.w = __fadd2_rn(A.w, w);
    A.w2 = __fadd2_rn(A.w2, w2);
    A.r = __ffma2_rn(make_float2(T.cv.x, T.cv.x), w, A.r);
    A.g = __ffma2_rn(make_float2(T.cv.y, T.cv.y), w, A.g);
    A.b = __ffma2_rn(make_float2(T.cv.z, T.cv.z), w, A.b);
    A.v = __ffma2_rn(make_float2(T.cv.w, T.cv.w), w2, A.v);
}

// one tap against ONE centre of the row (tap columns 0 and 5 reach one centre column only): same operations per lane
template <int CA, bool NS = true>
PAIR_FN void at_single(const AtTap &T, const AtCentre2 &C, AtAcc2 &A, float h) {
    const float2 d2 = at_dist2(T, C.nx_px[CA], C.ny_py[CA], C.nz_pz[CA]);
    const float dn = dist_of<NS>(d2.x), dp = dist_of<NS>(d2.y);
    const float lum = CA ? C.lum.y : C.lum.x, kl = CA ? C.kl.y : C.kl.x;
    const float e = fmaf(fabsf(T.lum - lum), kl, dn) + dp;
    const float w = pair_ex2(-e) * h, w2 = w * w;
    if (CA) { A.w.y += w; A.w2.y += w2; A.r.y = fmaf(T.cv.x, w, A.r.y); A.g.y = fmaf(T.cv.y, w, A.g.y); A.b.y = fmaf(T.cv.z, w, A.b.y); A.v.y = fmaf(T.cv.w, w2, A.v.y); }
    else { A.w.x += w; A.w2.x += w2; A.r.x = fmaf(T.cv.x, w, A.r.x); A.g.x = fmaf(T.cv.y, w, A.g.x); A.b.x = fmaf(T.cv.z, w, A.b.x); A.v.x = fmaf(T.cv.w, w2, A.v.x); }
}

// One tap column (window column `tt` of the thread's 6) against the thread's 2 x TY centres. DO0/DO1 select which of the
// two centre columns the tap column reaches (|i| <= 2), so the edge columns are peeled without wasted work.
template <class SH, bool DO0, bool DO1, bool NS = true>
PAIR_FN void at_column(const float4 *s_cv, const float4 *s_np, const float2 *s_zl, const float2 *s_lv, int c, int row0, int col,
                                          const AtCentre2 (&C)[SH::TY], AtAcc2 (&A)[SH::TY], float hi0, float hi1) {
    // h = hi * hj with hj in {3/8, 1/4, 1/16} for |j| = 0, 1, 2
    const float2 hh[3] = {make_float2(hi0 * 0.375f, hi1 * 0.375f), make_float2(hi0 * 0.25f, hi1 * 0.25f), make_float2(hi0 * 0.0625f, hi1 * 0.0625f)};
#pragma unroll
    for (int u = 0; u < SH::TY + 4; u++) {
        const int si = SH::idx(c, row0 + u, col);
        const float4 np = s_np[si];
        AtTap T;
        T.cv = s_cv[si];
        T.nx_px = make_float2(np.x, np.y); T.ny_py = make_float2(np.z, np.w); T.nz_pz = s_zl[si]; T.lum = s_lv[si].x;
#pragma unroll
        for (int cb = 0; cb < SH::TY; cb++) {
            const int j = u - 2 - cb, aj = j < 0 ? -j : j;
            if (aj > 2) continue;       // compile-time
            if (DO0 && DO1) at_twin<NS>(T, C[cb], A[cb], hh[aj]);
            else if (DO0) at_single<0, NS>(T, C[cb], A[cb], hh[aj].x);
            else at_single<1, NS>(T, C[cb], A[cb], hh[aj].y);
        }
    }
}

// Everything one thread does between the tile having landed and the sums being complete: its 2 x TY centres against the
// 6 x (TY + 4) tap window.
template <class SH, bool NS = true>
PAIR_FN void at_thread_compute(int c, int ap, int bq, const float4 *s_cv, const float4 *s_np, const float2 *s_zl, const float2 *s_lv,
                               const float (&c_kl)[AT_TX][SH::TY], AtAcc2 (&A)[SH::TY]) {
    constexpr int AT_TY = SH::TY;
    AtCentre2 C[AT_TY];
#pragma unroll
    for (int cb = 0; cb < AT_TY; cb++) {
        float cl[AT_TX];
#pragma unroll
        for (int ca = 0; ca < AT_TX; ca++) {
            const int si = SH::idx(c, AT_TY * bq + cb + 2, 2 * ap + ca + 2);
            const float4 np = s_np[si]; const float2 zl = s_zl[si];
            C[cb].nx_px[ca] = make_float2(-np.x, -np.y); C[cb].ny_py[ca] = make_float2(-np.z, -np.w);
            C[cb].nz_pz[ca] = make_float2(-zl.x, -zl.y); cl[ca] = s_lv[si].x;
        }
        C[cb].lum = make_float2(cl[0], cl[1]); C[cb].kl = make_float2(c_kl[0][cb], c_kl[1][cb]);
        const float2 z = make_float2(0.f, 0.f);
        A[cb].w = z; A[cb].w2 = z; A[cb].r = z; A[cb].g = z; A[cb].b = z; A[cb].v = z;
    }

    // ---- 6 tap columns x 8 tap rows. Columns 1..4 reach both centre columns and run as a rolled loop whose body is one
    // large basic block (8 taps, 40 independent pair evaluations: plenty of ILP for 12 warps/SM, 13 KB of SASS);
    // columns 0 and 5 reach one centre column each and are peeled. The centre tap takes the generic path: all
    // differences are 0, sqrt(0) = 0, ex2(-0) = 1 exactly. ----
    const int row0 = AT_TY * bq, col0 = 2 * ap;
    at_column<SH, true, false, NS>(s_cv, s_np, s_zl, s_lv, c, row0, col0 + 0, C, A, 0.0625f, 0.f);        // i = -2 for centre column 0
#pragma unroll 1
    for (int tt = 1; tt <= 4; tt++) {
        const int i0 = tt - 2, i1 = tt - 3;
        const float hi0 = i0 == 0 ? 0.375f : ((i0 == 1 || i0 == -1) ? 0.25f : 0.0625f);
        const float hi1 = i1 == 0 ? 0.375f : ((i1 == 1 || i1 == -1) ? 0.25f : 0.0625f);
        at_column<SH, true, true, NS>(s_cv, s_np, s_zl, s_lv, c, row0, col0 + tt, C, A, hi0, hi1);
    }
    at_column<SH, false, true, NS>(s_cv, s_np, s_zl, s_lv, c, row0, col0 + 5, C, A, 0.f, 0.0625f);        // i = +2 for centre column 1

}
