// atrous_slide_core.h -- per-lane code of the "sliding" a-trous kernel (csrc/atrous.cu: atrous_slide_kernel), the symmetric
// formulation of ATrousFilter (src/denoise.cu:77-170) with the pair distances kept in REGISTERS.
//
// The edge-stopping exponent of a (centre p, tap q) pair is
//     e_pq = |l_q - l_p| * kl_p  +  g_pq,      g_pq = |n'_q - n'_p| + |p'_q - p'_p| - log2 h_pq
// and g is symmetric in (p, q): of the three MUFU operations of a pair (two square roots, one ex2) and its ~23 FMA-pipe
// operations, the two square roots and 12 of the operations need to be done once per UNORDERED pair only. A first attempt
// (atrous_pair_core.h) passed g through shared memory and became shared-memory bound. Here it never leaves the register file:
//
//   * a warp owns 16 adjacent lattice columns (x 2 sub-columns: lane = 2 * column + c) of one residue class and slides DOWN
//     the rows: step T brings in tap row T (5 taps per lane: its own column and two either side);
//   * five centres per lane are in flight, the rows T-2 .. T+2 of the lane's column, each with its six running sums. A tap
//     row is therefore read from shared memory ONCE and scattered into five centres (one 48-byte tap read per 5 pairs);
//   * for the pairs whose centre lies BELOW the tap row (and to the right in the same row) the lane computes g itself and
//     keeps it; the mirrored pairs -- centre above the tap -- belong to the lane that owns the tap's column, which needs
//     exactly that value one or two steps later: it arrives by one __shfl_sync (10 shuffles per step and lane);
//   * the centres rotate through five register sets; the step is instantiated for the five phases of the rotation, so every
//     register index is a compile-time constant, and the pair arithmetic is packed (FADD2/FFMA2) across the sets (0,1), (2,3).
//
// Per ordered pair: 2 MUFU instead of 3 (1.92 with the centre tap), ~16 instead of ~23 FMA-pipe operations, ~15 instead of ~26
// issued instructions. The two columns at either edge of a warp only compute g for their neighbours (12 of 16 columns produce
// output).
//
// Like the other *_core.h files this compiles for the device (the kernel) and for the host, where tests/emu/slide_emu.cpp runs
// the 32 lanes of a warp in lock step against the oracle (tests/test_atrous_emu.py). Test infrastructure on the host; the
// product never runs it there.
#pragma once
#include "atrous_pair_core.h"       // float2/float4 and packed-fp32 intrinsics on the host, pair_sqrt / pair_ex2, pair_nlog2h

#ifdef __CUDACC__
#define SL_HD __host__ __device__ __forceinline__
#else
#define SL_HD inline
#endif

constexpr int SL_COLS = 16;         // staged lattice columns per warp
constexpr int SL_EDGE = 2;          // columns at either side that only serve their neighbours
constexpr int SL_USE = SL_COLS - 2 * SL_EDGE;
constexpr int SL_ROW = 2 * SL_COLS; // entries of one staged row = lanes of a warp: entry = 2 * column + c

struct SlTap { float4 cv; float2 nx_px, ny_py, nz_pz; float lum; };
struct SlRow { const float4 *cv, *np; const float2 *zl, *lv; };        // one staged row: SL_ROW entries per plane
struct SlCentre {                   // a centre in flight
    float2 nx_px, ny_py, nz_pz;     // its pre-scaled G-buffer record {kn n, kx p}
    float fw1[5], fw2[5];           // g of (this centre, tap column -2..2 of the row 1 / 2 above it), kept for the lanes that mirror them
};
struct SlAccP { float2 w, w2, r, g, b, v; };    // running sums of two centres (sets 0,1 or 2,3)
struct SlAccS { float w, w2, r, g, b, v; };     // ... of set 4
struct SlLane {
    SlCentre C[5];
    SlAccP A01, A23; SlAccS A4;
    float2 mlum01, kl01, mlum23, kl23; float mlum4, kl4;     // -luminance and luminance-weight scale of the five centres
};
// which register set holds the centre `dj` rows below the tap row in phase PHI (row T lives in set T mod 5), and back
PAIR_FN constexpr int sl_set(int phi, int dj) { return ((phi + dj) % 5 + 5) % 5; }
PAIR_FN constexpr int sl_dj(int phi, int k) { return ((k - phi) % 5 + 5) % 5 > 2 ? ((k - phi) % 5 + 5) % 5 - 5 : ((k - phi) % 5 + 5) % 5; }

PAIR_FN SlTap sl_load_tap(const SlRow &row, int e) {
    SlTap t;
    const float4 np = row.np[e];
    t.cv = row.cv[e]; t.nx_px = make_float2(np.x, np.y); t.ny_py = make_float2(np.z, np.w); t.nz_pz = row.zl[e]; t.lum = row.lv[e].x;
    return t;
}

// NaN distance -> 0: the reference clamps the normal/position weights with min(1, expf(-d/s)) and CUDA's fminf drops a NaN
// operand (denoise.cu:144-145; a mesh without vertex normals shades with normalize(0)), see dist_of() in atrous_tile_core.h
PAIR_FN float sl_dist(float d2) { return fmaxf(pair_sqrt(d2), 0.0f); }

PAIR_FN float sl_g(const SlCentre &c, const SlTap &t, float nlog2h) {
    const float2 dx = __fadd2_rn(t.nx_px, make_float2(-c.nx_px.x, -c.nx_px.y));
    const float2 dy = __fadd2_rn(t.ny_py, make_float2(-c.ny_py.x, -c.ny_py.y));
    const float2 dz = __fadd2_rn(t.nz_pz, make_float2(-c.nz_pz.x, -c.nz_pz.y));
    const float2 d2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
    return (sl_dist(d2.x) + sl_dist(d2.y)) + nlog2h;
}

// A centre enters flight two rows below the tap row: its own record from the staged row it lies in, kl from the pre-pass.
template <int K>
PAIR_FN void sl_enter(SlLane &L, const SlRow &row, int e, float kl) {
    const float4 np = row.np[e];
    L.C[K].nx_px = make_float2(np.x, np.y); L.C[K].ny_py = make_float2(np.z, np.w); L.C[K].nz_pz = row.zl[e];
    const float ml = -row.lv[e].x;
    const float2 z = make_float2(0.f, 0.f);
    if (K == 0) { L.mlum01.x = ml; L.kl01.x = kl; L.A01.w.x = 0.f; L.A01.w2.x = 0.f; L.A01.r.x = 0.f; L.A01.g.x = 0.f; L.A01.b.x = 0.f; L.A01.v.x = 0.f; }
    if (K == 1) { L.mlum01.y = ml; L.kl01.y = kl; L.A01.w.y = 0.f; L.A01.w2.y = 0.f; L.A01.r.y = 0.f; L.A01.g.y = 0.f; L.A01.b.y = 0.f; L.A01.v.y = 0.f; }
    if (K == 2) { L.mlum23.x = ml; L.kl23.x = kl; L.A23.w.x = 0.f; L.A23.w2.x = 0.f; L.A23.r.x = 0.f; L.A23.g.x = 0.f; L.A23.b.x = 0.f; L.A23.v.x = 0.f; }
    if (K == 3) { L.mlum23.y = ml; L.kl23.y = kl; L.A23.w.y = 0.f; L.A23.w2.y = 0.f; L.A23.r.y = 0.f; L.A23.g.y = 0.f; L.A23.b.y = 0.f; L.A23.v.y = 0.f; }
    if (K == 4) { L.mlum4 = ml; L.kl4 = kl; L.A4.w = 0.f; L.A4.w2 = 0.f; L.A4.r = 0.f; L.A4.g = 0.f; L.A4.b = 0.f; L.A4.v = 0.f; }
    (void)z;
}

// A step, in three parts.
// (1) Same-row pairs to the RIGHT (tap columns +1, +2): computed here, mirrored by the lanes owning those columns.
template <int PHI>
PAIR_FN void sl_same_row(const SlLane &L, const SlTap &t3, const SlTap &t4, float (&s)[2]) {
    constexpr int K0 = sl_set(PHI, 0);
    s[0] = sl_g(L.C[K0], t3, pair_nlog2h(1, 0));
    s[1] = sl_g(L.C[K0], t4, pair_nlog2h(2, 0));
}

// (2) What this lane holds for others: the value lane (a) needs for its centre 1 / 2 rows ABOVE the tap (a + i, T) was computed
// one / two steps ago by lane a + i, for ITS centre in row T against the tap column -i. The kernel shuffles these registers
// (srcLane = lane + 2 i); the host emulation reads them from the other lane's state.
template <int PHI> PAIR_FN float sl_offer_r1(const SlLane &L, int i) { return L.C[sl_set(PHI, 0)].fw1[2 - i]; }
template <int PHI> PAIR_FN float sl_offer_r2(const SlLane &L, int i) { return L.C[sl_set(PHI, 0)].fw2[2 - i]; }

PAIR_FN void sl_twin(SlAccP &A, float2 mlum, float2 kl, const SlTap &t, float2 g) {
    const float2 dl = __fadd2_rn(make_float2(t.lum, t.lum), mlum);
    const float2 e = __ffma2_rn(make_float2(fabsf(dl.x), fabsf(dl.y)), kl, g);
    const float2 w = make_float2(pair_ex2(-e.x), pair_ex2(-e.y));
    const float2 w2 = __fmul2_rn(w, w);
    A.w = __fadd2_rn(A.w, w); A.w2 = __fadd2_rn(A.w2, w2);
    A.r = __ffma2_rn(make_float2(t.cv.x, t.cv.x), w, A.r); A.g = __ffma2_rn(make_float2(t.cv.y, t.cv.y), w, A.g);
    A.b = __ffma2_rn(make_float2(t.cv.z, t.cv.z), w, A.b); A.v = __ffma2_rn(make_float2(t.cv.w, t.cv.w), w2, A.v);
}
PAIR_FN void sl_single(SlAccS &A, float mlum, float kl, const SlTap &t, float g) {
    const float w = pair_ex2(-fmaf(fabsf(t.lum + mlum), kl, g)), w2 = w * w;
    A.w += w; A.w2 += w2;
    A.r = fmaf(t.cv.x, w, A.r); A.g = fmaf(t.cv.y, w, A.g); A.b = fmaf(t.cv.z, w, A.b); A.v = fmaf(t.cv.w, w2, A.v);
}

// (3) One tap (column TI - 2 of the tap row) against the five centres in flight. n1/n2: g for the centres 1 / 2 rows BELOW the
// tap row, computed here and kept for the mirrored pair; r1/r2: g for the centres above, received; g0: g for the centre in the
// tap row (computed, received, or -log2 h(0,0) for the centre tap itself, which has no distance).
template <int PHI, int TI>
PAIR_FN void sl_tap(SlLane &L, const SlTap &t, float r1, float r2, float g0) {
    constexpr int K1 = sl_set(PHI, 1), K2 = sl_set(PHI, 2);
    const float n1 = sl_g(L.C[K1], t, pair_nlog2h(TI - 2, 1)), n2 = sl_g(L.C[K2], t, pair_nlog2h(TI - 2, 2));
    L.C[K1].fw1[TI] = n1; L.C[K2].fw2[TI] = n2;
    float g[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const int dj = sl_dj(PHI, k);       // compile-time
        g[k] = dj == 2 ? n2 : (dj == 1 ? n1 : (dj == 0 ? g0 : (dj == -1 ? r1 : r2)));
    }
    sl_twin(L.A01, L.mlum01, L.kl01, t, make_float2(g[0], g[1]));
    sl_twin(L.A23, L.mlum23, L.kl23, t, make_float2(g[2], g[3]));
    sl_single(L.A4, L.mlum4, L.kl4, t, g[4]);
}

// The finished centre (two rows above the tap row): {sum w, sum w^2, sum w rgb, sum w^2 var}.
template <int K>
PAIR_FN SlAccS sl_exit(const SlLane &L) {
    SlAccS o;
    if (K == 0) { o.w = L.A01.w.x; o.w2 = L.A01.w2.x; o.r = L.A01.r.x; o.g = L.A01.g.x; o.b = L.A01.b.x; o.v = L.A01.v.x; }
    if (K == 1) { o.w = L.A01.w.y; o.w2 = L.A01.w2.y; o.r = L.A01.r.y; o.g = L.A01.g.y; o.b = L.A01.b.y; o.v = L.A01.v.y; }
    if (K == 2) { o.w = L.A23.w.x; o.w2 = L.A23.w2.x; o.r = L.A23.r.x; o.g = L.A23.g.x; o.b = L.A23.b.x; o.v = L.A23.v.x; }
    if (K == 3) { o.w = L.A23.w.y; o.w2 = L.A23.w2.y; o.r = L.A23.r.y; o.g = L.A23.g.y; o.b = L.A23.b.y; o.v = L.A23.v.y; }
    if (K == 4) o = L.A4;
    return o;
}

// ---- work decomposition, shared by the kernel and the emulation ---------------------------------------------------------
// An item = (column group cg: sub-columns X0 = 2 cg, X0 + 1; row class yc; strip of SL_USE lattice columns; band of lattice rows).
struct SlItem { int X0, yc, a0, b_lo, b_hi; };      // a0 = lattice column of staged column 0; centres in rows [b_lo, b_hi)
struct SlGrid {
    int step, ncg, strips, bands, band_rows, b_first, b_end;    // lattice rows [b_first, b_end) touch the rank's rows
    SL_HD int items() const { return ncg * step * strips * bands; }
    SL_HD SlItem item(int n) const {
        SlItem it;
        const int cg = n % ncg; n /= ncg;
        const int strip = n % strips; n /= strips;
        const int band = n % bands; n /= bands;
        it.X0 = 2 * cg; it.yc = n; it.a0 = strip * SL_USE - SL_EDGE;
        it.b_lo = b_first + band * band_rows;
        it.b_hi = it.b_lo + band_rows < b_end ? it.b_lo + band_rows : b_end;
        return it;
    }
};
