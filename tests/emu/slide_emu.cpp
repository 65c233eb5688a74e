// slide_emu.cpp -- host emulation of atrous_slide_kernel (csrc/atrous.cu): the per-lane code of csrc/atrous_slide_core.h -- the
// very functions the kernel calls -- run for the 32 lanes of a warp in lock step, with the register shuffles replaced by reads
// of the other lane's state, on rows staged the way the TMA row loads + edge fix-up stage them (zero fill outside the tensor,
// aliasing into the next row right of the image, zero padding below it), for every item of the grid. Checks on the CPU what a
// GPU is not needed for: which unordered pair every (centre, tap) uses and who computes it, the five-phase register rotation,
// band run-in/run-out, image borders, strips of a sharded frame. TEST INFRASTRUCTURE (CPU suite).
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../cuda-path-tracer-denoising_b200/csrc/atrous_slide_core.h"

namespace {

struct Planes { const float *cv, *gnp, *gzl, *lv, *kl; int W, H, step, row_begin, row_end; float *out; };

struct Staged { float4 cv[SL_ROW], np[SL_ROW]; float2 zl[SL_ROW], lv[SL_ROW]; };

// what cp.async.bulk.tensor delivers for staged row (lattice row b) of an item, then the kernel's fix-up (sl_wait)
void stage_row(const Planes &P, const SlItem &it, int b, Staged &S) {
    const int s = P.step, ncells = (P.W + s - 1) / s, nbands = (P.H + s - 1) / s;
    const size_t px = (size_t)P.W * P.H;
    const bool need_fix = (it.a0 < 0) || (it.X0 + (it.a0 + SL_COLS - 1) * s + 1 >= P.W);
    for (int e = 0; e < SL_ROW; e++) {
        const int a = e >> 1, c = e & 1, la = it.a0 + a;
        const long x = it.X0 + (long)la * s + c, y = it.yc + (long)b * s;
        S.cv[e] = float4{0, 0, 0, 0}; S.np[e] = float4{0, 0, 0, 0}; S.zl[e] = float2{0, 0}; S.lv[e] = float2{0, 0};
        if (la >= 0 && la < ncells && b >= 0 && b < nbands) {
            const size_t q = (size_t)x + (size_t)y * P.W;       // columns >= W alias the next row, rows >= H are zero padding
            if (q < px && y < P.H) {
                S.cv[e] = float4{P.cv[4 * q], P.cv[4 * q + 1], P.cv[4 * q + 2], P.cv[4 * q + 3]};
                S.np[e] = float4{P.gnp[4 * q], P.gnp[4 * q + 1], P.gnp[4 * q + 2], P.gnp[4 * q + 3]};
                S.zl[e] = float2{P.gzl[2 * q], P.gzl[2 * q + 1]}; S.lv[e] = float2{P.lv[2 * q], P.lv[2 * q + 1]};
            }
        }
        const bool xv = la >= 0 && x < P.W;
        if (need_fix && !xv) S.lv[e].x = 3e38f;
    }
}

struct Warp {
    SlLane L[32];
    int x[32]; bool xv[32];
};

template <int PHI>
void run_step(const Planes &P, const SlItem &it, Warp &w, std::vector<Staged> &rows, int rt, int nrows) {
    constexpr int KE = sl_set(PHI, 2), KX = sl_set(PHI, -2);
    const int b0 = it.b_lo - 2, s = P.step;
    // enter
    {
        const SlRow row{rows[rt + 2].cv, rows[rt + 2].np, rows[rt + 2].zl, rows[rt + 2].lv};
        const long y = it.yc + (long)(b0 + rt + 2) * s;
        for (int l = 0; l < 32; l++) {
            const float kl = (w.xv[l] && b0 + rt + 2 >= 0 && y < P.H) ? P.kl[w.x[l] + (size_t)y * P.W] : 0.f;
            sl_enter<KE>(w.L[l], row, l, kl);
        }
    }
    const int b = b0 + rt; const long y = it.yc + (long)b * s;
    if (b >= 0 && y < P.H) {
        const SlRow row{rows[rt].cv, rows[rt].np, rows[rt].zl, rows[rt].lv};
        float r1[32][5], r2[32][5], sr[32][2], bk[32][2];
        int e[32][5];
        for (int l = 0; l < 32; l++) for (int ti = 0; ti < 5; ti++) { int v = l + 2 * (ti - 2); e[l][ti] = v < 0 ? 0 : (v > 31 ? 31 : v); }
        for (int l = 0; l < 32; l++) sl_same_row<PHI>(w.L[l], sl_load_tap(row, e[l][3]), sl_load_tap(row, e[l][4]), sr[l]);
        // the shuffles: __shfl_sync(v, lane + d) reads lane (lane + d) mod 32
        for (int l = 0; l < 32; l++) {
            for (int ti = 0; ti < 5; ti++) {
                const int i = ti - 2, src = ((l + 2 * i) % 32 + 32) % 32;
                r1[l][ti] = sl_offer_r1<PHI>(w.L[i == 0 ? l : src], i); r2[l][ti] = sl_offer_r2<PHI>(w.L[i == 0 ? l : src], i);
            }
            bk[l][0] = sr[((l - 2) % 32 + 32) % 32][0]; bk[l][1] = sr[((l - 4) % 32 + 32) % 32][1];
        }
        for (int l = 0; l < 32; l++) {
            sl_tap<PHI, 3>(w.L[l], sl_load_tap(row, e[l][3]), r1[l][3], r2[l][3], sr[l][0]);
            sl_tap<PHI, 4>(w.L[l], sl_load_tap(row, e[l][4]), r1[l][4], r2[l][4], sr[l][1]);
            sl_tap<PHI, 2>(w.L[l], sl_load_tap(row, e[l][2]), r1[l][2], r2[l][2], pair_nlog2h(0, 0));
            sl_tap<PHI, 1>(w.L[l], sl_load_tap(row, e[l][1]), r1[l][1], r2[l][1], bk[l][0]);
            sl_tap<PHI, 0>(w.L[l], sl_load_tap(row, e[l][0]), r1[l][0], r2[l][0], bk[l][1]);
        }
    }
    const int bo = b - 2; const long yo = it.yc + (long)bo * s;
    for (int l = 0; l < 32; l++) {
        const int a = l >> 1;
        if (rt >= 4 && rt < nrows + 4 && a >= SL_EDGE && a < SL_COLS - SL_EDGE && w.x[l] < P.W && yo >= P.row_begin && yo < P.row_end) {
            const SlAccS o = sl_exit<KX>(w.L[l]);
            float *d = P.out + 4 * ((size_t)w.x[l] + (size_t)yo * P.W);
            d[0] = o.r / o.w; d[1] = o.g / o.w; d[2] = o.b / o.w; d[3] = o.v / o.w2;
        }
    }
}

}  // namespace

// bands <= 0: one band per (class, strip)
extern "C" int slide_emu_level(const float *cv, const float *gnp, const float *gzl, const float *lv, const float *kl, int W, int H, int step,
                               int row_begin, int row_end, int bands, float *out_cv) {
    Planes P{cv, gnp, gzl, lv, kl, W, H, step, row_begin, row_end, out_cv};
    SlGrid g;
    g.step = step; g.ncg = step / 2;
    const int lat_w = (W + step - 1) / step;
    g.strips = (lat_w + SL_USE - 1) / SL_USE;
    g.b_first = row_begin / step; g.b_end = (row_end - 1) / step + 1;
    const int lat_rows = g.b_end - g.b_first;
    if (bands <= 0) bands = 1;
    if (bands > lat_rows) bands = lat_rows > 0 ? lat_rows : 1;
    g.band_rows = (lat_rows + bands - 1) / bands;
    g.bands = (lat_rows + g.band_rows - 1) / g.band_rows;
    for (int n = 0; n < g.items(); n++) {
        const SlItem it = g.item(n);
        const int nrows = it.b_hi - it.b_lo;
        if (nrows <= 0) continue;
        Warp w;
        memset(&w, 0, sizeof(w));
        for (int l = 0; l < 32; l++) {
            const int a = l >> 1, c = l & 1;
            w.x[l] = it.X0 + (it.a0 + a) * step + c; w.xv[l] = (it.a0 + a >= 0) && (w.x[l] < W);
        }
        std::vector<Staged> rows(nrows + 6);
        for (int r = 0; r < nrows + 6; r++) stage_row(P, it, it.b_lo - 2 + r, rows[r]);
        for (int r = 0; r < 2; r++) {
            const SlRow row{rows[r].cv, rows[r].np, rows[r].zl, rows[r].lv};
            const long y = it.yc + (long)(it.b_lo - 2 + r) * step;
            for (int l = 0; l < 32; l++) {
                const float k0 = (w.xv[l] && it.b_lo - 2 + r >= 0 && y < H) ? kl[w.x[l] + (size_t)y * W] : 0.f;
                if (r == 0) sl_enter<0>(w.L[l], row, l, k0); else sl_enter<1>(w.L[l], row, l, k0);
            }
        }
        const int steps = nrows + 4;
        for (int rt = 0; rt < steps; rt++) {
            switch (rt % 5) {
                case 0: run_step<0>(P, it, w, rows, rt, nrows); break;
                case 1: run_step<1>(P, it, w, rows, rt, nrows); break;
                case 2: run_step<2>(P, it, w, rows, rt, nrows); break;
                case 3: run_step<3>(P, it, w, rows, rt, nrows); break;
                default: run_step<4>(P, it, w, rows, rt, nrows); break;
            }
        }
    }
    return 0;
}
