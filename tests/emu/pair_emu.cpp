// pair_emu.cpp -- host emulation of atrous_pair_kernel (csrc/atrous.cu): the kernel's two phases (csrc/atrous_pair_core.h) are
// executed item by item on a staged tile that is filled the way the TMA load + border fix-up fill shared memory, for every
// tile of the grid. TEST INFRASTRUCTURE: it lets the CPU suite check the kernel's indexing (forward/backward pair lookup,
// aprons, ragged borders, strips) against the oracle without a GPU. Built by tests/test_atrous_emu.py with g++.
#include <cstring>
#include <vector>

#include "../../cuda-path-tracer-denoising_b200/csrc/atrous_pair_core.h"

template <class SH>
static void run_level(const float *cv, const float *gnp, const float *gzl, const float *lv, const float *kl, int W, int H, int step,
                      int row_begin, int row_end, float *out_cv) {
    const int ncg = step / SH::C, lat_w = (W + step - 1) / step, b_first = row_begin / step;
    const int lat_rows = (row_end - 1) / step - b_first + 1;
    const int tiles_x = (lat_w + SH::LX - 1) / SH::LX, tiles_y = (lat_rows + SH::LY - 1) / SH::LY;
    // one block laid out like the kernel's shared memory (cv | np | zl | lv | g): phase 1 may read two entries past a plane
    std::vector<float> smem(SH::TILE * 12 + (size_t)SH::NOFF * SH::GN);
    float4 *s_cv = reinterpret_cast<float4 *>(smem.data()), *s_np = s_cv + SH::TILE;
    float2 *s_zl = reinterpret_cast<float2 *>(s_np + SH::TILE), *s_lv = s_zl + SH::TILE;
    float *s_g = reinterpret_cast<float *>(s_lv + SH::TILE);
    for (int tile_y = 0; tile_y < tiles_y; tile_y++) for (int yc = 0; yc < step; yc++)
    for (int tile_x = 0; tile_x < tiles_x; tile_x++) for (int cg = 0; cg < ncg; cg++) {
        const int X0 = cg * SH::C, a0 = tile_x * SH::LX - 2, b0 = b_first + tile_y * SH::LY - 2;
        for (int c = 0; c < SH::C; c++) for (int tb = 0; tb < SH::SH; tb++) for (int ta = 0; ta < SH::SW; ta++) {
            const int x = X0 + (a0 + ta) * step + c, y = yc + (b0 + tb) * step, si = SH::idx(c, tb, ta);
            if (a0 + ta >= 0 && b0 + tb >= 0 && x < W && y < H) {
                const size_t q = x + (size_t)y * W;
                s_cv[si] = float4{cv[4 * q], cv[4 * q + 1], cv[4 * q + 2], cv[4 * q + 3]};
                s_np[si] = float4{gnp[4 * q], gnp[4 * q + 1], gnp[4 * q + 2], gnp[4 * q + 3]};
                s_zl[si] = float2{gzl[2 * q], gzl[2 * q + 1]}; s_lv[si] = float2{lv[2 * q], lv[2 * q + 1]};
            } else {
                s_cv[si] = float4{0, 0, 0, 0}; s_np[si] = float4{0, 0, 0, 0}; s_zl[si] = float2{0, 0}; s_lv[si] = float2{3e38f, 0};
            }
        }
        std::fill(s_g, s_g + (size_t)SH::NOFF * SH::GN, -7777.0f);    // poison: a pair phase 1 did not produce must show up
        for (int n = 0; n < SH::ITEMS; n++) pair_phase1_item<SH>(n, s_np, s_zl, s_g);
        for (int tid = 0; tid < SH::THREADS; tid++) {
            constexpr int PR = SH::PR;
            const int c = tid & 1, ap = (tid >> 1) % (SH::LX / 2), bq = tid / SH::LX;
            float k4[2][PR]; long op[2][PR]; bool live = false;
            for (int ca = 0; ca < 2; ca++) for (int cb = 0; cb < PR; cb++) {
                const int x = X0 + (a0 + 2 * ap + ca + 2) * step + c, y = yc + (b0 + PR * bq + cb + 2) * step;
                const bool ok = x < W && y >= row_begin && y < row_end;
                op[ca][cb] = ok ? x + (long)y * W : -1; k4[ca][cb] = ok ? kl[x + (size_t)y * W] : 0.f; live |= ok;
            }
            if (!live) continue;
            PairAcc A[PR];
            pair_phase2_thread<SH>(c, ap, bq, s_cv, s_lv, s_g, k4, A);
            for (int ca = 0; ca < 2; ca++) for (int cb = 0; cb < PR; cb++) {
                if (op[ca][cb] < 0) continue;
                const PairAcc &a = A[cb];
                const float w = ca ? a.w.y : a.w.x, w2 = ca ? a.w2.y : a.w2.x;
                float *o = out_cv + 4 * op[ca][cb];
                o[0] = (ca ? a.r.y : a.r.x) / w; o[1] = (ca ? a.g.y : a.g.x) / w; o[2] = (ca ? a.b.y : a.b.x) / w; o[3] = (ca ? a.v.y : a.v.x) / w2;
            }
        }
    }
}

extern "C" int pair_emu_level(const float *cv, const float *gnp, const float *gzl, const float *lv, const float *kl, int W, int H, int step,
                              int row_begin, int row_end, int shape, float *out_cv) {
    if (shape == 0) run_level<PairShape<16, 16, 2>>(cv, gnp, gzl, lv, kl, W, H, step, row_begin, row_end, out_cv);
    else if (shape == 1) run_level<PairShape<16, 12, 2>>(cv, gnp, gzl, lv, kl, W, H, step, row_begin, row_end, out_cv);
    else if (shape == 2) run_level<PairShape<16, 16, 1>>(cv, gnp, gzl, lv, kl, W, H, step, row_begin, row_end, out_cv);
    else if (shape == 3) run_level<PairShape<16, 12, 1>>(cv, gnp, gzl, lv, kl, W, H, step, row_begin, row_end, out_cv);
    else return -1;
    return 0;
}
