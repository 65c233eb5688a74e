// tile_emu.cpp -- host emulation of atrous_tiled_kernel (csrc/atrous.cu), the production a-trous kernel: every thread's work
// (csrc/atrous_tile_core.h: at_thread_compute -- the very code the kernel runs) on tiles staged the way the TMA load + border
// fix-up stage them, for every tile of the grid and every tile shape of the kernel's table. TEST INFRASTRUCTURE (CPU suite).
#include <vector>

#include "../../cuda-path-tracer-denoising_b200/csrc/atrous_tile_core.h"

template <int LX, int LY, int TY>
static void run_level(const float *cv, const float *gnp, const float *gzl, const float *lv, const float *kl, int W, int H, int step,
                      int row_begin, int row_end, float *out_cv) {
    using SH = AtShape<LX, LY, TY>;
    static_assert(SH::OK, "tile shape");
    const int ncg = step / AT_C, lat_w = (W + step - 1) / step, b_first = row_begin / step;
    const int lat_rows = (row_end - 1) / step - b_first + 1;
    const int tiles_x = (lat_w + LX - 1) / LX, tiles_y = (lat_rows + LY - 1) / LY;
    std::vector<float4> s_cv(SH::TILEP), s_np(SH::TILEP);
    std::vector<float2> s_zl(SH::TILEP), s_lv(SH::TILEP);
    for (int tile_y = 0; tile_y < tiles_y; tile_y++) for (int yc = 0; yc < step; yc++)
    for (int tile_x = 0; tile_x < tiles_x; tile_x++) for (int cg = 0; cg < ncg; cg++) {
        const int X0 = cg * AT_C, a0 = tile_x * LX - 2, b0 = b_first + tile_y * LY - 2;
        for (int c = 0; c < AT_C; c++) for (int tb = 0; tb < SH::SH; tb++) for (int ta = 0; ta < SH::SW; ta++) {
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
        for (int tid = 0; tid < SH::THREADS; tid++) {
            const int c = tid & 1, ap = (tid >> 1) % (LX / 2), bq = tid / LX;
            float k4[AT_TX][TY]; long op[AT_TX][TY]; bool live = false;
            for (int ca = 0; ca < AT_TX; ca++) for (int cb = 0; cb < TY; cb++) {
                const int x = X0 + (a0 + 2 * ap + ca + 2) * step + c, y = yc + (b0 + TY * bq + cb + 2) * step;
                const bool ok = x < W && y >= row_begin && y < row_end;
                op[ca][cb] = ok ? x + (long)y * W : -1; k4[ca][cb] = ok ? kl[x + (size_t)y * W] : 0.f; live |= ok;
            }
            if (!live) continue;
            AtAcc2 A[TY];
            at_thread_compute<SH>(c, ap, bq, s_cv.data(), s_np.data(), s_zl.data(), s_lv.data(), k4, A);
            for (int ca = 0; ca < AT_TX; ca++) for (int cb = 0; cb < TY; cb++) {
                if (op[ca][cb] < 0) continue;
                const AtAcc2 &a = A[cb];
                const float w = ca ? a.w.y : a.w.x, w2 = ca ? a.w2.y : a.w2.x;
                float *o = out_cv + 4 * op[ca][cb];
                o[0] = (ca ? a.r.y : a.r.x) / w; o[1] = (ca ? a.g.y : a.g.x) / w; o[2] = (ca ? a.b.y : a.b.x) / w; o[3] = (ca ? a.v.y : a.v.x) / w2;
            }
        }
    }
}

// `shape` indexes the same table as g_at_shapes in csrc/atrous.cu
extern "C" int tile_emu_level(const float *cv, const float *gnp, const float *gzl, const float *lv, const float *kl, int W, int H, int step,
                              int row_begin, int row_end, int shape, float *out_cv) {
#define RUN(LX, LY, TY) run_level<LX, LY, TY>(cv, gnp, gzl, lv, kl, W, H, step, row_begin, row_end, out_cv); return 0
    switch (shape) {
        case 0: RUN(16, 32, 4); case 1: RUN(32, 16, 4); case 2: RUN(16, 16, 2); case 3: RUN(16, 32, 2); case 4: RUN(32, 16, 2);
        case 5: RUN(16, 24, 4); case 6: RUN(32, 12, 4); case 7: RUN(16, 32, 2); case 8: RUN(32, 8, 2); case 9: RUN(16, 12, 2);
        case 10: RUN(32, 12, 2); case 11: RUN(16, 16, 2); case 12: RUN(16, 16, 2); case 13: RUN(16, 18, 3); case 14: RUN(16, 16, 1);
    }
    return -1;
}
