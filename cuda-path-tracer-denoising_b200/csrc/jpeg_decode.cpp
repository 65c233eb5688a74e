// jpeg_decode.cpp -- SURVEY.md 8(f) N2, last piece: textures of the reference's scenes are JPEG files which its loader decodes
// with the vendored stb_image (src/sceneStructs.h:198-199, stbi_load(file, &w, &h, &comp, 0)). The pixels enter the image through
// Texture::getColor, so "the same scene" means THE SAME BYTES: this is a JPEG decoder (ITU T.81 baseline and progressive Huffman,
// 8-bit, 1 or 3 components) whose arithmetic after the entropy decoder -- the 12-bit fixed-point inverse DCT, the chroma
// up-sampling filter of 4:2:0 images (3/4, 1/4 taps in both directions), the 20-bit fixed-point YCbCr -> RGB conversion with
// its truncated green cross term -- is the arithmetic stb_image uses, rounding included. tests/test_scene_ingest.py compares its
// output byte for byte with the pixels the reference loader produced for both textures of the repo's scenes (one progressive
// 4:4:4 file with an Adobe marker, one baseline 4:2:0 file). Host only.
#include <cstdint>
#include <cstring>
#include <new>
#include <string>
#include <vector>

namespace {

const uint8_t kZig[64 + 15] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
                               6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
                               39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
                               63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};      // a corrupt run may step past 63

struct Huff {
    // canonical code: for every length, the first code of that length and the index of its first symbol
    int mincode[17], maxcode[18], valptr[17];
    uint8_t vals[256];
    bool build(const uint8_t *counts, const uint8_t *v, int nv) {
        memcpy(vals, v, nv);
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            valptr[l] = k; mincode[l] = code;
            code += counts[l - 1]; k += counts[l - 1];
            maxcode[l] = counts[l - 1] ? code - 1 : -1;
            if (code > (1 << l)) return false;
            code <<= 1;
        }
        return k == nv;
    }
};

struct Comp {
    int id, h, v, tq, td, ta;
    int x, y, w2, h2;               // size in samples; allocated size (whole MCUs)
    int dc_pred;
    std::vector<uint8_t> data;
    std::vector<short> coeff;       // progressive: w2/8 x h2/8 blocks of 64
    int coeff_w, coeff_h;
};

struct Dec {
    const uint8_t *p, *end;
    std::string err;
    int W = 0, H = 0, ncomp = 0, progressive = 0;
    int hmax = 1, vmax = 1, mcu_w = 0, mcu_h = 0, mcu_x = 0, mcu_y = 0;
    uint16_t dequant[4][64];
    Huff hdc[4], hac[4];
    bool have_dc[4] = {false, false, false, false}, have_ac[4] = {false, false, false, false};
    Comp comp[4];
    int restart_interval = 0;
    int jfif = 0, app14_transform = -1;
    // scan
    int order[4], scan_n = 0, spec_start = 0, spec_end = 0, succ_high = 0, succ_low = 0, eob_run = 0;
    // bit reader
    uint32_t code_buffer = 0; int code_bits = 0; bool nomore = false; int marker = 0xff;

    bool fail(const char *m) { if (err.empty()) err = m; return false; }
    int get8() { return p < end ? *p++ : 0; }
    int get16() { const int a = get8(); return (a << 8) | get8(); }

    void grow() {
        do {
            unsigned b = nomore ? 0u : (unsigned)get8();
            if (b == 0xff) {
                int c = get8();
                while (c == 0xff) c = get8();
                if (c != 0) { marker = c; nomore = true; b = 0; }       // past a marker the stream reads as zero bits
            }
            code_buffer |= b << (24 - code_bits);
            code_bits += 8;
        } while (code_bits <= 24);
    }
    int getbits(int n) {
        if (n == 0) return 0;
        if (code_bits < n) grow();
        const uint32_t k = code_buffer >> (32 - n);
        code_buffer <<= n; code_bits -= n;
        return (int)k;
    }
    int getbit() { return getbits(1); }
    // n bits as the signed value of the JPEG "extend" procedure
    int extend_receive(int n) {
        if (n == 0) return 0;
        const int v = getbits(n);
        return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
    }
    int huff_decode(const Huff &h) {
        int code = 0;
        for (int l = 1; l <= 16; l++) {
            code = (code << 1) | getbit();
            if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
        }
        fail("bad huffman code");
        return -1;
    }
    void reset_scan() {
        code_bits = 0; code_buffer = 0; nomore = false; marker = 0xff; eob_run = 0;
        for (int i = 0; i < 4; i++) comp[i].dc_pred = 0;
    }

    // ---- entropy decoding of one block ----
    bool block_baseline(short data[64], const Huff &dc, const Huff &ac, int c, const uint16_t *dq) {
        int t = huff_decode(dc);
        if (t < 0 || t > 15) return fail("bad DC code");
        memset(data, 0, 64 * sizeof(short));
        const int diff = t ? extend_receive(t) : 0;
        const int d = (int)((uint32_t)comp[c].dc_pred + (uint32_t)diff);       // wraps on a damaged file, never overflows
        comp[c].dc_pred = d;
        data[0] = (short)((uint32_t)d * dq[0]);
        int k = 1;
        do {
            const int rs = huff_decode(ac);
            if (rs < 0) return false;
            const int s = rs & 15, r = rs >> 4;
            if (s == 0) {
                if (rs != 0xf0) break;      // end of block
                k += 16;
            } else {
                k += r;
                const int zig = kZig[k++];
                data[zig] = (short)(extend_receive(s) * dq[zig]);
            }
        } while (k < 64);
        return true;
    }
    bool block_prog_dc(short data[64], const Huff &dc, int c) {
        if (spec_end != 0) return fail("cannot merge DC and AC");
        if (succ_high == 0) {       // first scan for the DC coefficient
            memset(data, 0, 64 * sizeof(short));
            const int t = huff_decode(dc);
            if (t < 0 || t > 15) return fail("bad DC code");
            const int diff = t ? extend_receive(t) : 0;
            const int d = (int)((uint32_t)comp[c].dc_pred + (uint32_t)diff);
            comp[c].dc_pred = d;
            data[0] = (short)((uint32_t)d << succ_low);
        } else if (getbit()) {      // refinement: one more bit
            data[0] += (short)(1 << succ_low);
        }
        return true;
    }
    bool block_prog_ac(short data[64], const Huff &ac) {
        if (spec_start == 0) return fail("cannot merge DC and AC");
        if (succ_high == 0) {
            const int shift = succ_low;
            if (eob_run) { --eob_run; return true; }
            int k = spec_start;
            do {
                const int rs = huff_decode(ac);
                if (rs < 0) return false;
                const int s = rs & 15, r = rs >> 4;
                if (s == 0) {
                    if (r < 15) {
                        eob_run = 1 << r;
                        if (r) eob_run += getbits(r);
                        --eob_run;
                        break;
                    }
                    k += 16;
                } else {
                    k += r;
                    const int zig = kZig[k++];
                    data[zig] = (short)(extend_receive(s) * (1 << shift));
                }
            } while (k <= spec_end);
        } else {                    // refinement scan
            const short bit = (short)(1 << succ_low);
            if (eob_run) {
                --eob_run;
                for (int k = spec_start; k <= spec_end; ++k) {
                    short *q = &data[kZig[k]];
                    if (*q != 0 && getbit() && (*q & bit) == 0) { if (*q > 0) *q += bit; else *q -= bit; }
                }
            } else {
                int k = spec_start;
                do {
                    const int rs = huff_decode(ac);
                    if (rs < 0) return false;
                    int s = rs & 15, r = rs >> 4;
                    if (s == 0) {
                        if (r < 15) {
                            eob_run = (1 << r) - 1;
                            if (r) eob_run += getbits(r);
                            r = 64;     // force the end of the block
                        }
                        // r == 15: a run of 16 zero coefficients, handled by the loop below
                    } else {
                        if (s != 1) return fail("bad huffman code");
                        s = getbit() ? bit : -bit;
                    }
                    while (k <= spec_end) {     // advance by r zero-history coefficients, refining the others on the way
                        short *q = &data[kZig[k++]];
                        if (*q != 0) {
                            if (getbit() && (*q & bit) == 0) { if (*q > 0) *q += bit; else *q -= bit; }
                        } else {
                            if (r == 0) { *q = (short)s; break; }
                            --r;
                        }
                    }
                } while (k <= spec_end);
            }
        }
        return true;
    }

    // ---- inverse DCT: 12-bit fixed point, constants and rounding of stb_image's stbi__idct_block ----
    static inline int f2f(float x) { return (int)(x * 4096 + 0.5); }
    static inline uint8_t clamp8(int x) { return (unsigned)x > 255 ? (x < 0 ? 0 : 255) : (uint8_t)x; }
    // All sums and products wrap modulo 2^32 (what stb_image's int arithmetic does on every real compiler): a damaged
    // file may hold coefficients that overflow, and the result must be defined and the same everywhere.
    struct W32 {
        uint32_t u;
        W32() : u(0) {}
        W32(int x) : u((uint32_t)x) {}
        friend W32 operator+(W32 a, W32 b) { W32 r; r.u = a.u + b.u; return r; }
        friend W32 operator-(W32 a, W32 b) { W32 r; r.u = a.u - b.u; return r; }
        friend W32 operator*(W32 a, W32 b) { W32 r; r.u = a.u * b.u; return r; }
        W32 &operator+=(W32 b) { u += b.u; return *this; }
        int sar(int n) const { return (int)(int32_t)u >> n; }       // arithmetic shift of the two's-complement value
    };
    static void idct(uint8_t *out, int stride, const short d[64]) {
        W32 val[64], *v = val;
#define SVGF_IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)                                                                          \
    W32 t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3;                                                                   \
    p2 = s2; p3 = s6;                                                                                                         \
    p1 = (p2 + p3) * W32(f2f(0.5411961f));                                                                                    \
    t2 = p1 + p3 * W32(f2f(-1.847759065f));                                                                                   \
    t3 = p1 + p2 * W32(f2f(0.765366865f));                                                                                    \
    p2 = s0; p3 = s4;                                                                                                         \
    t0 = (p2 + p3) * W32(4096); t1 = (p2 - p3) * W32(4096);                                                                   \
    x0 = t0 + t3; x3 = t0 - t3; x1 = t1 + t2; x2 = t1 - t2;                                                                   \
    t0 = s7; t1 = s5; t2 = s3; t3 = s1;                                                                                       \
    p3 = t0 + t2; p4 = t1 + t3; p1 = t0 + t3; p2 = t1 + t2;                                                                   \
    p5 = (p3 + p4) * W32(f2f(1.175875602f));                                                                                  \
    t0 = t0 * W32(f2f(0.298631336f)); t1 = t1 * W32(f2f(2.053119869f));                                                       \
    t2 = t2 * W32(f2f(3.072711026f)); t3 = t3 * W32(f2f(1.501321110f));                                                       \
    p1 = p5 + p1 * W32(f2f(-0.899976223f)); p2 = p5 + p2 * W32(f2f(-2.562915447f));                                           \
    p3 = p3 * W32(f2f(-1.961570560f)); p4 = p4 * W32(f2f(-0.390180644f));                                                     \
    t3 += p1 + p4; t2 += p2 + p3; t1 += p2 + p4; t0 += p1 + p3;
        const short *dd = d;
        for (int i = 0; i < 8; ++i, ++dd, ++v) {        // columns
            if (dd[8] == 0 && dd[16] == 0 && dd[24] == 0 && dd[32] == 0 && dd[40] == 0 && dd[48] == 0 && dd[56] == 0) {
                const W32 dcterm = W32(dd[0]) * W32(4);
                v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dcterm;
            } else {
                SVGF_IDCT_1D(W32(dd[0]), W32(dd[8]), W32(dd[16]), W32(dd[24]), W32(dd[32]), W32(dd[40]), W32(dd[48]), W32(dd[56]))
                const W32 r(512);
                x0 += r; x1 += r; x2 += r; x3 += r;
                v[0] = (x0 + t3).sar(10); v[56] = (x0 - t3).sar(10);
                v[8] = (x1 + t2).sar(10); v[48] = (x1 - t2).sar(10);
                v[16] = (x2 + t1).sar(10); v[40] = (x2 - t1).sar(10);
                v[24] = (x3 + t0).sar(10); v[32] = (x3 - t0).sar(10);
            }
        }
        v = val;
        uint8_t *o = out;
        for (int i = 0; i < 8; ++i, v += 8, o += stride) {      // rows; the rounding constant also re-centres the samples on 128
            SVGF_IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
            const W32 r(65536 + (128 << 17));
            x0 += r; x1 += r; x2 += r; x3 += r;
            o[0] = clamp8((x0 + t3).sar(17)); o[7] = clamp8((x0 - t3).sar(17));
            o[1] = clamp8((x1 + t2).sar(17)); o[6] = clamp8((x1 - t2).sar(17));
            o[2] = clamp8((x2 + t1).sar(17)); o[5] = clamp8((x2 - t1).sar(17));
            o[3] = clamp8((x3 + t0).sar(17)); o[4] = clamp8((x3 - t0).sar(17));
        }
#undef SVGF_IDCT_1D
    }

    // ---- markers ----
    bool process_marker(int m) {
        switch (m) {
        case 0xDD:
            if (get16() != 4) return fail("bad DRI length");
            restart_interval = get16();
            return true;
        case 0xDB: {
            int L = get16() - 2;
            while (L > 0) {
                const int q = get8(), prec = q >> 4, t = q & 15;
                if (prec > 1 || t > 3) return fail("bad DQT");
                for (int i = 0; i < 64; ++i) dequant[t][kZig[i]] = (uint16_t)(prec ? get16() : get8());
                L -= prec ? 129 : 65;
            }
            return L == 0 ? true : fail("bad DQT length");
        }
        case 0xC4: {
            int L = get16() - 2;
            while (L > 0) {
                uint8_t counts[16]; uint8_t vals[256];
                const int q = get8(), tc = q >> 4, th = q & 15;
                if (tc > 1 || th > 3) return fail("bad DHT header");
                int n = 0;
                for (int i = 0; i < 16; ++i) { counts[i] = (uint8_t)get8(); n += counts[i]; }
                if (n > 256) return fail("bad DHT header");
                L -= 17;
                for (int i = 0; i < n; ++i) vals[i] = (uint8_t)get8();
                L -= n;
                Huff &h = tc ? hac[th] : hdc[th];
                if (!h.build(counts, vals, n)) return fail("bad code lengths");
                (tc ? have_ac : have_dc)[th] = true;
            }
            return L == 0 ? true : fail("bad DHT length");
        }
        }
        if ((m >= 0xE0 && m <= 0xEF) || m == 0xFE) {        // APPn, COM
            int L = get16();
            if (L < 2) return fail(m == 0xFE ? "bad COM length" : "bad APP length");
            L -= 2;
            if (m == 0xE0 && L >= 5) {
                static const char tag[5] = {'J', 'F', 'I', 'F', '\0'};
                bool ok = true;
                for (int i = 0; i < 5; ++i) if (get8() != tag[i]) ok = false;
                L -= 5;
                if (ok) jfif = 1;
            } else if (m == 0xEE && L >= 12) {
                static const char tag[6] = {'A', 'd', 'o', 'b', 'e', '\0'};
                bool ok = true;
                for (int i = 0; i < 6; ++i) if (get8() != tag[i]) ok = false;
                L -= 6;
                if (ok) { get8(); get16(); get16(); app14_transform = get8(); L -= 6; }
            }
            p += L; if (p > end) p = end;
            return true;
        }
        return fail("unknown marker");
    }
    int next_marker() {
        if (marker != 0xff) { const int m = marker; marker = 0xff; return m; }
        int x = get8();
        if (x != 0xff) return 0xff;
        while (x == 0xff) x = get8();
        return x;
    }
    bool frame_header() {
        const int Lf = get16();
        if (Lf < 11) return fail("bad SOF length");
        if (get8() != 8) return fail("only 8-bit JPEG");
        H = get16(); W = get16();
        if (H <= 0 || W <= 0) return fail("empty image");
        if ((size_t)W * (size_t)H > ((size_t)1 << 28)) return fail("image too large");
        ncomp = get8();
        if (ncomp != 1 && ncomp != 3) return fail("1 or 3 components only");
        if (Lf != 8 + 3 * ncomp) return fail("bad SOF length");
        for (int i = 0; i < ncomp; ++i) {
            Comp &c = comp[i];
            c.id = get8();
            const int q = get8();
            c.h = q >> 4; c.v = q & 15; c.tq = get8();
            if (!c.h || c.h > 4 || !c.v || c.v > 4 || c.tq > 3) return fail("bad component");
            if (c.h > hmax) hmax = c.h;
            if (c.v > vmax) vmax = c.v;
        }
        for (int i = 0; i < ncomp; ++i) if (hmax % comp[i].h || vmax % comp[i].v) return fail("bad sampling factors");
        mcu_w = hmax * 8; mcu_h = vmax * 8;
        mcu_x = (W + mcu_w - 1) / mcu_w; mcu_y = (H + mcu_h - 1) / mcu_h;
        for (int i = 0; i < ncomp; ++i) {
            Comp &c = comp[i];
            c.x = (W * c.h + hmax - 1) / hmax; c.y = (H * c.v + vmax - 1) / vmax;
            c.w2 = mcu_x * c.h * 8; c.h2 = mcu_y * c.v * 8;
            c.data.assign((size_t)c.w2 * c.h2, 0);
            if (progressive) { c.coeff_w = c.w2 / 8; c.coeff_h = c.h2 / 8; c.coeff.assign((size_t)c.w2 * c.h2, 0); }
        }
        return true;
    }
    bool scan_header() {
        const int Ls = get16();
        scan_n = get8();
        if (scan_n < 1 || scan_n > 4 || scan_n > ncomp) return fail("bad SOS component count");
        if (Ls != 6 + 2 * scan_n) return fail("bad SOS length");
        for (int i = 0; i < scan_n; ++i) {
            const int id = get8(), q = get8();
            int which = 0;
            for (; which < ncomp; ++which) if (comp[which].id == id) break;
            if (which == ncomp) return false;
            comp[which].td = q >> 4; comp[which].ta = q & 15;
            if (comp[which].td > 3 || comp[which].ta > 3) return fail("bad table index");
            order[i] = which;
        }
        spec_start = get8(); spec_end = get8();
        const int aa = get8();
        succ_high = aa >> 4; succ_low = aa & 15;
        if (progressive) {
            if (spec_start > 63 || spec_end > 63 || spec_start > spec_end || succ_high > 13 || succ_low > 13) return fail("bad SOS");
        } else {
            if (spec_start != 0 || succ_high != 0 || succ_low != 0) return fail("bad SOS");
            spec_end = 63;
        }
        return true;
    }
    bool restart_check(int &todo) {
        if (--todo > 0) return true;
        if (code_bits < 24) grow();
        if (!(marker >= 0xD0 && marker <= 0xD7)) { todo = 0x7fffffff; return false; }       // no restart marker: the scan is over
        reset_scan();
        todo = restart_interval ? restart_interval : 0x7fffffff;
        return true;
    }
    bool entropy_scan() {
        reset_scan();
        int todo = restart_interval ? restart_interval : 0x7fffffff;
        if (!progressive) {
            short data[64];
            if (scan_n == 1) {      // non-interleaved: blocks of the component in raster order, only those that hold samples
                const int n = order[0];
                Comp &c = comp[n];
                const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
                for (int j = 0; j < h; ++j)
                    for (int i = 0; i < w; ++i) {
                        if (!have_dc[c.td] || !have_ac[c.ta]) return fail("missing huffman table");
                        if (!block_baseline(data, hdc[c.td], hac[c.ta], n, dequant[c.tq])) return false;
                        idct(&c.data[(size_t)c.w2 * j * 8 + i * 8], c.w2, data);
                        if (!restart_check(todo)) return true;
                    }
            } else {
                for (int j = 0; j < mcu_y; ++j)
                    for (int i = 0; i < mcu_x; ++i) {
                        for (int k = 0; k < scan_n; ++k) {
                            const int n = order[k];
                            Comp &c = comp[n];
                            if (!have_dc[c.td] || !have_ac[c.ta]) return fail("missing huffman table");
                            for (int y = 0; y < c.v; ++y)
                                for (int x = 0; x < c.h; ++x) {
                                    const int x2 = (i * c.h + x) * 8, y2 = (j * c.v + y) * 8;
                                    if (!block_baseline(data, hdc[c.td], hac[c.ta], n, dequant[c.tq])) return false;
                                    idct(&c.data[(size_t)c.w2 * y2 + x2], c.w2, data);
                                }
                        }
                        if (!restart_check(todo)) return true;
                    }
            }
            return true;
        }
        if (scan_n == 1) {
            const int n = order[0];
            Comp &c = comp[n];
            const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
            for (int j = 0; j < h; ++j)
                for (int i = 0; i < w; ++i) {
                    short *data = &c.coeff[64 * (size_t)(i + j * c.coeff_w)];
                    if (spec_start == 0) {
                        if (!have_dc[c.td]) return fail("missing huffman table");
                        if (!block_prog_dc(data, hdc[c.td], n)) return false;
                    } else {
                        if (!have_ac[c.ta]) return fail("missing huffman table");
                        if (!block_prog_ac(data, hac[c.ta])) return false;
                    }
                    if (!restart_check(todo)) return true;
                }
        } else {        // interleaved progressive scans carry DC only
            for (int j = 0; j < mcu_y; ++j)
                for (int i = 0; i < mcu_x; ++i) {
                    for (int k = 0; k < scan_n; ++k) {
                        const int n = order[k];
                        Comp &c = comp[n];
                        if (!have_dc[c.td]) return fail("missing huffman table");
                        for (int y = 0; y < c.v; ++y)
                            for (int x = 0; x < c.h; ++x) {
                                const int x2 = i * c.h + x, y2 = j * c.v + y;
                                if (!block_prog_dc(&c.coeff[64 * (size_t)(x2 + y2 * c.coeff_w)], hdc[c.td], n)) return false;
                            }
                    }
                    if (!restart_check(todo)) return true;
                }
        }
        return true;
    }
    void finish_progressive() {
        for (int n = 0; n < ncomp; ++n) {
            Comp &c = comp[n];
            const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
            for (int j = 0; j < h; ++j)
                for (int i = 0; i < w; ++i) {
                    short *data = &c.coeff[64 * (size_t)(i + j * c.coeff_w)];
                    for (int q = 0; q < 64; ++q) data[q] = (short)(data[q] * dequant[c.tq][q]);
                    idct(&c.data[(size_t)c.w2 * j * 8 + i * 8], c.w2, data);
                }
        }
    }

    bool decode() {
        if (get8() != 0xff || get8() != 0xD8) return fail("no SOI");
        int m = next_marker();
        while (!(m == 0xC0 || m == 0xC1 || m == 0xC2)) {
            if (!process_marker(m)) return false;
            m = next_marker();
            while (m == 0xff) { if (p >= end) return fail("no SOF"); m = next_marker(); }
        }
        progressive = m == 0xC2;
        if (!frame_header()) return false;
        m = next_marker();
        while (m != 0xD9) {
            if (m == 0xDA) {
                if (!scan_header()) return fail(err.empty() ? "bad SOS" : err.c_str());
                if (!entropy_scan()) return false;
                if (marker == 0xff) {       // the entropy decoder did not run into the next marker: skip what is left of the scan
                    while (p < end) {
                        const int x = get8();
                        if (x == 0xff) {
                            const int y = get8();
                            if (y != 0 && y != 0xff) { marker = y; break; }
                        }
                    }
                }
            } else if (m == 0xDC) {
                get16(); get16();       // DNL: ignored
            } else if (!process_marker(m)) {
                return false;
            }
            if (p >= end && marker == 0xff) break;
            m = next_marker();
        }
        if (progressive) finish_progressive();
        return true;
    }
};

// chroma up-sampling, the four cases stb_image distinguishes (output: one row of `w * hs` samples)
const uint8_t *resample_1(uint8_t *, const uint8_t *near, const uint8_t *, int, int) { return near; }
const uint8_t *resample_v2(uint8_t *out, const uint8_t *near, const uint8_t *far, int w, int) {
    for (int i = 0; i < w; ++i) out[i] = (uint8_t)((3 * near[i] + far[i] + 2) >> 2);
    return out;
}
const uint8_t *resample_h2(uint8_t *out, const uint8_t *in, const uint8_t *, int w, int) {
    if (w == 1) { out[0] = out[1] = in[0]; return out; }
    out[0] = in[0];
    out[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
    int i;
    for (i = 1; i < w - 1; ++i) {
        const int n = 3 * in[i] + 2;
        out[i * 2] = (uint8_t)((n + in[i - 1]) >> 2);
        out[i * 2 + 1] = (uint8_t)((n + in[i + 1]) >> 2);
    }
    out[i * 2] = (uint8_t)((in[w - 2] * 3 + in[w - 1] + 2) >> 2);
    out[i * 2 + 1] = in[w - 1];
    return out;
}
const uint8_t *resample_hv2(uint8_t *out, const uint8_t *near, const uint8_t *far, int w, int) {
    if (w == 1) { out[0] = out[1] = (uint8_t)((3 * near[0] + far[0] + 2) >> 2); return out; }
    int t1 = 3 * near[0] + far[0];
    out[0] = (uint8_t)((t1 + 2) >> 2);
    for (int i = 1; i < w; ++i) {
        const int t0 = t1;
        t1 = 3 * near[i] + far[i];
        out[i * 2 - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
        out[i * 2] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
    }
    out[w * 2 - 1] = (uint8_t)((t1 + 2) >> 2);
    return out;
}
const uint8_t *resample_generic(uint8_t *out, const uint8_t *near, const uint8_t *, int w, int hs) {
    for (int i = 0; i < w; ++i) for (int j = 0; j < hs; ++j) out[i * hs + j] = near[i];
    return out;
}

inline int float2fixed(float x) { return ((int)(x * 4096.0f + 0.5f)) << 8; }

}  // namespace

// Decodes a JPEG held in memory into 8-bit pixels, `*components` per pixel (1: grey, 3: RGB), row-major, top row first -- what
// stbi_load(file, &w, &h, &comp, 0) returns for the same file.
bool svgf_jpeg_decode(const unsigned char *bytes, size_t n, int *width, int *height, int *components, std::vector<unsigned char> &pixels, std::string &err) {
    Dec d;
    d.p = bytes; d.end = bytes + n;
    bool ok = false;
    try { ok = d.decode(); } catch (const std::bad_alloc &) { d.err = "out of memory"; }
    if (!ok) { err = d.err.empty() ? "corrupt JPEG" : d.err; return false; }
    const int W = d.W, H = d.H, nc = d.ncomp;
    pixels.assign((size_t)W * H * nc, 0);
    typedef const uint8_t *(*resample_fn)(uint8_t *, const uint8_t *, const uint8_t *, int, int);
    struct R { resample_fn fn; const uint8_t *line0, *line1; int hs, vs, w_lores, ystep, ypos; std::vector<uint8_t> buf; } r[3];
    for (int k = 0; k < nc; ++k) {
        const Comp &c = d.comp[k];
        r[k].hs = d.hmax / c.h; r[k].vs = d.vmax / c.v;
        r[k].ystep = r[k].vs >> 1; r[k].w_lores = (W + r[k].hs - 1) / r[k].hs; r[k].ypos = 0;
        r[k].line0 = r[k].line1 = c.data.data();
        r[k].buf.assign((size_t)W + 3 + 16, 0);
        r[k].fn = (r[k].hs == 1 && r[k].vs == 1) ? resample_1 : (r[k].hs == 1 && r[k].vs == 2) ? resample_v2 :
                  (r[k].hs == 2 && r[k].vs == 1) ? resample_h2 : (r[k].hs == 2 && r[k].vs == 2) ? resample_hv2 : resample_generic;
    }
    int rgb_ids = 0;
    for (int k = 0; k < nc; ++k) { static const char rgb[3] = {'R', 'G', 'B'}; if (nc == 3 && d.comp[k].id == rgb[k]) rgb_ids++; }
    const bool is_rgb = nc == 3 && (rgb_ids == 3 || (d.app14_transform == 0 && !d.jfif));
    for (int j = 0; j < H; ++j) {
        const uint8_t *co[3] = {nullptr, nullptr, nullptr};
        for (int k = 0; k < nc; ++k) {
            R &q = r[k];
            const bool y_bot = q.ystep >= (q.vs >> 1);
            co[k] = q.fn(q.buf.data(), y_bot ? q.line1 : q.line0, y_bot ? q.line0 : q.line1, q.w_lores, q.hs);
            if (++q.ystep >= q.vs) {
                q.ystep = 0;
                q.line0 = q.line1;
                if (++q.ypos < d.comp[k].y) q.line1 += d.comp[k].w2;
            }
        }
        uint8_t *out = &pixels[(size_t)nc * W * j];
        if (nc == 1) { memcpy(out, co[0], W); continue; }
        if (is_rgb) {
            for (int i = 0; i < W; ++i) { out[0] = co[0][i]; out[1] = co[1][i]; out[2] = co[2][i]; out += 3; }
            continue;
        }
        for (int i = 0; i < W; ++i) {       // 20-bit fixed point; the green cross term is truncated to 16 bits before the add
            const int y_fixed = (co[0][i] << 20) + (1 << 19);
            const int cb = co[1][i] - 128, cr = co[2][i] - 128;
            int rr = y_fixed + cr * float2fixed(1.40200f);
            int gg = y_fixed + (cr * -float2fixed(0.71414f)) + ((cb * -float2fixed(0.34414f)) & 0xffff0000);
            int bb = y_fixed + cb * float2fixed(1.77200f);
            rr >>= 20; gg >>= 20; bb >>= 20;
            if ((unsigned)rr > 255) rr = rr < 0 ? 0 : 255;
            if ((unsigned)gg > 255) gg = gg < 0 ? 0 : 255;
            if ((unsigned)bb > 255) bb = bb < 0 ? 0 : 255;
            out[0] = (uint8_t)rr; out[1] = (uint8_t)gg; out[2] = (uint8_t)bb;
            out += 3;
        }
    }
    *width = W; *height = H; *components = nc;
    return true;
}
