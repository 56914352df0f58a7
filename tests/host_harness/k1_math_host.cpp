// Test-only host instantiation of gandtr_b200/csrc/clahe_math.cuh (the arithmetic the sm_100a kernels inline), built by
// tests/test_host_k1_math.py with g++ -ffp-contract=off and compared with oracle/clahe_np.py. Never part of the product.
#include <stdint.h>
#include <vector>

#include "../../gandtr_b200/csrc/clahe_math.cuh"

using namespace gdt;

static std::vector<uint32_t> g_L, g_AB, g_rec;
static bool g_rec_ok = false;
static Lab2RgbConst g_K;
static float g_spline[4096], g_fy[1024];

extern "C" void k1h_init(const int16_t* lut33) {
    g_L.assign((size_t)kLabCells * 4, 0u);
    g_AB.assign((size_t)kLabCells * 8, 0u);
    pack_lab_lut(lut33, g_L.data(), g_AB.data());
    g_rec.assign((size_t)kLabCells * 8, 0u);
    g_rec_ok = pack_lab_rec32(lut33, g_rec.data());
    build_lab2rgb_const(g_K);
    build_inv_gamma_spline(g_spline);
    build_fy_table(g_K, g_fy);
}

// pass A of one uint8 pixel: -> CLAHE input byte and packed Q14 chroma (a | b << 16)
extern "C" void k1h_stage_a_u8(const uint8_t* rgb, long n, uint8_t* l8, uint32_t* ab) {
    for (long i = 0; i < n; ++i) {
        int tr, tg, tb, fr, fg, fb;
        lab_cell_u8(rgb[i * 3], tr, fr);
        lab_cell_u8(rgb[i * 3 + 1], tg, fg);
        lab_cell_u8(rgb[i * 3 + 2], tb, fb);
        const uint32_t* wl = &g_L[(size_t)lab_cell_index(tr, tg, tb) * 4];
        const uint32_t* wc = &g_AB[(size_t)lab_cell_index(tr, tg, tb) * 8];
        l8[i] = (uint8_t)lab_l8_int(lab_trilinear(wl[0], wl[1], wl[2], wl[3], fr, fg, fb));
        const int oa = lab_trilinear(wc[0], wc[1], wc[2], wc[3], fr, fg, fb);
        const int ob = lab_trilinear(wc[4], wc[5], wc[6], wc[7], fr, fg, fb);
        ab[i] = (uint32_t)oa | ((uint32_t)ob << 16);
    }
}

// the same for float pixels already in [0,1] (ClahePost path)
extern "C" void k1h_stage_a_f32(const float* rgb, long n, uint8_t* l8, uint32_t* ab) {
    for (long i = 0; i < n; ++i) {
        int tr, tg, tb, fr, fg, fb;
        lab_cell(clamp01(rgb[i * 3]), tr, fr);
        lab_cell(clamp01(rgb[i * 3 + 1]), tg, fg);
        lab_cell(clamp01(rgb[i * 3 + 2]), tb, fb);
        const uint32_t* wl = &g_L[(size_t)lab_cell_index(tr, tg, tb) * 4];
        const uint32_t* wc = &g_AB[(size_t)lab_cell_index(tr, tg, tb) * 8];
        l8[i] = (uint8_t)lab_l8_int(lab_trilinear(wl[0], wl[1], wl[2], wl[3], fr, fg, fb));
        const int oa = lab_trilinear(wc[0], wc[1], wc[2], wc[3], fr, fg, fb);
        const int ob = lab_trilinear(wc[4], wc[5], wc[6], wc[7], fr, fg, fb);
        ab[i] = (uint32_t)oa | ((uint32_t)ob << 16);
    }
}

// pass B after the CLAHE blend: (CLAHE output byte, packed chroma) -> normalised RGB, SIMD-body sequence with the
// lightness half taken from the 256-entry table (use_table) or recomputed, or OpenCV's scalar-tail sequence (tail)
extern "C" void k1h_stage_b(const uint8_t* dst, const uint32_t* ab, long n, int use_table, int tail, const float* mean,
                            const float* std_, float* out_rgb) {
    float rstd[3];
    for (int c = 0; c < 3; ++c) { volatile float r = 1.0f / std_[c]; rstd[c] = r; }
    if (use_table == 2 && !tail) {
        // the packed two-pixel sequence (sm_100 f32x2 on the device), pixels (i, i + 1); an odd last pixel pairs with itself
        for (long i = 0; i < n; i += 2) {
            const long j = i + 1 < n ? i + 1 : i;
            const f2 A = lab_chroma_fast2((int)(ab[i] & 0xffffu), (int)(ab[j] & 0xffffu));
            const f2 B = lab_chroma_fast2((int)(ab[i] >> 16), (int)(ab[j] >> 16));
            f2 y, fy, lin[3];
            lab_fy_body2(lab_l_from_u8_fast2(dst[i], dst[j]), y, fy);
            lab2lin_body2(fy, y, A, B, g_K, lin[0], lin[1], lin[2]);
            for (int c = 0; c < 3; ++c) {
                int i0, i1;
                const f2 x = spline_index2(lin[c], i0, i1);
                const float *s = &g_spline[i0 * 4], *t = &g_spline[i1 * 4];
                const f2 e = spline_eval2(x, mk2(s[0], t[0]), mk2(s[1], t[1]), mk2(s[2], t[2]), mk2(s[3], t[3]));
                const f2 o = normalize_px_fast2(e, mean[c], std_[c], rstd[c]);
                out_rgb[i * 3 + c] = o.x;
                out_rgb[j * 3 + c] = o.y;
            }
        }
        return;
    }
    for (long i = 0; i < n; ++i) {
        const float a2 = lab_chroma_fast((int)(ab[i] & 0xffffu)), b2 = lab_chroma_fast((int)(ab[i] >> 16));
        float lin[3];
        if (use_table && !tail) {
            const float* t = &g_fy[dst[i] * 4];
            lab2lin_body_from_fy(t[0], t[1], t[2], t[3], a2, b2, g_K, lin[0], lin[1], lin[2]);
        } else {
            lab2lin(lab_l_from_u8_fast(dst[i]), a2, b2, tail != 0, g_K, lin[0], lin[1], lin[2]);
        }
        for (int c = 0; c < 3; ++c) {
            int ix;
            const float x = spline_index(lin[c], ix);
            const float e = spline_eval(x, g_spline[ix * 4], g_spline[ix * 4 + 1], g_spline[ix * 4 + 2], g_spline[ix * 4 + 3]);
            out_rgb[i * 3 + c] = normalize_px_fast(e, mean[c], std_[c], rstd[c]);
        }
    }
}

// Compressed 32-byte lattice record against the uncompressed corner records: every cell, every (fr, fg, fb) in
// [0, 16)^3 with stride `fstep` in each fraction. Returns the number of mismatching (cell, fraction, channel) triples,
// -1 when the table did not fit the record format.
extern "C" long k1h_rec32_check(int fstep) {
    if (!g_rec_ok) return -1;
    long bad = 0;
    for (int cell = 0; cell < kLabCells; ++cell) {
        const uint32_t* wl = &g_L[(size_t)cell * 4];
        const uint32_t* wc = &g_AB[(size_t)cell * 8];
        const uint32_t* w = &g_rec[(size_t)cell * 8];
        for (int fr = 0; fr < 16; fr += fstep)
            for (int fg = 0; fg < 16; fg += fstep)
                for (int fb = 0; fb < 16; fb += fstep) {
                    int oL, oa, ob;
                    lab_from_rec32(w, lab_weights(fr, fg, fb), oL, oa, ob);
                    bad += oL != lab_trilinear(wl[0], wl[1], wl[2], wl[3], fr, fg, fb);
                    bad += oa != lab_trilinear(wc[0], wc[1], wc[2], wc[3], fr, fg, fb);
                    bad += ob != lab_trilinear(wc[4], wc[5], wc[6], wc[7], fr, fg, fb);
                }
    }
    return bad;
}
