// Per-pixel arithmetic of the CLAHE path (K1). Every function here is `__host__ __device__` so the
// exact same source is (a) inlined into the sm_100a kernels of clahe_sm100.cu and (b) compiled by
// g++ into a test-only harness (tests/host_harness) that checks the arithmetic against live cv2 on
// a machine without a GPU. The product never runs the host instantiation.
//
// Reference behaviour restated here (OpenCV 4.13.0 float paths reached from
// mdir/components/data/transform/functional.py:35,63,148 -- see SURVEY.md App. A):
//   RGB2Lab : int16 LUT (33^3 lattice) + integer trilinear interpolation, Q14 fixed point
//   CLAHE   : bilinear blend of four tile LUTs, separate f32 roundings, cvRound
//   Lab2RGB : f32, SIMD-body / scalar-tail op sequences, cubic-spline inverse gamma
// All float ops are single IEEE-754 roundings: no FMA contraction (intrinsics on the device,
// -ffp-contract=off on the host).
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define GDT_HD __host__ __device__ __forceinline__
#else
#define GDT_HD inline
#endif

namespace gdt {

GDT_HD float f_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
GDT_HD float f_add(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
GDT_HD float f_sub(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
GDT_HD float f_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
// cvRound: round half to even
GDT_HD int f_rint(float a) {
#if defined(__CUDA_ARCH__)
    return __float2int_rn(a);
#else
    return (int)lrintf(a);
#endif
}
GDT_HD int f_floor(float a) {
#if defined(__CUDA_ARCH__)
    return __float2int_rd(a);
#else
    return (int)floorf(a);
#endif
}
GDT_HD int f_trunc(float a) {
#if defined(__CUDA_ARCH__)
    return __float2int_rz(a);
#else
    return (int)a;
#endif
}
GDT_HD float clamp01(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); }
GDT_HD float f_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}

// ---- IEEE division by a constant without the divider ---------------------------------------------
// a / b == RN-correct result of  q = a*r; { e = fma(-b, q, a); q = fma(e, r, q); } x ITERS  with r = RN(1/b)
// (Markstein's correction step: the residual e is exact, each step at least squares the error). ITERS = 2 is what
// the hardware division sequence itself runs and is correctly rounded for every normal a (b's significand not all
// ones, no under/overflow); ITERS = 1 is used only on small enumerated domains that tests/ check exhaustively.
template <int ITERS>
GDT_HD float div_by_const(float a, float b, float r) {
    float q = f_mul(a, r);
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int i = 0; i < ITERS; ++i) {
        const float e = f_fma(-b, q, a);
        q = f_fma(e, r, q);
    }
    return q;
}
// true when div_by_const<2>(a, b, RN(1/b)) is guaranteed for all a in [-4, 4] with |a| >= 2^-40 or a == 0
inline bool div_by_const_ok(float b) {
    uint32_t u;
    memcpy(&u, &b, 4);
    const uint32_t man = u & 0x7fffffu, ex = (u >> 23) & 0xffu;
    return man != 0x7fffffu && ex >= 127 - 20 && ex <= 127 + 20;
}

// ---- RGB -> Lab (Q14 integer LUT path) ----------------------------------------------------------

// Quantise a [0,1] float channel to the lattice index `t` (0..32) and the 4-bit fraction `f` (0..15).
// OpenCV: c = cvRound(x * 16384); t = c >> 9; f = (c >> 5) & 15, neighbour index clamped to 32. c == 16384 gives
// (t = 32, f = 0): the packed tables below have 33 cells per axis, the "+1" neighbours of the last cell are the clamped
// ones and carry weight 0.
GDT_HD void lab_cell(float x01, int& t, int& f) {
    const int c = f_rint(f_mul(x01, 16384.0f));
    t = c >> 9;
    f = (c >> 5) & 15;
}

// 8-bit input channel: c = cvRound((float(v) / 255.0f) * 16384) == (v * 32768 + 255) / 510 for every v in [0, 255]
// (v * 16384 / 255 is never within 1/510 of a half-integer, far outside the float rounding error), and
// t << 4 | f == c >> 5 == (v * 514 + 4) >> 8  (one multiply-add and one shift; all 256 values are checked in
// tests/test_clahe_fastmath.py). v == 255 gives t = 32, f = 0.
GDT_HD void lab_cell_u8(int v, int& t, int& f) {
    const unsigned tf = ((unsigned)v * 514u + 4u) >> 8;
    t = (int)(tf >> 4);
    f = (int)(tf & 15u);
}

// index of a lattice cell in the packed 33^3 tables
GDT_HD int lab_cell_index(int tr, int tg, int tb) { return (tr * 33 + tg) * 33 + tb; }
constexpr int kLabCells = 33 * 33 * 33;

// One channel of the trilinear interpolation. `w` holds the four (dx,dy) corner pairs of the cell,
// each 32-bit word = value(dz=0) | value(dz=1) << 16 (values in [0,16384]).
GDT_HD int lab_trilinear(uint32_t w00, uint32_t w01, uint32_t w10, uint32_t w11, int fr, int fg, int fb) {
    const int wz1 = fb, wz0 = 16 - fb;
#if defined(__CUDA_ARCH__)
    // value(dz=0) * wz0 + value(dz=1) * wz1 is one 2-way dot product of the packed 16-bit pair with two 8-bit weights
    const unsigned wz = (unsigned)wz0 | ((unsigned)wz1 << 8);
    const int i00 = (int)__dp2a_lo(w00, wz, 0u);
    const int i01 = (int)__dp2a_lo(w01, wz, 0u);
    const int i10 = (int)__dp2a_lo(w10, wz, 0u);
    const int i11 = (int)__dp2a_lo(w11, wz, 0u);
#else
    const int i00 = (int)(w00 & 0xffffu) * wz0 + (int)(w00 >> 16) * wz1;
    const int i01 = (int)(w01 & 0xffffu) * wz0 + (int)(w01 >> 16) * wz1;
    const int i10 = (int)(w10 & 0xffffu) * wz0 + (int)(w10 >> 16) * wz1;
    const int i11 = (int)(w11 & 0xffffu) * wz0 + (int)(w11 >> 16) * wz1;
#endif
    const int wx1 = fr, wx0 = 16 - fr, wy1 = fg, wy0 = 16 - fg;
    const int acc = i00 * (wx0 * wy0) + i01 * (wx0 * wy1) + i10 * (wx1 * wy0) + i11 * (wx1 * wy1);
    return (acc + 2048) >> 12;
}

// Q14 lightness -> the uint8 the reference hands to cv2 CLAHE:
//   L = (float(o0) * 2^-14) * 100 ; spc = (L + 0) / 100 ; L8 = uint8(trunc(spc * 255))
// (functional.py:35 and :148 -- truncation, not rounding).
GDT_HD int lab_l8(int o0) {
    const float L = f_mul(f_mul((float)o0, 1.0f / 16384.0f), 100.0f);
    const float spc = f_div(L, 100.0f);
    return f_trunc(f_mul(spc, 255.0f));
}

// lab_l8 without the divider: bit-identical for every o0 in [0, 16384] (exhaustively tested).
GDT_HD int lab_l8_fast(int o0) {
    const float L = f_mul(f_mul((float)o0, 1.0f / 16384.0f), 100.0f);
    const float spc = div_by_const<1>(L, 100.0f, 1.0f / 100.0f);
    return f_trunc(f_mul(spc, 255.0f));
}

// lab_l8 in integer arithmetic: trunc(((o * 2^-14) * 100 / 100) * 255) == (o * 255) >> 14 for every o in [0, 16384]
// (o * 255 / 16384 is an integer only at o = 0 and 16384, and is otherwise at least 2^-14 away from one -- four float
// ulps at 255; exhaustively tested).
GDT_HD int lab_l8_int(int o0) { return (o0 * 255) >> 14; }

// Q14 chroma -> the a (or b) value handed to LAB2RGB after the reference's normalise/denormalise
// round trip:  a = o*2^-14*256 - 128 ; spc = (a + 128) / 255 ; a' = spc*255 - 128.
GDT_HD float lab_chroma(int o) {
    const float a = f_sub(f_mul(f_mul((float)o, 1.0f / 16384.0f), 256.0f), 128.0f);
    const float spc = f_div(f_add(a, 128.0f), 255.0f);
    return f_sub(f_mul(spc, 255.0f), 128.0f);
}

// lab_chroma without the divider: (o * 2^-14) * 256 - 128 == o / 64 - 128 exactly, so (a + 128) == o / 64 exactly;
// bit-identical for every o in [0, 16384] (exhaustively tested).
GDT_HD float lab_chroma_fast(int o) {
    const float x = f_mul((float)o, 1.0f / 64.0f);
    const float spc = div_by_const<1>(x, 255.0f, 1.0f / 255.0f);
    return f_sub(f_mul(spc, 255.0f), 128.0f);
}
GDT_HD float lab_l_from_u8_fast(int v) { return f_mul(div_by_const<1>((float)v, 255.0f, 1.0f / 255.0f), 100.0f); }

// CLAHE output byte -> L handed to LAB2RGB:  (float(v) / 255) * 100 - 0.
GDT_HD float lab_l_from_u8(int v) { return f_mul(f_div((float)v, 255.0f), 100.0f); }

// ---- CLAHE bilinear blend -----------------------------------------------------------------------

struct ClaheAxis {
    int i1, i2;      // clamped tile indices
    float a, a1;     // weights (computed before clamping)
};

GDT_HD ClaheAxis clahe_axis(int coord, float inv_tile, int ntiles) {
    ClaheAxis r;
    const float tf = f_sub(f_mul((float)coord, inv_tile), 0.5f);
    const int t1 = f_floor(tf);
    r.a = f_sub(tf, (float)t1);
    r.a1 = f_sub(1.0f, r.a);
    const int t2 = t1 + 1;
    r.i1 = t1 < 0 ? 0 : t1;
    r.i2 = t2 > ntiles - 1 ? ntiles - 1 : t2;
    return r;
}

GDT_HD int clahe_blend(int l11, int l12, int l21, int l22, float xa, float xa1, float ya, float ya1) {
    const float top = f_add(f_mul((float)l11, xa1), f_mul((float)l12, xa));
    const float bot = f_add(f_mul((float)l21, xa1), f_mul((float)l22, xa));
    const float res = f_add(f_mul(top, ya1), f_mul(bot, ya));
    int v = f_rint(res);
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// ---- Lab -> RGB (f32) ---------------------------------------------------------------------------

struct Lab2RgbConst {
    float C[9];        // float32(XYZ2sRGB_D65[k][j] * whitePt[j]), row-major
};

// SIMD-body sequence, lightness half: L -> (y, fy). Depends only on the CLAHE output byte, see build_fy_table.
GDT_HD void lab_fy_body(float L, float& y, float& fy) {
    const float c16 = 16.0f / 116.0f;
    const float r903 = 1.0f / 903.3f, r116 = 1.0f / 116.0f;
    if (L <= 8.0f) {
        y = f_mul(L, r903);
        fy = f_add(f_mul(y, 7.787f), c16);
    } else {
        fy = f_mul(f_add(L, 16.0f), r116);
        y = f_mul(f_mul(fy, fy), fy);
    }
}

// SIMD-body sequence, chroma half: `y1`, `y4`, `y7` are the products C[1]*y, C[4]*y, C[7]*y.
GDT_HD void lab2lin_body_from_fy(float fy, float y1, float y4, float y7, float a, float b, const Lab2RgbConst& K,
                                 float& r, float& g, float& bl) {
    const float c16 = 16.0f / 116.0f;
    const float fth = 6.0f / 29.0f;
    const float r500 = 1.0f / 500.0f, r200 = 1.0f / 200.0f, r7787 = 1.0f / 7.787f;
    const float fx = f_add(f_mul(a, r500), fy);
    const float fz = f_sub(fy, f_mul(b, r200));
    const float X = fx <= fth ? f_mul(f_sub(fx, c16), r7787) : f_mul(f_mul(fx, fx), fx);
    const float Z = fz <= fth ? f_mul(f_sub(fz, c16), r7787) : f_mul(f_mul(fz, fz), fz);
    r = f_add(f_mul(K.C[0], X), f_add(y1, f_mul(K.C[2], Z)));
    g = f_add(f_mul(K.C[3], X), f_add(y4, f_mul(K.C[5], Z)));
    bl = f_add(f_mul(K.C[6], X), f_add(y7, f_mul(K.C[8], Z)));
}

// The same with the chroma terms already multiplied: ar = a * (1/500), bz = b * (1/200) (pass A's float chroma scratch).
GDT_HD void lab2lin_body_from_fy_pre(float fy, float y1, float y4, float y7, float ar, float bz, const Lab2RgbConst& K,
                                     float& r, float& g, float& bl) {
    const float c16 = 16.0f / 116.0f;
    const float fth = 6.0f / 29.0f;
    const float r7787 = 1.0f / 7.787f;
    const float fx = f_add(ar, fy);
    const float fz = f_sub(fy, bz);
    const float X = fx <= fth ? f_mul(f_sub(fx, c16), r7787) : f_mul(f_mul(fx, fx), fx);
    const float Z = fz <= fth ? f_mul(f_sub(fz, c16), r7787) : f_mul(f_mul(fz, fz), fz);
    r = f_add(f_mul(K.C[0], X), f_add(y1, f_mul(K.C[2], Z)));
    g = f_add(f_mul(K.C[3], X), f_add(y4, f_mul(K.C[5], Z)));
    bl = f_add(f_mul(K.C[6], X), f_add(y7, f_mul(K.C[8], Z)));
}
// Q14 chroma pair -> the two products the SIMD-body sequence adds to / subtracts from fy
GDT_HD void lab_chroma_terms(int oa, int ob, float& ar, float& bz) {
    ar = f_mul(lab_chroma_fast(oa), 1.0f / 500.0f);
    bz = f_mul(lab_chroma_fast(ob), 1.0f / 200.0f);
}

// `tail` selects OpenCV's scalar-tail sequence (last W % 8 pixels of every row): true divisions and
// ((C0*X + C1*y) + C2*Z); the SIMD body multiplies by f32 reciprocals and uses C0*X + (C1*y + C2*Z).
// Returns the three *linear* channels, unclipped.
GDT_HD void lab2lin(float L, float a, float b, bool tail, const Lab2RgbConst& K, float& r, float& g, float& bl) {
    const float c16 = 16.0f / 116.0f;
    const float fth = 6.0f / 29.0f;
    float y, fy, fx, fz, X, Z;
    if (!tail) {
        lab_fy_body(L, y, fy);
        lab2lin_body_from_fy(fy, f_mul(K.C[1], y), f_mul(K.C[4], y), f_mul(K.C[7], y), a, b, K, r, g, bl);
    } else {
        if (L <= 8.0f) {
            y = f_div(L, 903.3f);
            fy = f_add(f_mul(y, 7.787f), c16);
        } else {
            fy = f_div(f_add(L, 16.0f), 116.0f);
            y = f_mul(f_mul(fy, fy), fy);
        }
        fx = f_add(f_div(a, 500.0f), fy);
        fz = f_sub(fy, f_div(b, 200.0f));
        X = fx <= fth ? f_div(f_sub(fx, c16), 7.787f) : f_mul(f_mul(fx, fx), fx);
        Z = fz <= fth ? f_div(f_sub(fz, c16), 7.787f) : f_mul(f_mul(fz, fz), fz);
        r = f_add(f_add(f_mul(K.C[0], X), f_mul(K.C[1], y)), f_mul(K.C[2], Z));
        g = f_add(f_add(f_mul(K.C[3], X), f_mul(K.C[4], y)), f_mul(K.C[5], Z));
        bl = f_add(f_add(f_mul(K.C[6], X), f_mul(K.C[7], y)), f_mul(K.C[8], Z));
    }
}

// sRGB inverse gamma through the 1024-segment cubic spline: `seg` = {f, b, c, d} of segment ix.
GDT_HD float spline_index(float lin, int& ix) {
#if defined(__CUDA_ARCH__)
    float x = f_mul(__saturatef(lin), 1024.0f);     // == clamp01 for every non-NaN input (-0 -> +0 changes no result)
#else
    float x = f_mul(clamp01(lin), 1024.0f);
#endif
    ix = f_trunc(x);                                // x in [0, 1024]
    ix = ix > 1023 ? 1023 : ix;
    return f_sub(x, (float)ix);
}
GDT_HD float spline_eval(float x, float s0, float s1, float s2, float s3) {
    return f_add(f_mul(f_add(f_mul(f_add(f_mul(s3, x), s2), x), s1), x), s0);
}

// Normalize: (x - mean) / std, sub then true division (core_transforms.py:64-67).
GDT_HD float normalize_px(float x, float mean, float std) { return f_div(f_sub(x, mean), std); }
GDT_HD float normalize_px_fast(float x, float mean, float std, float rstd) {
    return div_by_const<2>(f_sub(x, mean), std, rstd);
}

// ---- compressed lattice record: all three channels of a cell in ONE 32-byte sector ------------------
// The trilinear sum over the 8 corners of a cell equals the multilinear polynomial through them:
//   4096 * v(fr, fg, fb) = 4096 B + 256 (D1 fr + D2 fg + D3 fb) + 16 (M12 fr fg + M13 fr fb + M23 fg fb) + M123 fr fg fb
// with B the (0,0,0) corner, D the first differences along r / g / b, M the mixed second / third differences (an exact
// integer identity: expand the weights (16 - f) and f). For the Lab lattice the D are one-signed and fit 10 bits (L, a, b
// rise or fall monotonically along each RGB axis), the mixed terms are tiny (|M| <= 52, one signed byte): 15 + 3 * 10 +
// 4 * 8 bits per channel, 231 bits per cell instead of 3 * 128. One gather of one 32-byte sector fetches what took a
// 16-byte and a 32-byte record in two sectors. pack_lab_rec32() verifies the ranges against the actual table and refuses
// otherwise.
//
// Layout: every field costs ONE instruction to extract, and the four mixed terms of a channel cost ONE dp4a.
//   w[c]     (c = 0, 1, 2 = L, a, b):  signed bytes  M12_c | M13_c | M23_c | M123_c          (dp4a against fr fg | fr fb | fg fb | 0)
//   w[3 + c]:                          |D1_c| [0,10)   |D2_c| [10,20)   |D3_c| [22,32)        (signs: lab_d_sign)
//   w[6]:                              B_L [0,15)   B_a [16,31)          w[7]:  B_b [0,15)
// The sum is accumulated times 4 (acc4 = 4 * 4096 * v + rounding) so that the middle field can be used where it lies:
// (w & (0x3ff << 10)) = 1024 |D2| is exactly its term per unit of fg; low and top fields take the weights 1024 fr / fb.
// sign of D1 (r), D2 (g), D3 (b) for L, a, b: a falls with green, b falls with blue
GDT_HD constexpr int lab_d_sign(int c, int k) { return ((c == 1 && k == 1) || (c == 2 && k == 2)) ? -1 : 1; }

struct LabWeights {      // per pixel, shared by the three channels
    int r1024, g1, b1024;    // 1024 fr, fg, 1024 fb
    uint32_t m;              // bytes fr fg | fr fb | fg fb | 0   (each <= 225)
    int rgb4;                // 4 fr fg fb
};
GDT_HD LabWeights lab_weights(int fr, int fg, int fb) {
    LabWeights W;
    W.r1024 = fr * 1024; W.g1 = fg; W.b1024 = fb * 1024;
    const int rg = fr * fg, rb = fr * fb, gb = fg * fb;
    W.m = (uint32_t)rg | ((uint32_t)rb << 8) | ((uint32_t)gb << 16);
    W.rgb4 = rg * (fb * 4);
    return W;
}
// signed bytes of `a` times unsigned bytes of `b`, summed, plus c
GDT_HD int dp4a_su(uint32_t a, uint32_t b, int c) {
#if defined(__CUDA_ARCH__)
    int d;
    asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    for (int i = 0; i < 4; ++i) c += (int)(int8_t)(a >> (8 * i)) * (int)((b >> (8 * i)) & 255u);
    return c;
#endif
}
// -> Q14 (L, a, b) of the pixel, identical to lab_trilinear() on the uncompressed corners
GDT_HD void lab_from_rec32(const uint32_t* w, const LabWeights& W, int& oL, int& oa, int& ob) {
    const int base[3] = {(int)(w[6] & 0xffffu), (int)(w[6] >> 16), (int)w[7]};
    int o[3];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int c = 0; c < 3; ++c) {
        const uint32_t wm = w[c], wd = w[3 + c];
        int a = base[c] * 16384 + 8192;                                   // 4 * (4096 B + 2048)
        a += lab_d_sign(c, 0) * ((int)(wd & 0x3ffu) * W.r1024);
        a += lab_d_sign(c, 1) * ((int)(wd & (0x3ffu << 10)) * W.g1);
        a += lab_d_sign(c, 2) * ((int)(wd >> 22) * W.b1024);
        a += dp4a_su(wm, W.m, 0) * 64;
        a += ((int)wm >> 24) * W.rgb4;
        o[c] = a >> 14;
    }
    oL = o[0]; oa = o[1]; ob = o[2];
}

// lut33: the [33][33][33][3] int16 table. rec: 33^3 * 8 words. Returns false (rec partially written) when a difference does
// not fit its field or has the wrong sign: the caller then keeps the uncompressed records.
inline bool pack_lab_rec32(const int16_t* lut33, uint32_t* rec) {
    auto at = [&](int r, int g, int b, int c) -> int {
        r = r < 32 ? r : 32; g = g < 32 ? g : 32; b = b < 32 ? b : 32;   // clamped neighbours carry weight 0
        return lut33[((r * 33 + g) * 33 + b) * 3 + c];
    };
    for (int tr = 0; tr < 33; ++tr)
        for (int tg = 0; tg < 33; ++tg)
            for (int tb = 0; tb < 33; ++tb) {
                uint32_t* w = rec + (size_t)((tr * 33 + tg) * 33 + tb) * 8;
                for (int i = 0; i < 8; ++i) w[i] = 0u;
                for (int c = 0; c < 3; ++c) {
                    const int v000 = at(tr, tg, tb, c), v100 = at(tr + 1, tg, tb, c), v010 = at(tr, tg + 1, tb, c),
                              v001 = at(tr, tg, tb + 1, c), v110 = at(tr + 1, tg + 1, tb, c),
                              v101 = at(tr + 1, tg, tb + 1, c), v011 = at(tr, tg + 1, tb + 1, c),
                              v111 = at(tr + 1, tg + 1, tb + 1, c);
                    const int D[3] = {v100 - v000, v010 - v000, v001 - v000};
                    const int M[4] = {v110 - v100 - v010 + v000, v101 - v100 - v001 + v000, v011 - v010 - v001 + v000,
                                      v111 - v110 - v101 - v011 + v100 + v010 + v001 - v000};
                    if (v000 < 0 || v000 >= (1 << 15)) return false;
                    uint32_t mag[3];
                    for (int k = 0; k < 3; ++k) {
                        const int m = D[k] * lab_d_sign(c, k);
                        if (m < 0 || m >= (1 << 10)) return false;
                        mag[k] = (uint32_t)m;
                    }
                    for (int k = 0; k < 4; ++k) {
                        if (M[k] < -128 || M[k] > 127) return false;
                        w[c] |= ((uint32_t)M[k] & 255u) << (8 * k);
                    }
                    w[3 + c] = mag[0] | (mag[1] << 10) | (mag[2] << 22);
                    if (c == 0) w[6] |= (uint32_t)v000;
                    else if (c == 1) w[6] |= (uint32_t)v000 << 16;
                    else w[7] = (uint32_t)v000;
                }
            }
    return true;
}

// ---- two pixels per instruction --------------------------------------------------------------------
// sm_100 has packed fp32 arithmetic (mul / add / fma .rn.f32x2: SASS FMUL2 / FADD2 / FFMA2): both lanes round once per
// operation exactly like the scalar instruction, so a sequence written lane-wise stays bit-identical while it takes half
// the issue slots. Every *2 function below is the scalar function of the same name applied to two pixels.
#if defined(__CUDACC__)
typedef float2 f2;
#else
struct f2 { float x, y; };
#endif
GDT_HD f2 mk2(float a, float b) { f2 r; r.x = a; r.y = b; return r; }
GDT_HD f2 bc2(float a) { return mk2(a, a); }
GDT_HD f2 p_mul(f2 a, f2 b) {
#if defined(__CUDA_ARCH__)
    return __fmul2_rn(a, b);
#else
    return mk2(a.x * b.x, a.y * b.y);
#endif
}
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (it does not for the scalar .rn forms), which would
// drop a rounding. The packed addition is therefore issued as fma(a, 1, b) with a 1.0 the assembler cannot see through
// (a __constant__ it must assume the host may rewrite): exact product, one rounding of a + b, nothing left to contract.
#if defined(__CUDACC__)
__constant__ float2 k_opaque_one2 = {1.0f, 1.0f};
#endif
GDT_HD f2 p_add(f2 a, f2 b) {
#if defined(__CUDA_ARCH__)
    return __ffma2_rn(a, k_opaque_one2, b);
#else
    return mk2(a.x + b.x, a.y + b.y);
#endif
}
GDT_HD f2 p_fma(f2 a, f2 b, f2 c) {
#if defined(__CUDA_ARCH__)
    return __ffma2_rn(a, b, c);
#else
    return mk2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
// a - b as fma(b, -1, a): the product is exact, one rounding of the exact difference (signed zeros included)
GDT_HD f2 p_sub(f2 a, f2 b) { return p_fma(b, bc2(-1.0f), a); }
GDT_HD f2 p_sel(bool cx, bool cy, f2 a, f2 b) { return mk2(cx ? a.x : b.x, cy ? a.y : b.y); }

template <int ITERS>
GDT_HD f2 div_by_const2(f2 a, float b, float r) {
    f2 q = p_mul(a, bc2(r));
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int i = 0; i < ITERS; ++i) {
        const f2 e = p_fma(bc2(-b), q, a);
        q = p_fma(e, bc2(r), q);
    }
    return q;
}
GDT_HD f2 lab_chroma_fast2(int o0, int o1) {
    const f2 x = p_mul(mk2((float)o0, (float)o1), bc2(1.0f / 64.0f));
    const f2 spc = div_by_const2<1>(x, 255.0f, 1.0f / 255.0f);
    return p_add(p_mul(spc, bc2(255.0f)), bc2(-128.0f));
}
GDT_HD f2 lab_l_from_u8_fast2(int v0, int v1) {
    return p_mul(div_by_const2<1>(mk2((float)v0, (float)v1), 255.0f, 1.0f / 255.0f), bc2(100.0f));
}
GDT_HD void clahe_blend2(const int* l11, const int* l12, const int* l21, const int* l22, f2 xa, f2 xa1, float ya, float ya1,
                         int& d0, int& d1) {
    const f2 top = p_add(p_mul(mk2((float)l11[0], (float)l11[1]), xa1), p_mul(mk2((float)l12[0], (float)l12[1]), xa));
    const f2 bot = p_add(p_mul(mk2((float)l21[0], (float)l21[1]), xa1), p_mul(mk2((float)l22[0], (float)l22[1]), xa));
    const f2 res = p_add(p_mul(top, bc2(ya1)), p_mul(bot, bc2(ya)));
    d0 = f_rint(res.x); d1 = f_rint(res.y);
    d0 = d0 < 0 ? 0 : (d0 > 255 ? 255 : d0);
    d1 = d1 < 0 ? 0 : (d1 > 255 ? 255 : d1);
}
// lab_fy_body for two pixels: both branches are computed lane-wise, the scalar function's branch is selected per lane
GDT_HD void lab_fy_body2(f2 L, f2& y, f2& fy) {
    const float c16 = 16.0f / 116.0f;
    const float r903 = 1.0f / 903.3f, r116 = 1.0f / 116.0f;
    const f2 ylo = p_mul(L, bc2(r903));
    const f2 fylo = p_add(p_mul(ylo, bc2(7.787f)), bc2(c16));
    const f2 fyhi = p_mul(p_add(L, bc2(16.0f)), bc2(r116));
    const f2 yhi = p_mul(p_mul(fyhi, fyhi), fyhi);
    const bool lx = L.x <= 8.0f, ly = L.y <= 8.0f;
    y = p_sel(lx, ly, ylo, yhi);
    fy = p_sel(lx, ly, fylo, fyhi);
}
GDT_HD void lab2lin_body2(f2 fy, f2 y, f2 a, f2 b, const Lab2RgbConst& K, f2& r, f2& g, f2& bl) {
    const float c16 = 16.0f / 116.0f;
    const float fth = 6.0f / 29.0f;
    const float r500 = 1.0f / 500.0f, r200 = 1.0f / 200.0f, r7787 = 1.0f / 7.787f;
    const f2 fx = p_add(p_mul(a, bc2(r500)), fy);
    const f2 fz = p_sub(fy, p_mul(b, bc2(r200)));
    const f2 X = p_sel(fx.x <= fth, fx.y <= fth, p_mul(p_add(fx, bc2(-c16)), bc2(r7787)), p_mul(p_mul(fx, fx), fx));
    const f2 Z = p_sel(fz.x <= fth, fz.y <= fth, p_mul(p_add(fz, bc2(-c16)), bc2(r7787)), p_mul(p_mul(fz, fz), fz));
    r = p_add(p_mul(bc2(K.C[0]), X), p_add(p_mul(bc2(K.C[1]), y), p_mul(bc2(K.C[2]), Z)));
    g = p_add(p_mul(bc2(K.C[3]), X), p_add(p_mul(bc2(K.C[4]), y), p_mul(bc2(K.C[5]), Z)));
    bl = p_add(p_mul(bc2(K.C[6]), X), p_add(p_mul(bc2(K.C[7]), y), p_mul(bc2(K.C[8]), Z)));
}
GDT_HD f2 spline_index2(f2 lin, int& ix0, int& ix1) {
#if defined(__CUDA_ARCH__)
    const f2 x = p_mul(mk2(__saturatef(lin.x), __saturatef(lin.y)), bc2(1024.0f));
#else
    const f2 x = p_mul(mk2(clamp01(lin.x), clamp01(lin.y)), bc2(1024.0f));
#endif
    ix0 = f_trunc(x.x); ix1 = f_trunc(x.y);
    ix0 = ix0 > 1023 ? 1023 : ix0;
    ix1 = ix1 > 1023 ? 1023 : ix1;
    return p_sub(x, mk2((float)ix0, (float)ix1));
}
GDT_HD f2 spline_eval2(f2 x, f2 s0, f2 s1, f2 s2, f2 s3) {
    return p_add(p_mul(p_add(p_mul(p_add(p_mul(s3, x), s2), x), s1), x), s0);
}
GDT_HD f2 normalize_px_fast2(f2 x, float mean, float std, float rstd) {
    return div_by_const2<2>(p_add(x, bc2(-mean)), std, rstd);
}

// ---- host-side table builders (used by gdt_init; plain C++) --------------------------------------

// Natural cubic spline of the sRGB inverse gamma, built exactly like OpenCV's splineBuild on
// float32 samples f[i] evaluated in double (color_lab.cpp). tab: 1024*4 floats.
inline void build_inv_gamma_spline(float* tab) {
    const int n = 1024;
    static float f[1025];
    for (int i = 0; i <= n; ++i) {
        const double xi = (double)i / n;
        f[i] = (float)(xi <= 0.0031308 ? xi * 12.92 : 1.055 * pow(xi, 1.0 / 2.4) - 0.055);
    }
    tab[0] = tab[1] = 0.0f;
    for (int i = 1; i < n; ++i) {
        volatile float t0 = f[i] * 2.0f;
        volatile float t1 = f[i + 1] - t0;
        volatile float t2 = t1 + f[i - 1];
        volatile float t = t2 * 3.0f;
        volatile float den = 4.0f - tab[(i - 1) * 4];
        volatile float l = 1.0f / den;
        tab[i * 4] = l;
        volatile float num = t - tab[(i - 1) * 4 + 1];
        tab[i * 4 + 1] = num * l;
    }
    float cn = 0.0f;
    for (int i = n - 1; i >= 0; --i) {
        volatile float p0 = tab[i * 4] * cn;
        volatile float c = tab[i * 4 + 1] - p0;
        volatile float d0 = f[i + 1] - f[i];
        volatile float c2 = c * 2.0f;
        volatile float s0 = cn + c2;
        volatile float s1 = s0 / 3.0f;
        volatile float b = d0 - s1;
        volatile float e0 = cn - c;
        volatile float d = e0 / 3.0f;
        tab[i * 4] = f[i];
        tab[i * 4 + 1] = b;
        tab[i * 4 + 2] = c;
        tab[i * 4 + 3] = d;
        cn = c;
    }
}

inline void build_lab2rgb_const(Lab2RgbConst& K) {
    static const double M[9] = {3.240479, -1.53715, -0.498535, -0.969256, 1.875991, 0.041556,
                                0.055648, -0.204043, 1.057311};
    static const double wp[3] = {0.950456, 1.0, 1.088754};
    for (int k = 0; k < 3; ++k)
        for (int j = 0; j < 3; ++j) K.C[k * 3 + j] = (float)(M[k * 3 + j] * wp[j]);
}

// Lightness half of Lab->RGB per CLAHE output byte v (SIMD-body sequence): {fy, C[1]*y, C[4]*y, C[7]*y} of
// L = (float(v) / 255) * 100. tab: 256 * 4 floats. Built with the same functions the kernels inline.
inline void build_fy_table(const Lab2RgbConst& K, float* tab) {
    for (int v = 0; v < 256; ++v) {
        float y, fy;
        lab_fy_body(lab_l_from_u8(v), y, fy);
        tab[v * 4 + 0] = fy;
        tab[v * 4 + 1] = f_mul(K.C[1], y);
        tab[v * 4 + 2] = f_mul(K.C[4], y);
        tab[v * 4 + 3] = f_mul(K.C[7], y);
    }
}

// Re-pack the 33^3 x 3 lattice table into per-cell words (see lab_trilinear):
//   lutL [cell][dxdy]          4 words = 16 B per cell   (lightness)
//   lutAB[cell][ch(a,b)][dxdy] 8 words = 32 B per cell   (chroma)
// cell = lab_cell_index(tr, tg, tb), tr/tg/tb in 0..32; neighbour indices are clamped to 32 (weight 0 there).
inline void pack_lab_lut(const int16_t* lut33, uint32_t* lutL, uint32_t* lutAB) {
    for (int tr = 0; tr < 33; ++tr)
        for (int tg = 0; tg < 33; ++tg)
            for (int tb = 0; tb < 33; ++tb) {
                const int cell = lab_cell_index(tr, tg, tb);
                const int tb1 = tb < 32 ? tb + 1 : 32;
                for (int dx = 0; dx < 2; ++dx)
                    for (int dy = 0; dy < 2; ++dy) {
                        const int r1 = tr + dx < 32 ? tr + dx : 32, g1 = tg + dy < 32 ? tg + dy : 32;
                        const int16_t* p0 = lut33 + ((r1 * 33 + g1) * 33 + tb) * 3;
                        const int16_t* p1 = lut33 + ((r1 * 33 + g1) * 33 + tb1) * 3;
                        const int q = dx * 2 + dy;
                        lutL[cell * 4 + q] = (uint32_t)(uint16_t)p0[0] | ((uint32_t)(uint16_t)p1[0] << 16);
                        lutAB[cell * 8 + q] = (uint32_t)(uint16_t)p0[1] | ((uint32_t)(uint16_t)p1[1] << 16);
                        lutAB[cell * 8 + 4 + q] = (uint32_t)(uint16_t)p0[2] | ((uint32_t)(uint16_t)p1[2] << 16);
                    }
            }
}

}  // namespace gdt
