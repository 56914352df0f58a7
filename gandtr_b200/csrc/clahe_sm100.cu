// K1 -- CLAHE preprocessing for sm_100a: fused `pil2np | apply_clahe | totensor | normalize`
// (mdir/components/data/transform/{core_transforms.py:35-100, photometric_transforms.py:28-36,
// functional.py:28-35,55-63,81-85,140-161}) and the ClahePost wrapper
// (mdir/components/data/wrapper.py:325-348), bit-exact against the reference's OpenCV 4.13.0 path.
//
// Two launches per batch, both streaming with coalesced 16-byte accesses (defaults):
//   pass A  clahe_hist_kernel   one CTA per (image, tile): RGB -> lattice cell + 4-bit fractions (integer arithmetic, one
//           multiply-shift per channel) -> ONE 32-byte gather of the cell's compressed record (clahe_math.cuh: base + first
//           differences + mixed differences of L, a, b) -> Q14 L, a, b by the multilinear form -> uint8 L8 scratch (integer
//           formula) + 4-byte chroma scratch (a | b << 16) + 256-bin shared-memory histogram (bank-skewed copies per warp)
//           -> clip, redistribute, prefix sum -> tile LUT (transposed rows).
//   pass B  clahe_apply_kernel  persistent, one 1024-thread CTA per SM = four 256-thread groups walking over (image, row
//           band, 1024-px column chunk) items: LUT rows of the band staged per item; the inverse-gamma spline and the
//           lightness table held in shared memory eight times (conflict-free lookups); per pixel: bilinear LUT blend ->
//           Lab->RGB -> spline inverse gamma -> normalise -> planar float4 stores. No lattice access, no texture.
// Algorithmic HBM bytes per pixel: 3 in + 12 out (u8 variant), 12 + 12 (f32 variant). Scratch: 5 B/px written by A and
// read by B (the input itself is read and quantised once).
// Round 1's arrangement (pass A interpolates the lightness only from a 16-byte record and stores a cell code, pass B
// fetches the 32-byte chroma record through the texture pipe; non-persistent pass B) is compiled in, bit-identical, and
// selected automatically when the lattice does not fit the compressed record; g_k1_* / gdt_debug_k1_* switch every
// variant at run time for A/B timing (profiles/README.md).
#include <stdlib.h>

#include "clahe_math.cuh"
#include "common.cuh"

namespace gdt {

struct ClaheTables {
    uint4* lutL = nullptr;     // [33^3]      packed lightness corners
    uint4* lutAB = nullptr;    // [33^3][2]   packed chroma corners (a words, b words)
    uint32_t* rec32 = nullptr; // [33^3][8]   compressed record of all three channels (clahe_math.cuh), when it fits
    bool rec_ok = false;
    // std values for which div_by_const<1> (ONE Markstein correction) equals IEEE division for every numerator K1 can
    // produce: checked exhaustively on this device at gdt_init (div1_check_kernel)
    float div1_std[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int n_div1 = 0;
    bool div1_verified(float s) const {
        for (int i = 0; i < n_div1; ++i)
            if (div1_std[i] == s) return true;
        return false;
    }
    float4* spline = nullptr;  // [1024]
    float4* fytab = nullptr;   // [256]       {fy, C1*y, C4*y, C7*y} per CLAHE output byte (build_fy_table)
    cudaTextureObject_t texL = 0, texAB = 0, texSpline = 0, texFy = 0;   // the same tables behind the texture path
    Lab2RgbConst K;
    float spline_host[4096];
    float fy_host[1024];
    bool ready = false;
};

static ClaheTables g_tables[32];

// A/B switches (gdt_debug_k1_config): texab bit 0 = pass A fetches the chroma lattice records through the texture pipe
// (when chroma_a), bit 1 / bit 2 = pass A fetches all / every other lightness record through the texture pipe (when
// !chroma_a);
// spltex = 0..1 spline lookups of pass B through the texture pipe; fytex = lightness half of Lab->RGB from the 256-entry
// table (texture pipe) instead of recomputing it. Every combination is bit-identical; only the pipe balance differs.
//   chroma_a = interpolate the chroma in pass A (one lattice visit per pixel) instead of pass B (gather hidden under
//   pass B's arithmetic); occ_a = resident CTAs per SM pass A is compiled for (4 or 6).
// Defaults = the fastest combination measured on B200 (profiles/k1_v2_ab_r1q.log): chroma in pass B (its gather hides
// under pass B's arithmetic; in pass A it is exposed: 1.03 vs 1.20 ms per 128 images), everything else recomputed.
static int g_k1_texab = 0, g_k1_spltex = 0, g_k1_fytex = 0, g_k1_occ_a = 4;
// chroma_a: -1 = automatic: pass A interpolates all three channels from ONE 32-byte compressed lattice record per pixel
// (a single sector gather; pass B then touches no lattice at all) whenever the table fits that format, else 0.
// 1 with g_k1_rec32 == 0 is round 1's uncompressed three-gather variant.
static int g_k1_chroma_a = -1;
static int g_k1_rec32 = 1;
static int g_k1_chroma_f = 0; // 1: pass A hands pass B the chroma as float terms (8 B/px scratch; gdt_debug_k1_chroma_f). Measured
                              // SLOWER (0.999 vs 0.901 ms per 128 images, profiles/k1_chroma_f_ab_r2ae.log): pass A sits on the
                              // L1 data pipe, the wider stores land exactly there
static int g_k1_div1 = 1;     // one-correction-step normalisation for the std values verified at gdt_init (gdt_debug_k1_div1)
static int g_k1_pack = 0;     // pass B: packed f32x2 arithmetic (two pixels per instruction); 0 = scalar (gdt_debug_k1_pack)
static int g_k1_persist = 1;  // pass B: persistent 1024-thread CTAs with conflict-free spline copies (gdt_debug_k1_persist)
static int g_k1_rows = 0;     // > 0: rows per pass-B CTA forced (gdt_debug_k1_rows), 0: pass_b_rows()
static int g_k1_chunk = 0;    // images per (pass A, pass B) launch pair: 0 whole batch at once (default: measured fastest,
                              // profiles/k1_chunk_ab_r2a.log), -1 sized so that a chunk's scratch stays in L2, > 0 forced
                              // (gdt_debug_k1_chunk)

const ClaheTables* clahe_tables_for_current_device() {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return nullptr;
    return g_tables[dev].ready ? &g_tables[dev] : nullptr;
}

struct Norm3 {
    float mean[3];
    float std[3];
};

// ---- pixel helpers ------------------------------------------------------------------------------

// One 8-bit pixel -> (cell, fr, fg, fb), pure integer arithmetic (no table)
__device__ __forceinline__ void cell_from_u8(int r, int g, int b, int& cell, int& fr, int& fg, int& fb) {
    int tr, tg, tb;
    lab_cell_u8(r, tr, fr);
    lab_cell_u8(g, tg, fg);
    lab_cell_u8(b, tb, fb);
    cell = lab_cell_index(tr, tg, tb);
}

__device__ __forceinline__ void cell_from_f32(float r, float g, float b, const Norm3& in, int& cell, int& fr, int& fg,
                                              int& fb) {
    // ClahePost: tensor.mul(std).add(mean) (wrapper.py:342), then cv2 clips to [0,1]
    int tr, tg, tb;
    lab_cell(clamp01(f_add(f_mul(r, in.std[0]), in.mean[0])), tr, fr);
    lab_cell(clamp01(f_add(f_mul(g, in.std[1]), in.mean[1])), tg, fg);
    lab_cell(clamp01(f_add(f_mul(b, in.std[2]), in.mean[2])), tb, fb);
    cell = lab_cell_index(tr, tg, tb);
}

// Per-pixel "cell code" (chroma interpolated in pass B): bits [0,16) lattice cell, [16,20) fr, [20,24) fg, [24,28) fb
__device__ __forceinline__ uint32_t pack_code(int cell, int fr, int fg, int fb) {
    return (uint32_t)cell | ((uint32_t)fr << 16) | ((uint32_t)fg << 20) | ((uint32_t)fb << 24);
}
__device__ __forceinline__ void unpack_code(uint32_t c, int& cell, int& fr, int& fg, int& fb) {
    cell = (int)(c & 0xffffu);
    fr = (int)((c >> 16) & 15u);
    fg = (int)((c >> 20) & 15u);
    fb = (int)(c >> 24);
}

// Per-pixel scratch written by pass A and consumed by pass B: the CLAHE input byte and EITHER the two Q14 chroma channels
// packed as a | b << 16 (each in [0, 16384]; CHROMA_A: the lattice is interpolated once, in pass A) OR the cell code
// (pass B fetches the chroma records itself, overlapping the gather with its arithmetic).
__device__ __forceinline__ void lab_from_records(const uint4& wl, const uint4& wa, const uint4& wb, int fr, int fg, int fb,
                                                 int& l8, uint32_t& ab) {
    l8 = lab_l8_int(lab_trilinear(wl.x, wl.y, wl.z, wl.w, fr, fg, fb));
    const int oa = lab_trilinear(wa.x, wa.y, wa.z, wa.w, fr, fg, fb);
    const int ob = lab_trilinear(wb.x, wb.y, wb.z, wb.w, fr, fg, fb);
    ab = (uint32_t)oa | ((uint32_t)ob << 16);
}

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// Shared-memory histogram increment. Neighbouring pixels of a tile share few grey levels, so a single histogram would
// serialise same-address atomics inside a warp: every warp owns kHistCopies copies (lane & 3 picks one; the copy stride
// is padded by 8 words so equal levels in different copies fall into different banks). `key` > 255 means "no pixel".
constexpr int kHistCopies = 4;
constexpr int kHistStride = 256 + 8;
__device__ __forceinline__ void hist_add(int* my_hist, int key) {
    if (key < 256) atomicAdd(&my_hist[key], 1);
}

// ---- pass A -------------------------------------------------------------------------------------
// tile LUT layout written by pass A and read by pass B: lutT[img][ty][v][tx] with 8 or 16 bytes per (ty, v) row, so
// the four tile LUT values a pixel blends sit in two 16-byte rows (one per tile row).
// Row stride: 8 bytes for grids up to 8x8 (the default), 16 bytes up to 16x16 -- `lsh` = log2(stride).
__host__ __device__ inline int lut_row_shift(int grid) { return grid <= 8 ? 3 : 4; }

// 32-byte gather of one chroma cell record (a words then b words) as a single 256-bit read-only load
__device__ __forceinline__ void ld_cell_ab(const uint4* __restrict__ lutAB, int cell, uint4& wa, uint4& wb) {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(wa.x), "=r"(wa.y), "=r"(wa.z), "=r"(wa.w), "=r"(wb.x), "=r"(wb.y), "=r"(wb.z), "=r"(wb.w)
                 : "l"(lutAB + cell * 2));
}

// the 32-byte compressed record of a cell (one sector) as a single 256-bit read-only load
__device__ __forceinline__ void ld_cell_rec32(const uint32_t* __restrict__ rec32, int cell, uint32_t* w) {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(rec32 + (size_t)cell * 8));
}
__device__ __forceinline__ void lab_from_rec32_px(const uint32_t* w, int fr, int fg, int fb, int& l8, uint32_t& ab) {
    int oL, oa, ob;
    lab_from_rec32(w, lab_weights(fr, fg, fb), oL, oa, ob);
    l8 = lab_l8_int(oL);
    ab = (uint32_t)oa | ((uint32_t)ob << 16);
}

// `gq`, `gr` = 256 / gw, 256 % gw (gw = 4-pixel groups per tile row): the vectorised loop walks (row, group) incrementally,
// no division per step. TEXAB: chroma records through the texture pipe (idle otherwise), lightness through the LSU pipe.
// TEXL: which lightness-record gathers take the texture pipe instead of the LSU pipe (pass A is bound by the LSU pipe's
// scattered 16-byte gathers): 0 none, 1 all, 2 every other pixel (both pipes gather in parallel).
// REC32 (with CHROMA_A): all three channels from the compressed 32-byte record, one gather per pixel.
// VEC1: the host guarantees vec_ok == 1 (every tile a whole number of aligned 4-pixel groups inside the image: the common
// sizes), so the ragged / reflected / generic paths and their per-group tests are compiled out.
// CHROMA_F (VEC1, REC32, widths without scalar-tail pixels): the chroma leaves pass A as the two float terms of Lab->RGB
// (8 B/px: a / 500 | b / 200 in the SIMD-body operation order) instead of the Q14 pair; see clahe_apply_kernel.
template <bool U8, bool TEXAB, bool CHROMA_A, int MINB, int TEXL, bool REC32, bool VEC1, bool CHROMA_F>
__global__ void __launch_bounds__(256, MINB)
clahe_hist_kernel(const void* __restrict__ in_, uint8_t* __restrict__ L8, uint32_t* __restrict__ AB,
                  uint8_t* __restrict__ lutT, int h, int w, int pitch,
                  int grid, int th, int tw, int clip, float lut_scale, int vec_ok_, int gq, int gr,
                  const uint4* __restrict__ lutL, const uint4* __restrict__ lutAB, Norm3 in_norm,
                  cudaTextureObject_t texAB, cudaTextureObject_t texL, const uint32_t* __restrict__ rec32) {
    __shared__ int hist_all[8 * kHistCopies * kHistStride];
    __shared__ int warp_tmp[8];
    const int tid = threadIdx.x;
    const int img = blockIdx.y;
    const int ty = blockIdx.x / grid, tx = blockIdx.x % grid;
    for (int i = tid; i < 8 * kHistCopies * kHistStride; i += 256) hist_all[i] = 0;
    int* hist = hist_all + ((tid >> 5) * kHistCopies + (tid & (kHistCopies - 1))) * kHistStride;
    __syncthreads();

    const size_t plane = (size_t)h * w;
    const uint8_t* in8 = (const uint8_t*)in_ + (size_t)img * plane * 3;
    const float* inf = (const float*)in_ + (size_t)img * plane * 3;
    // scratch rows are `pitch` = align4(w) elements apart: 4-pixel groups are 16-byte aligned for every width
    uint8_t* l8img = L8 + (size_t)img * h * pitch;
    uint32_t* abimg = AB + (size_t)img * h * pitch * (CHROMA_F ? 2 : 1);

    const int vec_ok = VEC1 ? 1 : vec_ok_;
    if (vec_ok) {
        // 4 consecutive pixels per thread, groups aligned to 4 pixels of the image row. vec_ok == 1: the tile is a whole
        // number of groups inside the image. vec_ok == 2 (image padded by OpenCV and / or tile width not a multiple of
        // 4): the groups cover the tile's REAL pixels [xa, xb) x [ya, yb); pixels of a straddling group that belong to
        // the neighbouring tile are computed (the scratch vector both CTAs store is identical) but not counted; the
        // reflected padding is counted by the scalar loop below.
        const int xa = tx * tw, xb = min(xa + tw, w), ya = ty * th, yb = min(ya + th, h);
        const int g0 = xa >> 2, ngx = ((xb + 3) >> 2) - g0, nry = yb - ya;
        const int vq = ngx > 0 ? 256 / ngx : 0, vr = ngx > 0 ? 256 - vq * ngx : 0;
        int row = ngx > 0 ? tid / ngx : nry, c4 = ngx > 0 ? tid - row * ngx : 0;
        for (; row < nry; ) {
            const int y = ya + row, x0 = (g0 + c4) << 2;
            const size_t p = (size_t)y * w + x0, ps = (size_t)y * pitch + x0;
            int cell[4], fr[4], fg[4], fb[4];
            if (U8) {
                // 12 bytes of 4 packed RGB pixels. Rows of odd widths start at any byte: read the aligned words that
                // cover them and realign by a funnel shift. A group cut by the row end (width not a multiple of 4) is
                // read bytewise; its missing pixels are zeros (their scratch slots are padding, never counted).
                uint32_t a0, a1, a2;
                const uint8_t* src8 = in8 + p * 3;
                if (vec_ok == 1) {                               // aligned rows: the 12 bytes are three aligned words
                    const uint32_t* src = (const uint32_t*)src8;
                    a0 = __ldg(src); a1 = __ldg(src + 1); a2 = __ldg(src + 2);
                } else if (x0 + 4 <= w) {
                    const unsigned mis = (unsigned)((uintptr_t)src8 & 3);
                    const uint32_t* src = (const uint32_t*)(src8 - mis);
                    const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
                    const uint32_t w3 = mis ? __ldg(src + 3) : 0u;      // mis == 0: the 12 bytes end with w2
                    a0 = __funnelshift_r(w0, w1, mis * 8);
                    a1 = __funnelshift_r(w1, w2, mis * 8);
                    a2 = __funnelshift_r(w2, w3, mis * 8);
                } else {
                    uint32_t b[3] = {0u, 0u, 0u};
                    const int nb = (w - x0) * 3;
                    for (int k = 0; k < nb; ++k) b[k >> 2] |= (uint32_t)__ldg(src8 + k) << (8 * (k & 3));
                    a0 = b[0]; a1 = b[1]; a2 = b[2];
                }
                const int rr[4] = {(int)(a0 & 255), (int)(a0 >> 24), (int)((a1 >> 16) & 255), (int)((a2 >> 8) & 255)};
                const int gg[4] = {(int)((a0 >> 8) & 255), (int)(a1 & 255), (int)(a1 >> 24), (int)((a2 >> 16) & 255)};
                const int bb[4] = {(int)((a0 >> 16) & 255), (int)((a1 >> 8) & 255), (int)(a2 & 255), (int)(a2 >> 24)};
#pragma unroll
                for (int i = 0; i < 4; ++i) cell_from_u8(rr[i], gg[i], bb[i], cell[i], fr[i], fg[i], fb[i]);
            } else {
                const float4 r4 = __ldg((const float4*)(inf + p));
                const float4 g4 = __ldg((const float4*)(inf + plane + p));
                const float4 b4 = __ldg((const float4*)(inf + 2 * plane + p));
                const float rr[4] = {r4.x, r4.y, r4.z, r4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w},
                            bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) cell_from_f32(rr[i], gg[i], bb[i], in_norm, cell[i], fr[i], fg[i], fb[i]);
            }
            int v[4];
            uint32_t ab[4];
            if (CHROMA_A && REC32) {
                uint32_t rw[4][8];
#pragma unroll
                for (int i = 0; i < 4; ++i) ld_cell_rec32(rec32, cell[i], rw[i]);      // four sector gathers in flight
#pragma unroll
                for (int i = 0; i < 4; ++i) lab_from_rec32_px(rw[i], fr[i], fg[i], fb[i], v[i], ab[i]);
            } else if (CHROMA_A) {
                uint4 wl[4], wa[4], wb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {                                   // twelve gathers in flight
                    wl[i] = __ldg(lutL + cell[i]);
                    if (TEXAB) {
                        wa[i] = tex1Dfetch<uint4>(texAB, cell[i] * 2);
                        wb[i] = tex1Dfetch<uint4>(texAB, cell[i] * 2 + 1);
                    } else {
                        ld_cell_ab(lutAB, cell[i], wa[i], wb[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) lab_from_records(wl[i], wa[i], wb[i], fr[i], fg[i], fb[i], v[i], ab[i]);
            } else {
                uint4 wl[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)                                      // four gathers in flight
                    wl[i] = (TEXL == 1 || (TEXL == 2 && (i & 1))) ? tex1Dfetch<uint4>(texL, cell[i]) : __ldg(lutL + cell[i]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    v[i] = lab_l8_int(lab_trilinear(wl[i].x, wl[i].y, wl[i].z, wl[i].w, fr[i], fg[i], fb[i]));
                    ab[i] = pack_code(cell[i], fr[i], fg[i], fb[i]);
                }
            }
            if (CHROMA_F) {
                uint32_t t[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float ar, bz;
                    lab_chroma_terms((int)(ab[i] & 0xffffu), (int)(ab[i] >> 16), ar, bz);
                    t[2 * i] = __float_as_uint(ar);
                    t[2 * i + 1] = __float_as_uint(bz);
                }
                *(uint4*)(abimg + ps * 2) = make_uint4(t[0], t[1], t[2], t[3]);
                *(uint4*)(abimg + ps * 2 + 4) = make_uint4(t[4], t[5], t[6], t[7]);
            } else {
                *(uint4*)(abimg + ps) = make_uint4(ab[0], ab[1], ab[2], ab[3]);
            }
            *(uint32_t*)(l8img + ps) = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (VEC1) atomicAdd(&hist[v[i]], 1);             // v is a byte by construction: no "no pixel" key to test
                else hist_add(hist, (vec_ok == 1 || (x0 + i >= xa && x0 + i < xb)) ? v[i] : 256);
            }
            // next (row, group) of this thread: + 256 groups
            row += vq;
            c4 += vr;
            if (c4 >= ngx) { c4 -= ngx; ++row; }
        }
        if (vec_ok == 2) {
            // reflected padding of this tile: right strip (all tile rows) then bottom strip (columns left of the right strip)
            const int xe = xa + tw, ye = ya + th;
            const int rx0 = max(xa, w), rw = xe - rx0;                 // right strip: cols [rx0, xe), rows [ya, ye)
            const int by0 = max(ya, h), bw = min(xe, w) - xa;          // bottom strip: rows [by0, ye), cols [xa, xa + bw)
            const int nright = rw > 0 ? rw * th : 0, nbottom = (ye > by0 && bw > 0) ? (ye - by0) * bw : 0;
            for (int i = tid; i < nright + nbottom; i += 256) {
                int ey, ex;
                if (i < nright) { ey = ya + i / rw; ex = rx0 + i % rw; }
                else { const int j = i - nright; ey = by0 + j / bw; ex = xa + j % bw; }
                const size_t p = (size_t)reflect101(ey, h) * w + reflect101(ex, w);
                int cell, fr, fg, fb;
                if (U8) cell_from_u8(in8[p * 3], in8[p * 3 + 1], in8[p * 3 + 2], cell, fr, fg, fb);
                else cell_from_f32(inf[p], inf[plane + p], inf[2 * plane + p], in_norm, cell, fr, fg, fb);
                const uint4 wl = __ldg(lutL + cell);
                hist_add(hist, lab_l8_int(lab_trilinear(wl.x, wl.y, wl.z, wl.w, fr, fg, fb)));
            }
        }
    } else {
        // generic: extended (REFLECT_101-padded) tile, one pixel per thread per step
        // (row, col) walk incrementally: + 256 pixels per step, gq / gr = 256 / tw, 256 % tw
        int row = tid / tw, col = tid - row * tw;
        for (; row < th; ) {
            int v = 256;
            {
                const int ey = ty * th + row, ex = tx * tw + col;
                const int sy = reflect101(ey, h), sx = reflect101(ex, w);
                const size_t p = (size_t)sy * w + sx;
                int cell, fr, fg, fb;
                if (U8) {
                    cell_from_u8(in8[p * 3], in8[p * 3 + 1], in8[p * 3 + 2], cell, fr, fg, fb);
                } else {
                    cell_from_f32(inf[p], inf[plane + p], inf[2 * plane + p], in_norm, cell, fr, fg, fb);
                }
                uint32_t ab;
                if (CHROMA_A && REC32) {
                    uint32_t rw[8];
                    ld_cell_rec32(rec32, cell, rw);
                    lab_from_rec32_px(rw, fr, fg, fb, v, ab);
                } else if (CHROMA_A) {
                    const uint4 wl = __ldg(lutL + cell);
                    uint4 wa, wb;
                    ld_cell_ab(lutAB, cell, wa, wb);
                    lab_from_records(wl, wa, wb, fr, fg, fb, v, ab);
                } else {
                    const uint4 wl = __ldg(lutL + cell);
                    v = lab_l8_int(lab_trilinear(wl.x, wl.y, wl.z, wl.w, fr, fg, fb));
                    ab = pack_code(cell, fr, fg, fb);
                }
                if (ey < h && ex < w) {
                    l8img[(size_t)sy * pitch + sx] = (uint8_t)v;
                    abimg[(size_t)sy * pitch + sx] = ab;
                }
            }
            hist_add(hist, v);
            row += gq;
            col += gr;
            if (col >= tw) { col -= tw; ++row; }
        }
    }
    __syncthreads();

    // ---- clip, redistribute, prefix sum -> tile LUT (OpenCV clahe.cpp CLAHE_CalcLut_Body) ----
    int hv = 0;
#pragma unroll 8
    for (int c = 0; c < 8 * kHistCopies; ++c) hv += hist_all[c * kHistStride + tid];
    const int lane = tid & 31, wid = tid >> 5;
    if (clip > 0) {
        int excess = hv > clip ? hv - clip : 0;
        hv = hv > clip ? clip : hv;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) excess += __shfl_xor_sync(0xffffffffu, excess, o);
        if (lane == 0) warp_tmp[wid] = excess;
        __syncthreads();
        int clipped = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) clipped += warp_tmp[i];
        __syncthreads();
        const int batch = clipped / 256;
        int resid = clipped - batch * 256;
        hv += batch;
        if (resid != 0) {
            const int step = max(256 / resid, 1);
            if ((tid % step) == 0 && (tid / step) < resid) hv += 1;
        }
    }
    int sum = hv;  // inclusive scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, sum, o);
        if (lane >= o) sum += t;
    }
    if (lane == 31) warp_tmp[wid] = sum;
    __syncthreads();
    for (int i = 0; i < wid; ++i) sum += warp_tmp[i];
    int lv = f_rint(f_mul((float)sum, lut_scale));
    lv = lv < 0 ? 0 : (lv > 255 ? 255 : lv);
    lutT[((((size_t)img * grid + ty) * 256 + tid) << lut_row_shift(grid)) + tx] = (uint8_t)lv;
}

// ---- pass B -------------------------------------------------------------------------------------

constexpr int kSpl8Bytes = 1025 * 128, kFy8Bytes = 256 * 128;

struct NormFast {
    float mean[3], std[3], rstd[3];
    int fast;   // div_by_const_ok() for all three std
};

// FAST = the common configuration compiled without per-pixel mode tests: 8-byte LUT rows (grid <= 8) and divider-free
// normalisation. Any width: OpenCV's scalar-tail pixels (the last w % 8 of a row) are handled by a warp-uniform split --
// warps without tail pixels run the SIMD-body sequence only -- and the planar float4 stores fall back to scalar stores
// on the (row, channel) combinations that are not 16-byte aligned (widths that are not a multiple of 4).
// SPLTEX = how many of the three inverse-gamma spline lookups go through the texture pipe instead of shared memory.
// FYTEX  = lightness half of Lab->RGB ({fy, C1*y, C4*y, C7*y}, a function of the CLAHE output byte) fetched from the
// 256-entry table through the texture pipe instead of being recomputed (FAST only).
// CHROMA_A = pass A already interpolated the chroma (AB holds a | b << 16); otherwise AB holds the cell code and the
// chroma records are fetched here through the texture pipe.
// ANYW = widths with scalar-tail pixels (w % 8 != 0) or unaligned planar rows; without it (the common case, chosen by the
// host) the tail split and the store alignment test are compiled out.
// PACK = two pixels per instruction (packed f32x2 arithmetic, clahe_math.cuh) on every warp without scalar-tail pixels
// (FAST, SPLTEX == 0, !FYTEX only); bit-identical to the scalar sequence, ~45 fewer issue slots per pixel.
// PERSIST = one 1024-thread CTA per SM that walks over (image, row band, column chunk) items as four independent
// 256-thread groups (named barriers), with the inverse-gamma spline held in shared memory EIGHT times: segment i, copy c
// at [i][c], 16 bytes each, so that a 128-byte row is the same segment for the 8 lanes of a quarter warp. A 16-byte
// shared-memory load is served one quarter warp at a time; lane l reads copy l & 7, i.e. its own 16-byte bank group,
// whatever its segment index: the three random spline lookups per pixel are conflict-free (4 wavefronts per request
// instead of ~18 measured for random 16-byte reads, tools/microbench/gather_rate.cu). 128 KB of shared memory per SM,
// staged once per launch. (FAST, SPLTEX == 0, !FYTEX only.)
// CHROMA_F (PERSIST, CHROMA_A, !ANYW, !PACK; A/B variant, off by default): pass A, which has issue slots to spare at its
// gather floor, already converted the chroma to the two float terms of Lab->RGB (8 B/px scratch: ar | bz); pass B, which is
// issue-bound, saves the conversion (16 instructions per pixel). Slower in total: see g_k1_chroma_f.
template <int MINB, bool FAST, int SPLTEX, bool FYTEX, bool CHROMA_A, bool ANYW, bool PACK, bool PERSIST, bool CHROMA_F>
__global__ void __launch_bounds__(PERSIST ? 1024 : 256, PERSIST ? 1 : MINB)
clahe_apply_kernel(const uint32_t* __restrict__ AB, const uint8_t* __restrict__ L8, const uint8_t* __restrict__ lutT,
                   float* __restrict__ out, int h, int w, int pitch, int grid, float inv_th, float inv_tw, int rows_per_cta,
                   const float4* __restrict__ spline, Lab2RgbConst K,
                   NormFast on, cudaTextureObject_t texSpline, cudaTextureObject_t texFy, cudaTextureObject_t texAB,
                   int xchunks, int nbands, int nitems, int lut_area_bytes, const float4* __restrict__ fytab, int div1) {
    if (FAST) on.fast = 1;
    extern __shared__ __align__(16) uint8_t smem[];
    // !PERSIST: inverse-gamma spline segments split into two 8-byte halves (random 8-byte shared-memory gathers conflict
    // less than 16-byte ones); PERSIST: eight conflict-free copies of the 16-byte segments
    float2* spl_fb = (float2*)smem;              // [1024] (f, b)
    float2* spl_cd = (float2*)(smem + 1024 * 8); // [1024] (c, d)
    // PERSIST: [1025][8] spline segments (entry 1024 = the value at x == 1024, so the index needs no clamp), then
    // [256][8] {fy, C1*y, C4*y, C7*y} per CLAHE output byte (lightness half of Lab->RGB, build_fy_table), then LUT areas
    const float4* spl8 = (const float4*)smem;
    const float4* fy8 = (const float4*)(smem + kSpl8Bytes);
    const int lsh = FAST ? 3 : lut_row_shift(grid);
    const int tid = PERSIST ? (threadIdx.x & 255) : threadIdx.x;      // thread within the 256-thread group
    const int grp = PERSIST ? (threadIdx.x >> 8) : 0;
    // LUT rows of the group's current item: [(ty_hi - ty_lo + 1)][256] rows of (1 << lsh) bytes
    uint2* luts = (uint2*)(smem + (PERSIST ? kSpl8Bytes + kFy8Bytes + (size_t)grp * lut_area_bytes : 1024 * 16));
    if (PERSIST) {
        float4* dst = (float4*)smem;
        for (int i = threadIdx.x; i < 1025; i += 1024) {
            float4 sgm = __ldg(spline + (i < 1024 ? i : 1023));
            if (i == 1024)      // x == 1024 (linear value >= 1): segment 1023 at its right end, in spline_eval's op order
                sgm = make_float4(spline_eval(1.0f, sgm.x, sgm.y, sgm.z, sgm.w), 0.f, 0.f, 0.f);
#pragma unroll
            for (int c = 0; c < 8; ++c) dst[i * 8 + c] = sgm;
        }
        float4* dfy = (float4*)(smem + kSpl8Bytes);
        for (int i = threadIdx.x; i < 256; i += 1024) {
            const float4 t = __ldg(fytab + i);
#pragma unroll
            for (int c = 0; c < 8; ++c) dfy[i * 8 + c] = t;
        }
    } else if (SPLTEX < 3) {
        for (int i = tid; i < 1024; i += 256) {
            const float4 sgm = __ldg(spline + i);
            spl_fb[i] = make_float2(sgm.x, sgm.y);
            spl_cd[i] = make_float2(sgm.z, sgm.w);
        }
    }
    __syncthreads();
    auto group_sync = [&]() {
        if (PERSIST) asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory");
        else __syncthreads();
    };
    const int lane8 = threadIdx.x & 7;
    // byte offset of this lane's copy of spline segment ix, straight from the bits of 2^23 + ix (see `pixel`)
    const uint32_t spl_off0 = (uint32_t)lane8 * 16u - 0x80000000u;

    for (int item = PERSIST ? (int)blockIdx.x * 4 + grp : (int)blockIdx.x; item < nitems;
         item += PERSIST ? (int)gridDim.x * 4 : nitems) {
    const int bx = item % xchunks, by = (item / xchunks) % nbands, img = item / (xchunks * nbands);
    const int y0 = by * rows_per_cta;
    const int y1 = min(y0 + rows_per_cta, h);
    const int ty_lo = clahe_axis(y0, inv_th, grid).i1;
    const int ty_hi = clahe_axis(y1 - 1, inv_th, grid).i2;
    if (PERSIST) group_sync();          // the group's previous item no longer reads the LUT rows
    {
        const int nwords = ((ty_hi - ty_lo + 1) * 256) << (lsh - 3);
        const uint2* src = (const uint2*)(lutT + ((((size_t)img * grid + ty_lo) * 256) << lsh));
        for (int i = tid; i < nwords; i += 256) luts[i] = __ldg(src + i);
    }
    group_sync();

    const int x0 = (bx * 256 + tid) * 4;
    const int wbody = (w >> 3) << 3;  // pixels >= wbody take OpenCV's scalar-tail op sequence
    // does this warp own any scalar-tail pixel? (uniform per warp and constant over the rows)
    const bool tail_warp = ANYW && __any_sync(0xffffffffu, x0 < w && x0 + 4 > wbody);
    if (x0 >= w) continue;
    const int npx = min(4, w - x0);

    ClaheAxis ax[4];
    uint32_t sel[4];   // byte-permute selectors (tile columns i1, i2) of the 8-byte LUT rows
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        ax[i] = clahe_axis(x0 + i, inv_tw, grid);
        sel[i] = (uint32_t)(ax[i].i1 & 7) | ((uint32_t)(ax[i].i2 & 7) << 4);
    }

    const size_t plane = (size_t)h * w;
    // scratch rows are `pitch` = align4(w) elements apart (CHROMA_F: two words per element)
    const uint32_t* abimg = AB + (size_t)img * h * pitch * (CHROMA_F ? 2 : 1);
    const uint8_t* l8img = L8 + (size_t)img * h * pitch;
    float* outimg = out + (size_t)img * plane * 3;
    const uint8_t* lut_bytes = (const uint8_t*)luts;

    // one pixel: CLAHE blend of the lightness, chroma, Lab -> RGB, inverse gamma, normalisation
    auto pixel = [&](int i, int v, uint32_t abw, uint32_t abw2, const ClaheAxis& ay, const uint8_t* lrow1,
                     const uint8_t* lrow2, bool tail, float& o0, float& o1, float& o2) {
        // chroma: Q14 -> the a / b handed to LAB2RGB
        int oa = 0, ob = 0;
        if (CHROMA_F) {
        } else if (CHROMA_A) {
            oa = (int)(abw & 0xffffu);
            ob = (int)(abw >> 16);
        } else {
            int cell, fr, fg, fb;
            unpack_code(abw, cell, fr, fg, fb);
            const uint4 wa = tex1Dfetch<uint4>(texAB, cell * 2), wb = tex1Dfetch<uint4>(texAB, cell * 2 + 1);
            oa = lab_trilinear(wa.x, wa.y, wa.z, wa.w, fr, fg, fb);
            ob = lab_trilinear(wb.x, wb.y, wb.z, wb.w, fr, fg, fb);
        }
        const float a2 = CHROMA_F ? 0.f : lab_chroma_fast(oa), b2 = CHROMA_F ? 0.f : lab_chroma_fast(ob);
        // lightness through CLAHE: the two LUT rows hold the LUT value of every tile column at level v
        int l11, l12, l21, l22;
        if (lsh == 3) {
            // 8-byte rows: one 64-bit load per tile row, the two tile-column bytes picked by one byte permute
            const uint2 w1 = *(const uint2*)(lrow1 + (v << 3));
            const uint2 w2 = *(const uint2*)(lrow2 + (v << 3));
            const uint32_t p1 = __byte_perm(w1.x, w1.y, sel[i]), p2 = __byte_perm(w2.x, w2.y, sel[i]);
            l11 = p1 & 255; l12 = (p1 >> 8) & 255;
            l21 = p2 & 255; l22 = (p2 >> 8) & 255;
        } else {
            const uint8_t* r1 = lrow1 + (v << lsh);
            const uint8_t* r2 = lrow2 + (v << lsh);
            l11 = r1[ax[i].i1]; l12 = r1[ax[i].i2];
            l21 = r2[ax[i].i1]; l22 = r2[ax[i].i2];
        }
        const int dst = clahe_blend(l11, l12, l21, l22, ax[i].a, ax[i].a1, ay.a, ay.a1);
        float lr, lg, lb;
        if (PERSIST && !tail) {
            const float4 fy = fy8[(dst << 3) | lane8];          // this lane's own copy: conflict-free
            if (CHROMA_F) lab2lin_body_from_fy_pre(fy.x, fy.y, fy.z, fy.w, __uint_as_float(abw), __uint_as_float(abw2), K, lr, lg, lb);
            else lab2lin_body_from_fy(fy.x, fy.y, fy.z, fy.w, a2, b2, K, lr, lg, lb);
        } else if (FAST && FYTEX && !tail) {
            const float4 fy = tex1Dfetch<float4>(texFy, dst);
            lab2lin_body_from_fy(fy.x, fy.y, fy.z, fy.w, a2, b2, K, lr, lg, lb);
        } else {
            lab2lin(lab_l_from_u8_fast(dst), a2, b2, tail, K, lr, lg, lb);
        }
        float e[3];
        if (PERSIST) {
            // x = clamp01(lin) * 1024 in [0, 1024]; 2^23 + x rounded toward zero is 2^23 + trunc(x): its low mantissa bits
            // are the segment index (no float <-> int conversion, no clamp: entry 1024 exists), minus 2^23 it is float(ix)
            const float lin3[3] = {lr, lg, lb};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float x = f_mul(__saturatef(lin3[c]), 1024.0f);
                const float t = __fadd_rz(x, 8388608.0f);
                const float xs = f_sub(x, f_sub(t, 8388608.0f));
                const uint32_t off = (__float_as_uint(t) << 7) + spl_off0;
                const float4 sg = *(const float4*)(smem + off);
                e[c] = spline_eval(xs, sg.x, sg.y, sg.z, sg.w);
            }
        }
        int ix[3] = {0, 0, 0};
        float xs[3] = {0.f, 0.f, 0.f};
        if (!PERSIST) {
            xs[0] = spline_index(lr, ix[0]); xs[1] = spline_index(lg, ix[1]); xs[2] = spline_index(lb, ix[2]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (PERSIST) {
            } else if (c < SPLTEX) {       // texture pipe: no shared-memory bank conflicts
                const float4 sg = tex1Dfetch<float4>(texSpline, ix[c]);
                e[c] = spline_eval(xs[c], sg.x, sg.y, sg.z, sg.w);
            } else {
                const float2 s01 = spl_fb[ix[c]], s23 = spl_cd[ix[c]];
                e[c] = spline_eval(xs[c], s01.x, s01.y, s23.x, s23.y);
            }
        }
        if (PERSIST && div1) {      // std values for which ONE correction step was verified exhaustively (gdt_init)
            o0 = div_by_const<1>(f_sub(e[0], on.mean[0]), on.std[0], on.rstd[0]);
            o1 = div_by_const<1>(f_sub(e[1], on.mean[1]), on.std[1], on.rstd[1]);
            o2 = div_by_const<1>(f_sub(e[2], on.mean[2]), on.std[2], on.rstd[2]);
            return;
        }
        o0 = on.fast ? normalize_px_fast(e[0], on.mean[0], on.std[0], on.rstd[0]) : normalize_px(e[0], on.mean[0], on.std[0]);
        o1 = on.fast ? normalize_px_fast(e[1], on.mean[1], on.std[1], on.rstd[1]) : normalize_px(e[1], on.mean[1], on.std[1]);
        o2 = on.fast ? normalize_px_fast(e[2], on.mean[2], on.std[2], on.rstd[2]) : normalize_px(e[2], on.mean[2], on.std[2]);
    };

    // two pixels (i, i + 1) per instruction: the same operation sequence as `pixel`, packed lane-wise
    f2 xa2[2], xa12[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        xa2[j] = mk2(ax[2 * j].a, ax[2 * j + 1].a);
        xa12[j] = mk2(ax[2 * j].a1, ax[2 * j + 1].a1);
    }
    auto pixel_pair = [&](int j, const int* v, const uint32_t* abw, const ClaheAxis& ay, const uint8_t* lrow1,
                          const uint8_t* lrow2, f2& o0, f2& o1, f2& o2) {
        int oa[2], ob[2], l11[2], l12[2], l21[2], l22[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            if (CHROMA_A) {
                oa[t] = (int)(abw[t] & 0xffffu);
                ob[t] = (int)(abw[t] >> 16);
            } else {
                int cell, fr, fg, fb;
                unpack_code(abw[t], cell, fr, fg, fb);
                const uint4 wa = tex1Dfetch<uint4>(texAB, cell * 2), wb = tex1Dfetch<uint4>(texAB, cell * 2 + 1);
                oa[t] = lab_trilinear(wa.x, wa.y, wa.z, wa.w, fr, fg, fb);
                ob[t] = lab_trilinear(wb.x, wb.y, wb.z, wb.w, fr, fg, fb);
            }
            const uint2 w1 = *(const uint2*)(lrow1 + (v[t] << 3));
            const uint2 w2 = *(const uint2*)(lrow2 + (v[t] << 3));
            const uint32_t p1 = __byte_perm(w1.x, w1.y, sel[2 * j + t]), p2 = __byte_perm(w2.x, w2.y, sel[2 * j + t]);
            l11[t] = p1 & 255; l12[t] = (p1 >> 8) & 255;
            l21[t] = p2 & 255; l22[t] = (p2 >> 8) & 255;
        }
        const f2 a2 = lab_chroma_fast2(oa[0], oa[1]), b2 = lab_chroma_fast2(ob[0], ob[1]);
        int d0, d1;
        clahe_blend2(l11, l12, l21, l22, xa2[j], xa12[j], ay.a, ay.a1, d0, d1);
        f2 y, fy, lin[3];
        lab_fy_body2(lab_l_from_u8_fast2(d0, d1), y, fy);
        lab2lin_body2(fy, y, a2, b2, K, lin[0], lin[1], lin[2]);
        f2 e[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int i0, i1;
            const f2 xs = spline_index2(lin[c], i0, i1);
            if (PERSIST) {
                const float4 s0 = spl8[(i0 << 3) | lane8], s1 = spl8[(i1 << 3) | lane8];
                e[c] = spline_eval2(xs, mk2(s0.x, s1.x), mk2(s0.y, s1.y), mk2(s0.z, s1.z), mk2(s0.w, s1.w));
            } else {
                const float2 f0 = spl_fb[i0], c0 = spl_cd[i0], f1 = spl_fb[i1], c1 = spl_cd[i1];
                e[c] = spline_eval2(xs, mk2(f0.x, f1.x), mk2(f0.y, f1.y), mk2(c0.x, c1.x), mk2(c0.y, c1.y));
            }
        }
        o0 = normalize_px_fast2(e[0], on.mean[0], on.std[0], on.rstd[0]);
        o1 = normalize_px_fast2(e[1], on.mean[1], on.std[1], on.rstd[1]);
        o2 = normalize_px_fast2(e[2], on.mean[2], on.std[2], on.rstd[2]);
    };

    // software prefetch: the next row's scratch words are requested before this row's arithmetic
    uint4 nxc, nxc2 = make_uint4(0u, 0u, 0u, 0u);
    uint32_t nxl;
    {
        const size_t ps = (size_t)y0 * pitch + x0;
        if (CHROMA_F) {
            nxc = __ldg((const uint4*)(abimg + ps * 2));
            nxc2 = __ldg((const uint4*)(abimg + ps * 2 + 4));
        } else {
            nxc = __ldg((const uint4*)(abimg + ps));
        }
        nxl = __ldg((const uint32_t*)(l8img + ps));
    }
    for (int y = y0; y < y1; ++y) {
        const ClaheAxis ay = clahe_axis(y, inv_th, grid);
        const uint8_t* lrow1 = lut_bytes + (((size_t)(ay.i1 - ty_lo) * 256) << lsh);
        const uint8_t* lrow2 = lut_bytes + (((size_t)(ay.i2 - ty_lo) * 256) << lsh);
        const size_t p = (size_t)y * w + x0, ps = (size_t)y * pitch + x0;

        uint32_t lw = nxl;
        uint4 cw = nxc, cw2 = nxc2;
        if (y + 1 < y1) {
            if (CHROMA_F) {
                nxc = __ldg((const uint4*)(abimg + (ps + pitch) * 2));
                nxc2 = __ldg((const uint4*)(abimg + (ps + pitch) * 2 + 4));
            } else {
                nxc = __ldg((const uint4*)(abimg + ps + pitch));
            }
            nxl = __ldg((const uint32_t*)(l8img + ps + pitch));
        }
        if (ANYW && npx < 4) {  // slots past the row end are padding (possibly never written): neutral values
            if (npx < 2) { cw.y = 0u; lw &= 0xffu; }
            if (npx < 3) { cw.z = 0u; lw &= 0xffffu; }
            cw.w = 0u; lw &= 0xffffffu;
        }
        const int v[4] = {(int)(lw & 255), (int)((lw >> 8) & 255), (int)((lw >> 16) & 255), (int)(lw >> 24)};
        // CHROMA_F: pixel i's terms are (ar, bz) = words (2 i, 2 i + 1) of the 8-word group
        const uint32_t ab[4] = {cw.x, CHROMA_F ? cw.z : cw.y, CHROMA_F ? cw2.x : cw.z, CHROMA_F ? cw2.z : cw.w};
        const uint32_t ab2[4] = {cw.y, cw.w, cw2.y, cw2.w};

        float o[3][4];
        if (PACK && !tail_warp) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                f2 p0, p1, p2;
                pixel_pair(j, v + 2 * j, ab + 2 * j, ay, lrow1, lrow2, p0, p1, p2);
                o[0][2 * j] = p0.x; o[0][2 * j + 1] = p0.y;
                o[1][2 * j] = p1.x; o[1][2 * j + 1] = p1.y;
                o[2][2 * j] = p2.x; o[2][2 * j + 1] = p2.y;
            }
        } else if (tail_warp) {
#pragma unroll
            for (int i = 0; i < 4; ++i) pixel(i, v[i], ab[i], ab2[i], ay, lrow1, lrow2, (x0 + i) >= wbody, o[0][i], o[1][i], o[2][i]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) pixel(i, v[i], ab[i], ab2[i], ay, lrow1, lrow2, false, o[0][i], o[1][i], o[2][i]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float* dst = outimg + c * plane + p;
            if (!ANYW || (npx == 4 && (((uintptr_t)dst) & 15) == 0)) {
                __stcs((float4*)dst, make_float4(o[c][0], o[c][1], o[c][2], o[c][3]));
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < npx) __stcs(dst + i, o[c][i]);
            }
        }
    }
    }   // items
}

// ---- host ---------------------------------------------------------------------------------------

struct ClaheGeom {
    int eh, ew, th, tw, clip, rows_per_cta;
    float lut_scale, inv_th, inv_tw;
};

static int clahe_geometry(int n, int h, int w, double clip_limit, int grid, ClaheGeom& g) {
    if (n <= 0 || h <= 0 || w <= 0 || grid < 1 || grid > 16) return GDT_ERR_INVALID_ARGUMENT;
    if (h % grid == 0 && w % grid == 0) {
        g.eh = h; g.ew = w;
    } else {  // OpenCV quirk: a dimension that IS divisible still gets +grid
        g.eh = h + (grid - h % grid);
        g.ew = w + (grid - w % grid);
        if (g.eh - h >= h || g.ew - w >= w) return GDT_ERR_UNSUPPORTED;  // REFLECT_101 needs pad < size
    }
    g.th = g.eh / grid; g.tw = g.ew / grid;
    const int area = g.th * g.tw;
    g.lut_scale = 255.0f / (float)area;
    g.clip = 0;
    if (clip_limit > 0.0) {
        g.clip = (int)(clip_limit * area / 256);
        if (g.clip < 1) g.clip = 1;
    }
    g.inv_th = 1.0f / (float)g.th;
    g.inv_tw = 1.0f / (float)g.tw;
    return GDT_OK;
}

// rows per pass-B CTA: the largest candidate whose grid wastes < 3 % of its last wave, else the most efficient one
static int pass_b_rows(int h, long long ctas_per_band, long long slots) {
    static const int cand[] = {48, 40, 32, 28, 24, 20, 16, 12, 8, 6, 4, 2};
    int best = 2;
    double best_eff = -1.0;
    for (int r : cand) {
        const long long ctas = (long long)ceil_div(h, r) * ctas_per_band;
        const long long waves = ceil_div_ll(ctas, slots);
        // rows that do not divide h leave a short last band: count it as full work (conservative)
        const double eff = (double)ctas / (double)(waves * slots);
        if (ctas >= slots && eff >= 0.97) return r;
        if (eff > best_eff + 1e-9) { best_eff = eff; best = r; }
    }
    return best;
}

template <bool U8>
static int clahe_launch_chunk(const void* in, int n, int h, int w, double clip_limit, int grid, const Norm3& in_norm,
                              const Norm3& out_norm, float* out, void* ws, size_t ws_bytes, cudaStream_t stream) {
    const ClaheTables* T = clahe_tables_for_current_device();
    if (!T) return GDT_ERR_NOT_INITIALISED;
    if (!in || !out || !ws) return GDT_ERR_INVALID_ARGUMENT;
    ClaheGeom g;
    int rc = clahe_geometry(n, h, w, clip_limit, grid, g);
    if (rc != GDT_OK) return rc;
    if (ws_bytes < gdt_clahe_workspace_bytes(n, h, w, grid)) return GDT_ERR_WORKSPACE_TOO_SMALL;
    Workspace W(ws, ws_bytes);
    const int pitch = (w + 3) & ~3;          // scratch row pitch in elements: every 4-pixel group is 16-byte aligned
    uint8_t* L8 = W.take<uint8_t>((size_t)n * h * pitch);
    // 4 B/px (Q14 pair or cell code); 8 B/px only while the float-terms variant is switched on
    uint32_t* AB = W.take<uint32_t>((size_t)n * h * pitch * (g_k1_chroma_f ? 2 : 1));
    uint8_t* luts = W.take<uint8_t>(((size_t)n * grid * 256) << lut_row_shift(grid));
    if (!W.ok()) return GDT_ERR_WORKSPACE_TOO_SMALL;

    // Pass A, 4-pixel groups: 1 = every tile is a whole number of aligned groups inside the image; 2 = aligned groups
    // over ragged / padded tiles (uint8 input: any width, rows are realigned by funnel shifts; float input: planar
    // float4 loads need a width that is a multiple of 4 and an aligned base); 0 = one pixel per thread.
    const bool in_aligned = (((uintptr_t)in) & 15) == 0;
    int vec_hist;
    if (in_aligned && (w % 4) == 0 && g.eh == h && g.ew == w && (g.tw % 4) == 0) vec_hist = 1;
    else if (U8 ? (((uintptr_t)in) & 3) == 0 : (in_aligned && (w % 4) == 0)) vec_hist = 2;
    else vec_hist = 0;

    // A/B switches, see gdt_debug_k1_config (profiles/k1_v2_ab_r1q.log)
    const int texab = g_k1_texab, spltex = g_k1_spltex, fytex = g_k1_fytex, occ_a = g_k1_occ_a;
    const bool rec32 = g_k1_rec32 != 0 && T->rec_ok;
    const int chroma_a = g_k1_chroma_a < 0 ? (rec32 ? 1 : 0) : g_k1_chroma_a;
    // Rows per CTA: as many as possible (amortises the LUT / spline staging) while the grid still fills whole waves of
    // the machine: 4 resident CTAs per SM, and a last wave that is mostly empty costs up to a wave of time (32 rows on
    // 128 images of 768 rows = 5.2 waves; 24 rows = 6.9).
    const int sms = sm_count_current_device();
    const int xchunks = ceil_div(w, 1024);
    int rows = pass_b_rows(h, (long long)xchunks * n, 4LL * sms);
    if (g_k1_rows > 0) rows = g_k1_rows;
    const int nbands = ceil_div(h, rows);
    const long long nitems_ll = (long long)xchunks * nbands * n;
    if (nitems_ll > 0x7fffffffLL) return GDT_ERR_UNSUPPORTED;
    const int nitems = (int)nitems_ll;
    // spline table + the LUT rows of every tile row a band of `rows` image rows can touch
    int span = (rows + g.th - 1) / g.th + 2;
    if (span > grid) span = grid;
    const int lut_area = (span * 256) << lut_row_shift(grid);
    const size_t smem = 1024 * 16 + (size_t)lut_area;
    NormFast on;
    on.fast = 1;
    for (int c = 0; c < 3; ++c) {
        on.mean[c] = out_norm.mean[c];
        on.std[c] = out_norm.std[c];
        volatile float r = 1.0f / out_norm.std[c];
        on.rstd[c] = r;
        if (!div_by_const_ok(out_norm.std[c])) on.fast = 0;
    }
    // widths with OpenCV scalar-tail pixels or rows of the planar output that are not 16-byte aligned
    const bool anyw = (w & 7) != 0 || (((uintptr_t)out) & 15) != 0;
    if (smem > 48 * 1024) {
        GDT_CUDA(cudaFuncSetAttribute(clahe_apply_kernel<4, false, 0, false, true, true, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      1024 * 16 + 16 * 256 * 16));
        GDT_CUDA(cudaFuncSetAttribute(clahe_apply_kernel<4, false, 0, false, false, true, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      1024 * 16 + 16 * 256 * 16));
    }
    // non-persistent form: one 256-thread CTA per item, 4 resident CTAs per SM (64 registers; 6 and 8 CTAs per SM at
    // 40 / 32 registers measured no faster)
#define GDT_APPLY_P(FAST_, S_, F_, C_, A_, P_)                                                                            \
    clahe_apply_kernel<4, FAST_, S_, F_, C_, A_, P_, false, false><<<nitems, 256, smem, stream>>>(                         \
        AB, L8, luts, out, h, w, pitch, grid, g.inv_th, g.inv_tw, rows, T->spline, T->K, on, T->texSpline, T->texFy,       \
        T->texAB, xchunks, nbands, nitems, lut_area, T->fytab, 0)
#define GDT_APPLY(FAST_, S_, F_, C_, A_) GDT_APPLY_P(FAST_, S_, F_, C_, A_, false)
    // persistent form (the default for the common configuration): one 1024-thread CTA per SM = four 256-thread groups,
    // eight conflict-free copies of the spline in shared memory (128 KB) + one LUT area per group
    const size_t smem_p = kSpl8Bytes + kFy8Bytes + 4 * (size_t)lut_area;
    int div1 = g_k1_div1 != 0;
    for (int c = 0; c < 3; ++c) div1 = div1 && T->div1_verified(out_norm.std[c]);
#define GDT_APPLY_PERSIST(C_, A_, P_) GDT_APPLY_PERSIST_F(C_, A_, P_, false)
#define GDT_APPLY_PERSIST_F(C_, A_, P_, CF_)                                                                             \
    do {                                                                                                                 \
        auto kern = clahe_apply_kernel<4, true, 0, false, C_, A_, P_, true, CF_>;                                        \
        static bool attr_done[32] = {false};                                                                             \
        const int slot = current_device_slot();                                                                          \
        if (!attr_done[slot]) {                                                                                          \
            GDT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,                             \
                                          kSpl8Bytes + kFy8Bytes + 4 * 8 * 2048));                                       \
            attr_done[slot] = true;                                                                                      \
        }                                                                                                                \
        const int ctas = nitems < 4 * sms ? ceil_div(nitems, 4) : sms;                                                   \
        kern<<<ctas, 1024, smem_p, stream>>>(AB, L8, luts, out, h, w, pitch, grid, g.inv_th, g.inv_tw, rows, T->spline,  \
                                             T->K, on, T->texSpline, T->texFy, T->texAB, xchunks, nbands, nitems,        \
                                             lut_area, T->fytab, div1);                                                  \
    } while (0)
    const bool pack = g_k1_pack != 0;
    // measured (profiles/k1_v3_ab_r2n.log): persistent wins on the common widths (0.950 vs 1.008 ms per 128 images of
    // 1024x768), the any-width instantiation (scalar-tail split, more live state) is faster non-persistent; 2 = force
    const bool persist = (g_k1_persist == 2 || (g_k1_persist == 1 && !anyw)) && spltex == 0 && !fytex;
    const bool fast_b = grid <= 8 && on.fast && smem <= 48 * 1024;
    // float chroma scratch: only between the specialised pass A and the persistent scalar pass B, on widths without
    // OpenCV scalar-tail pixels (those use true divisions for the same terms)
    const bool chroma_f = g_k1_chroma_f != 0 && rec32 && chroma_a && vec_hist == 1 && fast_b && persist && !anyw && !pack &&
                          (w & 7) == 0 && occ_a < 6;
    dim3 gridA(grid * grid, n);
    const int gw = g.tw;                               // scalar path: walk unit = one pixel of a tile row
    const int gq = 256 / gw, gr = 256 % gw;
#define GDT_HIST_RVF(T_, C_, O_, L_, R_, V_, F_)                                                                       \
    clahe_hist_kernel<U8, T_, C_, O_, L_, R_, V_, F_><<<gridA, 256, 0, stream>>>(in, L8, AB, luts, h, w, pitch, grid, g.th, g.tw, \
                                                                          g.clip, g.lut_scale, vec_hist, gq, gr, T->lutL, \
                                                                          T->lutAB, in_norm, T->texAB, T->texL, T->rec32)
#define GDT_HIST_RV(T_, C_, O_, L_, R_, V_) GDT_HIST_RVF(T_, C_, O_, L_, R_, V_, false)
#define GDT_HIST_R(T_, C_, O_, L_, R_) GDT_HIST_RV(T_, C_, O_, L_, R_, false)
#define GDT_HIST(T_, C_, O_, L_) GDT_HIST_R(T_, C_, O_, L_, false)
    if (!chroma_a) {
        if (texab & 4) GDT_HIST(false, false, 4, 2);
        else if (texab & 2) GDT_HIST(false, false, 4, 1);
        else if (occ_a >= 6) GDT_HIST(false, false, 6, 0);
        else GDT_HIST(false, false, 4, 0);
    } else if (rec32) {
        if (occ_a >= 6) GDT_HIST_R(false, true, 6, 0, true);
        else if (chroma_f) GDT_HIST_RVF(false, true, 4, 0, true, true, true);         // + float chroma terms for pass B
        else if (vec_hist == 1) GDT_HIST_RV(false, true, 4, 0, true, true);           // the common sizes: specialised
        else GDT_HIST_R(false, true, 4, 0, true);
    } else if (texab & 1) {
        if (occ_a >= 6) GDT_HIST(true, true, 6, 0); else GDT_HIST(true, true, 4, 0);
    } else {
        GDT_HIST(false, true, 4, 0);
    }
#undef GDT_HIST_R
#undef GDT_HIST_RV
#undef GDT_HIST_RVF
#undef GDT_HIST
    GDT_LAUNCH_CHECK();

    if (fast_b) {
        if (persist && chroma_f) {
            GDT_APPLY_PERSIST_F(true, false, false, true);
        } else if (persist) {
            switch ((chroma_a ? 4 : 0) + (anyw ? 2 : 0) + (pack ? 1 : 0)) {
                case 0: GDT_APPLY_PERSIST(false, false, false); break;
                case 1: GDT_APPLY_PERSIST(false, false, true); break;
                case 2: GDT_APPLY_PERSIST(false, true, false); break;
                case 3: GDT_APPLY_PERSIST(false, true, true); break;
                case 4: GDT_APPLY_PERSIST(true, false, false); break;
                case 5: GDT_APPLY_PERSIST(true, false, true); break;
                case 6: GDT_APPLY_PERSIST(true, true, false); break;
                default: GDT_APPLY_PERSIST(true, true, true); break;
            }
        } else if (anyw) {     // the pipe variants are A/B material for the common case only
            if (chroma_a) { if (pack) GDT_APPLY_P(true, 0, false, true, true, true); else GDT_APPLY(true, 0, false, true, true); }
            else { if (pack) GDT_APPLY_P(true, 0, false, false, true, true); else GDT_APPLY(true, 0, false, false, true); }
        } else if (pack && spltex == 0 && !fytex) {
            if (chroma_a) GDT_APPLY_P(true, 0, false, true, false, true); else GDT_APPLY_P(true, 0, false, false, false, true);
        } else switch ((spltex > 1 ? 1 : spltex) * 4 + fytex * 2 + chroma_a) {
            case 0: GDT_APPLY(true, 0, false, false, false); break;
            case 1: GDT_APPLY(true, 0, false, true, false); break;
            case 2: GDT_APPLY(true, 0, true, false, false); break;
            case 3: GDT_APPLY(true, 0, true, true, false); break;
            case 4: GDT_APPLY(true, 1, false, false, false); break;
            case 5: GDT_APPLY(true, 1, false, true, false); break;
            case 6: GDT_APPLY(true, 1, true, false, false); break;
            default: GDT_APPLY(true, 1, true, true, false); break;
        }
    } else if (chroma_a) {
        GDT_APPLY(false, 0, false, true, true);
    } else {
        GDT_APPLY(false, 0, false, false, true);
    }
#undef GDT_APPLY
#undef GDT_APPLY_PERSIST
#undef GDT_APPLY_PERSIST_F
#undef GDT_APPLY_P
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}


// Images per launch pair. Pass A writes 5 B/px of scratch that pass B reads back: when the whole batch goes through pass
// A first, the scratch of a large batch (128 images of 1024x768: 503 MB) has left the 126 MB L2 long before pass B wants
// it (measured DRAM traffic 1.38x the algorithmic bytes). Running the two passes chunk by chunk (-1: a chunk sized to
// about half of L2, rounded to whole waves of pass A) keeps a chunk's scratch L2-resident -- and is SLOWER on B200
// (1.10 - 1.25 ms against 1.03 ms per 128 images, profiles/k1_chunk_ab_r2a.log): the kernels are bound by the SM's
// issue / LSU / texture pipes, not by DRAM (22 % busy), and every extra launch pair adds a partially filled last wave.
// The whole batch at once stays the default; the hook remains for A/B runs.
static int clahe_chunk_images(int n, int h, int w, int grid) {
    if (g_k1_chunk == 0) return n;
    if (g_k1_chunk > 0) return g_k1_chunk < n ? g_k1_chunk : n;
    const size_t pitch = ((size_t)w + 3) & ~(size_t)3;
    const size_t scratch_per_image = (size_t)h * pitch * 5;
    int dev = 0, l2 = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev) != cudaSuccess || l2 <= 0)
        l2 = 64 << 20;
    long long c = (long long)((size_t)l2 * 9 / 16 / (scratch_per_image ? scratch_per_image : 1));
    if (c < 1) c = 1;
    if (c >= n) return n;
    // whole waves of pass A: 4 resident CTAs per SM
    const long long slots = 4LL * sm_count_current_device(), per_img = (long long)grid * grid;
    const long long waves = (c * per_img) / slots;
    if (waves >= 1) {
        const long long fit = (waves * slots) / per_img;
        if (fit >= 1) c = fit;
    }
    // balance: the same number of chunks, equal sizes
    const long long chunks = ceil_div_ll(n, c);
    return (int)ceil_div_ll(n, chunks);
}

template <bool U8>
static int clahe_launch(const void* in, int n, int h, int w, double clip_limit, int grid, const Norm3& in_norm,
                        const Norm3& out_norm, float* out, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (n <= 0 || h <= 0 || w <= 0 || grid < 1 || grid > 16) return GDT_ERR_INVALID_ARGUMENT;
    if (ws_bytes < gdt_clahe_workspace_bytes(n, h, w, grid)) return GDT_ERR_WORKSPACE_TOO_SMALL;
    const int chunk = clahe_chunk_images(n, h, w, grid);
    const size_t in_stride = (size_t)h * w * 3 * (U8 ? 1 : 4), out_stride = (size_t)h * w * 3;
    for (int i0 = 0; i0 < n; i0 += chunk) {
        const int m = n - i0 < chunk ? n - i0 : chunk;
        // every chunk reuses the head of the workspace: the launches are stream-ordered
        const int rc = clahe_launch_chunk<U8>((const char*)in + (size_t)i0 * in_stride, m, h, w, clip_limit, grid, in_norm,
                                              out_norm, out + (size_t)i0 * out_stride, ws, ws_bytes, stream);
        if (rc != GDT_OK) return rc;
    }
    return GDT_OK;
}

// MeanStdPost / MeanStdPre (wrapper.py:149-194): y = ((x * s0 + m0) - m1) / s1 per channel, four separate roundings.
// One thread per 4 consecutive floats of a plane (planes that are a multiple of 4 long and 16-byte aligned), else scalar.
__global__ void __launch_bounds__(256)
meanstd_adapt_kernel(const float* __restrict__ x, float* __restrict__ y, long long planes, long long plane, Norm3 in_n,
                     NormFast out_n, int vec) {
    const long long per = vec ? plane >> 2 : plane;
    const long long total = planes * per;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long pl = i / per;
        const int c = (int)(pl % 3);
        const float s0 = in_n.std[c], m0 = in_n.mean[c], m1 = out_n.mean[c], s1 = out_n.std[c], r1 = out_n.rstd[c];
        auto one = [&](float v) {
            const float t = f_sub(f_add(f_mul(v, s0), m0), m1);
            return out_n.fast ? div_by_const<2>(t, s1, r1) : f_div(t, s1);
        };
        if (vec) {
            const float4 v = __ldcs((const float4*)x + i);
            __stcs((float4*)y + i, make_float4(one(v.x), one(v.y), one(v.z), one(v.w)));
        } else {
            y[i] = one(x[i]);
        }
    }
}

// debug: count floats a in the bit range [lo_bits, hi_bits] (both signs) for which div_by_const<ITERS> != a / b
template <int ITERS>
__global__ void __launch_bounds__(256)
div_check_kernel(float b, float r, uint32_t lo_bits, uint32_t hi_bits, unsigned long long* __restrict__ mismatches) {
    unsigned long long bad = 0;
    const unsigned long long n = (unsigned long long)hi_bits - lo_bits + 1;
    for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * 256) {
        const float a = __uint_as_float(lo_bits + (uint32_t)i);
        if (__float_as_uint(div_by_const<ITERS>(a, b, r)) != __float_as_uint(f_div(a, b))) ++bad;
        if (__float_as_uint(div_by_const<ITERS>(-a, b, r)) != __float_as_uint(f_div(-a, b))) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// Is ONE correction step enough for this divisor? Every numerator K1's normalisation can see is (spline output in
// [0, 1]) - mean, i.e. 0 or 2^-30 <= |a| <= 2: all ~7.5e8 such floats of both signs are tried on the device (~1 ms).
static bool verify_div1(float b) {
    unsigned long long* dcount = nullptr;
    if (cudaMalloc(&dcount, sizeof(unsigned long long)) != cudaSuccess) { cudaGetLastError(); return false; }
    bool ok = false;
    volatile float r = 1.0f / b;
    const float lo = 9.313225746154785e-10f /* 2^-30 */, hi = 2.0f;
    uint32_t lo_bits, hi_bits;
    memcpy(&lo_bits, &lo, 4);
    memcpy(&hi_bits, &hi, 4);
    unsigned long long h = 1;
    if (cudaMemset(dcount, 0, sizeof(unsigned long long)) == cudaSuccess) {
        div_check_kernel<1><<<sm_count_current_device() * 8, 256>>>(b, r, lo_bits, hi_bits, dcount);
        if (cudaGetLastError() == cudaSuccess &&
            cudaMemcpy(&h, dcount, sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess)
            ok = (h == 0);
    }
    cudaGetLastError();
    cudaFree(dcount);
    return ok;
}

}  // namespace gdt

using namespace gdt;

extern "C" int gdt_debug_div_check(float b, uint32_t lo_bits, uint32_t hi_bits, unsigned long long* mismatches_dev,
                                   void* stream) {
    if (!mismatches_dev || hi_bits < lo_bits) return GDT_ERR_INVALID_ARGUMENT;
    volatile float r = 1.0f / b;
    div_check_kernel<2><<<148 * 8, 256, 0, (cudaStream_t)stream>>>(b, r, lo_bits, hi_bits, mismatches_dev);
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}

extern "C" int gdt_meanstd_adapt(const float* x, long long n, long long plane, const float* host_in_mean,
                                 const float* host_in_std, const float* host_out_mean, const float* host_out_std, float* y,
                                 void* stream) {
    if (!x || !y || !host_in_mean || !host_in_std || !host_out_mean || !host_out_std || n < 0 || plane < 0)
        return GDT_ERR_INVALID_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return GDT_ERR_NO_DEVICE; }
    if (n == 0 || plane == 0) return GDT_OK;
    Norm3 i;
    NormFast o;
    o.fast = 0;     // arbitrary input range: always the true division (the kernel is HBM-bound either way)
    for (int c = 0; c < 3; ++c) {
        if (host_out_std[c] == 0.f) return GDT_ERR_INVALID_ARGUMENT;
        i.mean[c] = host_in_mean[c]; i.std[c] = host_in_std[c];
        o.mean[c] = host_out_mean[c]; o.std[c] = host_out_std[c];
        o.rstd[c] = 0.f;
    }
    const int vec = (plane % 4 == 0) && ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0;
    const long long work = n * 3 * (vec ? plane / 4 : plane);
    long long blocks = ceil_div_ll(work, 256);
    const long long cap = (long long)sm_count_current_device() * 16;
    if (blocks > cap) blocks = cap;
    meanstd_adapt_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, n * 3, plane, i, o, vec);
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}

extern "C" int gdt_debug_k1_config(int texab, int spltex, int fytex, int chroma_a, int occ_a) {
    if (spltex < 0 || spltex > 1 || (occ_a != 4 && occ_a != 6)) return GDT_ERR_INVALID_ARGUMENT;
    g_k1_texab = texab & 7;
    g_k1_spltex = spltex;
    g_k1_fytex = fytex ? 1 : 0;
    g_k1_chroma_a = chroma_a < 0 ? -1 : (chroma_a ? 1 : 0);
    g_k1_occ_a = occ_a;
    return GDT_OK;
}

extern "C" int gdt_debug_k1_chunk(int images_per_launch_pair) {
    if (images_per_launch_pair < -1) return GDT_ERR_INVALID_ARGUMENT;
    g_k1_chunk = images_per_launch_pair;
    return GDT_OK;
}

extern "C" int gdt_debug_k1_pack(int packed_f32x2) {
    g_k1_pack = packed_f32x2 ? 1 : 0;
    return GDT_OK;
}

extern "C" int gdt_debug_k1_rec32(int compressed_record) {
    g_k1_rec32 = compressed_record ? 1 : 0;
    return GDT_OK;
}

extern "C" int gdt_debug_k1_chroma_f(int float_terms) {
    g_k1_chroma_f = float_terms ? 1 : 0;
    return GDT_OK;
}

extern "C" int gdt_debug_k1_div1(int one_step) {
    g_k1_div1 = one_step ? 1 : 0;
    return GDT_OK;
}

extern "C" int gdt_debug_k1_div1_verified(float std) {
    const ClaheTables* T = clahe_tables_for_current_device();
    if (!T) return GDT_ERR_NOT_INITIALISED;
    return T->div1_verified(std) ? 1 : 0;
}

extern "C" int gdt_debug_k1_persist(int persistent) {
    g_k1_persist = persistent < 0 ? 0 : (persistent > 2 ? 2 : persistent);
    return GDT_OK;
}

extern "C" int gdt_debug_k1_rows(int rows_per_cta) {
    if (rows_per_cta < 0 || rows_per_cta > 64) return GDT_ERR_INVALID_ARGUMENT;
    g_k1_rows = rows_per_cta;
    return GDT_OK;
}

extern "C" int gdt_init(const int16_t* host_rgb2lab_lut) {
    if (!host_rgb2lab_lut) return GDT_ERR_INVALID_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return GDT_ERR_NO_DEVICE;
    }
    int dev = 0;
    GDT_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 32) return GDT_ERR_UNSUPPORTED;
    ClaheTables& T = g_tables[dev];
    if (T.ready) return GDT_OK;

    const size_t ncell = kLabCells;
    uint32_t* hL = (uint32_t*)malloc(ncell * 4 * sizeof(uint32_t));
    uint32_t* hAB = (uint32_t*)malloc(ncell * 8 * sizeof(uint32_t));
    if (!hL || !hAB) { free(hL); free(hAB); return GDT_ERR_INVALID_ARGUMENT; }
    pack_lab_lut(host_rgb2lab_lut, hL, hAB);
    build_inv_gamma_spline(T.spline_host);
    build_lab2rgb_const(T.K);
    build_fy_table(T.K, T.fy_host);
    int rc = GDT_OK;
    auto up = [&](void** dptr, const void* src, size_t bytes) -> int {
        cudaError_t e = cudaMalloc(dptr, bytes);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(tables)", __FILE__, __LINE__);
        e = cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(tables)", __FILE__, __LINE__);
        return GDT_OK;
    };
    if (rc == GDT_OK) rc = up((void**)&T.lutL, hL, ncell * 16);
    if (rc == GDT_OK) rc = up((void**)&T.lutAB, hAB, ncell * 32);
    if (rc == GDT_OK) {
        uint32_t* hR = (uint32_t*)malloc(ncell * 8 * sizeof(uint32_t));
        if (hR) {
            memset(hR, 0, ncell * 8 * sizeof(uint32_t));
            T.rec_ok = pack_lab_rec32(host_rgb2lab_lut, hR);
            if (T.rec_ok) rc = up((void**)&T.rec32, hR, ncell * 32);
            free(hR);
        }
    }
    if (rc == GDT_OK) rc = up((void**)&T.spline, T.spline_host, 4096 * sizeof(float));
    if (rc == GDT_OK) rc = up((void**)&T.fytab, T.fy_host, 1024 * sizeof(float));
    free(hL);
    free(hAB);
    auto make_tex = [&](cudaTextureObject_t* tex, void* ptr, size_t bytes, bool is_float) -> int {
        cudaResourceDesc rd;
        memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = ptr;
        rd.res.linear.desc = is_float ? cudaCreateChannelDesc<float4>() : cudaCreateChannelDesc<uint4>();
        rd.res.linear.sizeInBytes = bytes;
        cudaTextureDesc td;
        memset(&td, 0, sizeof(td));
        td.addressMode[0] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        cudaError_t e = cudaCreateTextureObject(tex, &rd, &td, nullptr);
        if (e != cudaSuccess) return cuda_fail(e, "cudaCreateTextureObject(tables)", __FILE__, __LINE__);
        return GDT_OK;
    };
    if (rc == GDT_OK) rc = make_tex(&T.texL, T.lutL, ncell * 16, false);
    if (rc == GDT_OK) rc = make_tex(&T.texAB, T.lutAB, ncell * 32, false);
    if (rc == GDT_OK) rc = make_tex(&T.texSpline, T.spline, 4096 * sizeof(float), true);
    if (rc == GDT_OK) rc = make_tex(&T.texFy, T.fytab, 1024 * sizeof(float), true);
    if (rc == GDT_OK) {
        // the normalisation constants of the reference's transforms (ImageNet std; 0.5 for the GAN tensors)
        const float cand[4] = {0.229f, 0.224f, 0.225f, 0.5f};
        for (float b : cand)
            if (div_by_const_ok(b) && T.n_div1 < 8 && verify_div1(b)) T.div1_std[T.n_div1++] = b;
    }
    if (rc == GDT_OK) T.ready = true;
    return rc;
}

extern "C" int gdt_is_initialised(void) { return clahe_tables_for_current_device() ? 1 : 0; }

extern "C" int gdt_debug_get_spline_table(float* host_out_4096) {
    if (!host_out_4096) return GDT_ERR_INVALID_ARGUMENT;
    static float tab[4096];
    static bool built = false;
    if (!built) { build_inv_gamma_spline(tab); built = true; }
    memcpy(host_out_4096, tab, sizeof(tab));
    return GDT_OK;
}

extern "C" size_t gdt_clahe_workspace_bytes(int n, int h, int w, int grid) {
    if (n <= 0 || h <= 0 || w <= 0 || grid < 1) return 0;
    const size_t pitch = ((size_t)w + 3) & ~(size_t)3;
    return align_up((size_t)n * h * pitch, 256) + align_up((size_t)n * h * pitch * (gdt::g_k1_chroma_f ? 8 : 4), 256) +
           align_up((size_t)n * grid * 256 * 16, 256) + 512;
}

extern "C" int gdt_clahe_u8(const uint8_t* rgb_hwc, int n, int h, int w, double clip_limit, int grid,
                            const float* host_mean, const float* host_std, float* out_chw, void* ws, size_t ws_bytes,
                            void* stream) {
    if (!host_mean || !host_std) return GDT_ERR_INVALID_ARGUMENT;
    Norm3 o, i;
    for (int c = 0; c < 3; ++c) { o.mean[c] = host_mean[c]; o.std[c] = host_std[c]; i.mean[c] = 0.f; i.std[c] = 1.f; }
    return clahe_launch<true>(rgb_hwc, n, h, w, clip_limit, grid, i, o, out_chw, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int gdt_clahe_f32(const float* in_chw, int n, int h, int w, double clip_limit, int grid,
                             const float* host_in_mean, const float* host_in_std, const float* host_out_mean,
                             const float* host_out_std, float* out_chw, void* ws, size_t ws_bytes, void* stream) {
    if (!host_in_mean || !host_in_std || !host_out_mean || !host_out_std) return GDT_ERR_INVALID_ARGUMENT;
    Norm3 o, i;
    for (int c = 0; c < 3; ++c) {
        o.mean[c] = host_out_mean[c]; o.std[c] = host_out_std[c];
        i.mean[c] = host_in_mean[c]; i.std[c] = host_in_std[c];
    }
    return clahe_launch<false>(in_chw, n, h, w, clip_limit, grid, i, o, out_chw, ws, ws_bytes, (cudaStream_t)stream);
}
