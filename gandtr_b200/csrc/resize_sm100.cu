// K5 -- dataset image geometry on the device for sm_100a: crop + `Image.thumbnail((imsize, imsize), LANCZOS)` of a decoded
// uint8 RGB image (mdir/external/cirtorch/datasets/genericdataset.py:86-97, datahelpers.py:75-82), bit-exact against
// Pillow's 8-bit resampler: optional integer box reduction (Reduce.c), then the two-pass fixed-point LANCZOS resample
// (Resample.c: horizontal pass over the rows the vertical pass needs, uint8 rounding in between, 22-bit coefficients).
// Integer / byte work, HBM- and L1-bound; the filter coefficients are computed on the host in double precision with the
// same libm calls Pillow makes (resize_math.h) when a plan is created, one plan per image geometry.
//   reduce_kernel     one thread per reduced pixel
//   resize_h5_kernel  the default horizontal pass (source rows 4-byte aligned): raw bytes of the next row batch arrive by
//                     cp.async while the current batch, split once into three byte planes, runs its tap loop; tap groups
//                     aligned to 4 input pixels: one aligned shared-memory word per channel and group, 9 dp4a per group
//   resize_h4_kernel  (any alignment) one CTA per (image, 128 output columns, band of rows): the input spans of 4 rows are staged in
//                     shared memory as aligned words; a thread owns one output pixel of each row and walks its taps four
//                     at a time: 3 word loads + funnel shifts + byte permutes gather the 4 same-channel bytes, and the
//                     22-bit coefficients, split on the host into three byte planes (c = c2 * 65536 + c1 * 256 + c0),
//                     are applied with dp4a -- 9 dp4a per 12 multiply-adds, exact integer arithmetic mod 2^32
//   resize_v4_kernel  one thread per 4 consecutive bytes of an output row: 4 tap rows are transposed in registers (8 byte
//                     permutes) and applied with 12 dp4a per 16 multiply-adds; coefficients are warp-uniform
//   resize_h_kernel / resize_v_kernel  the byte-wise forms, kept for the layouts the fast kernels do not take (vertical
//                     pass straight from a caller's strided image, rows that are not 4-byte aligned)
// Images sharing a geometry are processed by ONE launch per pass (gdt_resize_u8_batch, image index = blockIdx.z).
#include <new>

#include "common.cuh"
#include "resize_math.h"

namespace gdt {

struct ResizePlan {
    int in_w, in_h;           // the (cropped) source
    ThumbGeom g;
    int red_w, red_h;         // size after the box reduction (== in_w, in_h when fx == fy == 1)
    uint32_t red_mult[4];     // reciprocal of the box area: full, ragged last column, ragged last row, corner
    int need_h, need_v;
    int ksize_h, ksize_v;
    int ybox_first, ybox_last;
    int h_smem_bytes;         // largest staged input span of a 128-column chunk, in bytes
    int* bounds_h = nullptr;      // [out_w][2]
    int32_t* kk_h = nullptr;      // [ksize_h][out_w]   (transposed)
    int* bounds_v = nullptr;      // [out_h][2]         (first index relative to ybox_first)
    int32_t* kk_v = nullptr;      // [out_h][ksize_v]
    int groups_h = 0, groups_v = 0;   // tap groups of 4 (dp4a form)
    uint32_t* kk4_h = nullptr;    // [groups_h][3 planes][out_w]   c0 | c1 (unsigned bytes), c2 (signed bytes), 4 taps per word
    uint32_t* kk4_v = nullptr;    // [out_h][groups_v][3 planes]
    int row_smem = 0;             // bytes of one staged row of the fast horizontal kernel
    // planar form (resize_h5_kernel): tap groups aligned to 4 INPUT pixels
    int groups_h5 = 0;            // aligned groups a column's taps can touch
    int span5 = 0;                // aligned groups a 128-column chunk can touch (shared-memory words per plane and row)
    int* g0_h = nullptr;          // [out_w] first aligned group of the column's tap window
    uint32_t* kk5_h = nullptr;    // [groups_h5][3 planes][out_w], coefficients shifted to the aligned groups (zeros outside)
    int device = -1;
};

constexpr int kMaxBatch = 32;  // images per launch (pointers travel in the kernel parameter block)
struct BatchSrc {
    const uint8_t* ptr[kMaxBatch];
    unsigned long long stride[kMaxBatch];
};
constexpr int kH4Rows = 4;     // rows staged together by the fast horizontal kernel
constexpr int kH4Band = 16;    // rows per CTA

__device__ __forceinline__ int dp4a_uu(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint8_t clip8_planes(int s0, int s1, int s2) {
    // wrap-around int32 arithmetic, like Pillow's `ss` accumulator (the true sum fits)
    const int acc = (int)((unsigned)(1 << (kResizePrecisionBits - 1)) + (unsigned)s0 + ((unsigned)s1 << 8) + ((unsigned)s2 << 16));
    const int v = acc >> kResizePrecisionBits;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

constexpr int kHCols = 128;    // output columns per CTA of the horizontal pass
constexpr int kHRows = 8;      // rows per CTA of the horizontal pass

__device__ __forceinline__ uint8_t clip8(int acc) {
    const int v = acc >> kResizePrecisionBits;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

__global__ void __launch_bounds__(256)
reduce_kernel(const uint8_t* __restrict__ src, size_t src_stride, int w, int h, int fx, int fy, uint8_t* __restrict__ dst,
              int rw, int rh, uint32_t m_full, uint32_t m_lastcol, uint32_t m_lastrow, uint32_t m_corner) {
    const int ox = blockIdx.x * 256 + threadIdx.x, oy = blockIdx.y;
    if (ox >= rw) return;
    const int x0 = ox * fx, y0 = oy * fy;
    const int x1 = min(x0 + fx, w), y1 = min(y0 + fy, h);
    const bool rx = (x1 - x0) != fx, ry = (y1 - y0) != fy;
    const uint32_t mult = rx ? (ry ? m_corner : m_lastcol) : (ry ? m_lastrow : m_full);
    const uint32_t amend = (uint32_t)((x1 - x0) * (y1 - y0)) / 2u;
    uint32_t s0 = amend, s1 = amend, s2 = amend;
    for (int y = y0; y < y1; ++y) {
        const uint8_t* p = src + (size_t)y * src_stride + (size_t)x0 * 3;
        for (int x = 0; x < x1 - x0; ++x) {
            s0 += p[x * 3];
            s1 += p[x * 3 + 1];
            s2 += p[x * 3 + 2];
        }
    }
    uint8_t* o = dst + ((size_t)oy * rw + ox) * 3;
    o[0] = (uint8_t)((s0 * mult) >> 24);
    o[1] = (uint8_t)((s1 * mult) >> 24);
    o[2] = (uint8_t)((s2 * mult) >> 24);
}

__global__ void __launch_bounds__(kHCols)
resize_h_kernel(const uint8_t* __restrict__ src, size_t src_stride, int row0, int nrows, uint8_t* __restrict__ dst,
                size_t dst_stride, int out_w, const int* __restrict__ bounds, const int32_t* __restrict__ kkT) {
    extern __shared__ __align__(16) uint8_t span[];
    const int tid = threadIdx.x;
    const int xx0 = blockIdx.x * kHCols;
    const int xx = xx0 + tid;
    const int xl = min(xx0 + kHCols - 1, out_w - 1);
    const int first = __ldg(bounds + xx0 * 2);
    const int end = __ldg(bounds + xl * 2) + __ldg(bounds + xl * 2 + 1);
    const int nbytes = (end - first) * 3;
    int xmin = 0, cnt = 0;
    if (xx < out_w) {
        xmin = __ldg(bounds + xx * 2) - first;
        cnt = __ldg(bounds + xx * 2 + 1);
    }
    const int r0 = blockIdx.y * kHRows, r1 = min(r0 + kHRows, nrows);
    uint32_t* span32 = (uint32_t*)span;
    for (int r = r0; r < r1; ++r) {
        const uint8_t* row = src + (size_t)(row0 + r) * src_stride + (size_t)first * 3;
        // staged as aligned 32-bit words: the span starts `mis` bytes into the first word (rows of 3-byte pixels and
        // crop views start anywhere); the words lie inside the source row's allocation except possibly the last one,
        // which is read bytewise
        const int mis = (int)((uintptr_t)row & 3);
        const uint32_t* row32 = (const uint32_t*)(row - mis);
        const int nwords = (mis + nbytes + 3) >> 2;
        __syncthreads();                                  // the previous row's readers are done
        for (int i = tid; i < nwords - 1; i += kHCols) span32[i] = __ldg(row32 + i);
        if (tid == 0) {
            uint32_t wv = 0;
            for (int b = (nwords - 1) * 4; b < mis + nbytes; ++b) wv |= (uint32_t)__ldg(row - mis + b) << (8 * (b & 3));
            span32[nwords - 1] = wv;
        }
        __syncthreads();
        if (xx < out_w) {
            int a0 = 1 << (kResizePrecisionBits - 1), a1 = a0, a2 = a0;
            const uint8_t* p = span + mis + xmin * 3;
            for (int k = 0; k < cnt; ++k) {
                const int c = __ldg(kkT + (size_t)k * out_w + xx);
                a0 += (int)p[k * 3] * c;
                a1 += (int)p[k * 3 + 1] * c;
                a2 += (int)p[k * 3 + 2] * c;
            }
            uint8_t* o = dst + (size_t)r * dst_stride + (size_t)xx * 3;
            o[0] = clip8(a0);
            o[1] = clip8(a1);
            o[2] = clip8(a2);
        }
    }
}

// ---- fast (dp4a) kernels ------------------------------------------------------------------------------------

__global__ void __launch_bounds__(kHCols)
resize_h4_kernel(BatchSrc S, int row0, int nrows, uint8_t* __restrict__ dst, size_t dst_image_stride, size_t dst_stride,
                 int out_w, const int* __restrict__ bounds, const uint32_t* __restrict__ kk4, int groups, int row_smem) {
    extern __shared__ __align__(16) uint8_t span[];          // [kH4Rows][row_smem]
    __shared__ int mis_s[kH4Rows];
    const int tid = threadIdx.x;
    const int img = blockIdx.z;
    const uint8_t* __restrict__ src = S.ptr[img];
    const size_t src_stride = (size_t)S.stride[img];
    uint8_t* __restrict__ out = dst + (size_t)img * dst_image_stride;
    const int xx0 = blockIdx.x * kHCols;
    const int xx = xx0 + tid;
    const int xl = min(xx0 + kHCols - 1, out_w - 1);
    const int first = __ldg(bounds + xx0 * 2);
    const int end = __ldg(bounds + xl * 2) + __ldg(bounds + xl * 2 + 1);
    const int nbytes = (end - first) * 3;
    const int xmin = xx < out_w ? __ldg(bounds + xx * 2) - first : 0;
    const int band0 = blockIdx.y * kH4Band, band1 = min(band0 + kH4Band, nrows);
    for (int rb = band0; rb < band1; rb += kH4Rows) {
        const int nr = min(kH4Rows, band1 - rb);
        __syncthreads();                                      // the previous batch's readers are done
        for (int r = 0; r < nr; ++r) {
            const uint8_t* row = src + (size_t)(row0 + rb + r) * src_stride + (size_t)first * 3;
            const int mis = (int)((uintptr_t)row & 3);
            const uint32_t* row32 = (const uint32_t*)(row - mis);
            const int nwords = (mis + nbytes + 3) >> 2;
            uint32_t* s32 = (uint32_t*)(span + (size_t)r * row_smem);
            for (int i = tid; i < nwords - 1; i += kHCols) s32[i] = __ldg(row32 + i);
            if (tid == r) {                                   // the last word may reach past the allocation: bytewise
                uint32_t wv = 0;
                for (int b = (nwords - 1) * 4; b < mis + nbytes; ++b) wv |= (uint32_t)__ldg(row - mis + b) << (8 * (b & 3));
                s32[nwords - 1] = wv;
                mis_s[r] = mis;
            }
        }
        __syncthreads();
        if (xx < out_w) {
            int acc[kH4Rows][9];
            uint32_t wlast[kH4Rows];
            int wbase[kH4Rows];
            unsigned sh[kH4Rows];
#pragma unroll
            for (int r = 0; r < kH4Rows; ++r) {
#pragma unroll
                for (int a = 0; a < 9; ++a) acc[r][a] = 0;
                const int b0 = (r < nr ? mis_s[r] : 0) + xmin * 3;   // byte offset of the first tap; + 12 per group
                wbase[r] = r * (row_smem >> 2) + (b0 >> 2);
                sh[r] = (unsigned)(b0 & 3) * 8u;
                wlast[r] = ((const uint32_t*)span)[wbase[r]];
            }
            const uint32_t* s32 = (const uint32_t*)span;
            for (int g = 0; g < groups; ++g) {
                const uint32_t c0 = __ldg(kk4 + ((size_t)(g * 3 + 0)) * out_w + xx);
                const uint32_t c1 = __ldg(kk4 + ((size_t)(g * 3 + 1)) * out_w + xx);
                const uint32_t c2 = __ldg(kk4 + ((size_t)(g * 3 + 2)) * out_w + xx);
#pragma unroll
                for (int r = 0; r < kH4Rows; ++r) {
                    if (r < nr) {
                        const int wi = wbase[r] + g * 3;
                        const uint32_t w0 = wlast[r], w1 = s32[wi + 1], w2 = s32[wi + 2], w3 = s32[wi + 3];
                        wlast[r] = w3;
                        const uint32_t a0 = __funnelshift_r(w0, w1, sh[r]), a1 = __funnelshift_r(w1, w2, sh[r]),
                                       a2 = __funnelshift_r(w2, w3, sh[r]);
                        // 12 aligned bytes b0..b11 = 4 RGB pixels: gather the 4 bytes of each channel
                        const uint32_t pr = __byte_perm(__byte_perm(a0, a1, 0x0630), a2, 0x5210);   // b0 b3 b6 b9
                        const uint32_t pg = __byte_perm(__byte_perm(a0, a1, 0x0741), a2, 0x6210);   // b1 b4 b7 b10
                        const uint32_t pb = __byte_perm(__byte_perm(a0, a1, 0x0052), a2, 0x7410);   // b2 b5 b8 b11
                        acc[r][0] = dp4a_uu(pr, c0, acc[r][0]); acc[r][1] = dp4a_uu(pr, c1, acc[r][1]); acc[r][2] = dp4a_us(pr, c2, acc[r][2]);
                        acc[r][3] = dp4a_uu(pg, c0, acc[r][3]); acc[r][4] = dp4a_uu(pg, c1, acc[r][4]); acc[r][5] = dp4a_us(pg, c2, acc[r][5]);
                        acc[r][6] = dp4a_uu(pb, c0, acc[r][6]); acc[r][7] = dp4a_uu(pb, c1, acc[r][7]); acc[r][8] = dp4a_us(pb, c2, acc[r][8]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < kH4Rows; ++r) {
                if (r < nr) {
                    uint8_t* o = out + (size_t)(rb + r) * dst_stride + (size_t)xx * 3;
                    o[0] = clip8_planes(acc[r][0], acc[r][1], acc[r][2]);
                    o[1] = clip8_planes(acc[r][3], acc[r][4], acc[r][5]);
                    o[2] = clip8_planes(acc[r][6], acc[r][7], acc[r][8]);
                }
            }
        }
    }
}

// Planar form of the fast horizontal pass (source rows 4-byte aligned). The interleaved RGB bytes are split into three
// byte planes ONCE per row batch in shared memory (12 bytes -> 6 byte permutes -> one word per plane), and
// the tap groups are aligned to 4 INPUT pixels (the host shifted each column's coefficients accordingly, zeros outside its
// window): the 4 same-channel bytes of a tap group are then ONE aligned shared-memory word -- no funnel shifts or byte
// permutes in the tap loop, 3 loads + 9 dp4a per (row, group) instead of 3 + 3 + 6 + 9.
constexpr int kH5Rows = 4;     // rows per batch (accumulators: 4 x 9 registers)
constexpr int kH5Band = 32;    // rows per CTA

// The raw bytes of the NEXT row batch travel global -> shared memory with cp.async (no registers, double-buffered) while
// the current batch is split into planes and runs its tap loop, so the DRAM latency of the staging hides under the dp4a
// work.
__device__ __forceinline__ void cp_async_4(uint32_t smem_addr, const void* gptr, bool valid) {
    const int n = valid ? 4 : 0;          // src-size 0: the destination word is zero-filled, nothing is read
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_addr), "l"(gptr), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// NG > 0: the number of aligned tap groups is a compile-time constant (5 ... 8 cover every LANCZOS scale after Pillow's box
// pre-reduction): the tap loop is unrolled and the coefficient words of the next group load under the current group's dp4a.
template <int NG>
__global__ void __launch_bounds__(kHCols)
resize_h5_kernel(BatchSrc S, int row0, int nrows, int in_w, uint8_t* __restrict__ dst, size_t dst_image_stride, size_t dst_stride,
                 int out_w, const int* __restrict__ g0, const uint32_t* __restrict__ kk5, int groups_, int span) {
    const int groups = NG > 0 ? NG : groups_;
    extern __shared__ __align__(16) uint32_t h5smem[];       // planes [kH5Rows][3][span], then raw [2][kH5Rows][3 * span]
    uint32_t* planes = h5smem;
    uint32_t* raw = h5smem + kH5Rows * 3 * span;
    const int tid = threadIdx.x;
    const int img = blockIdx.z;
    const uint8_t* __restrict__ src = S.ptr[img];
    const size_t src_stride = (size_t)S.stride[img];
    uint8_t* __restrict__ out = dst + (size_t)img * dst_image_stride;
    const int xx0 = blockIdx.x * kHCols;
    const int xx = xx0 + tid;
    const int gcta = __ldg(g0 + xx0);                         // first aligned group of the chunk (g0 is non-decreasing)
    const int gmine = xx < out_w ? __ldg(g0 + xx) - gcta : 0;
    const int xl = min(xx0 + kHCols - 1, out_w - 1);
    const int ngroups_cta = __ldg(g0 + xl) + groups - gcta;   // <= span
    const int band0 = blockIdx.y * kH5Band, band1 = min(band0 + kH5Band, nrows);
    const int nwords = ngroups_cta * 3;                       // raw words per row: 12 bytes per aligned group
    const int row_words = (in_w * 3 + 3) >> 2;                // words that hold bytes of the row (rows are word-aligned)
    const int w0 = gcta * 3;                                  // first raw word of the chunk inside a row
    const uint32_t raw_s = (uint32_t)__cvta_generic_to_shared(raw);

    auto issue = [&](int rb, int buf) {                       // raw bytes of rows [rb, rb + kH5Rows) -> raw[buf]
        for (int r = 0; r < kH5Rows; ++r) {
            if (rb + r >= band1) break;
            const uint32_t* rp = (const uint32_t*)(src + (size_t)(row0 + rb + r) * src_stride) + w0;
            const uint32_t sbase = raw_s + (uint32_t)(((buf * kH5Rows + r) * 3 * span) << 2);
            for (int i = tid; i < nwords; i += kHCols)        // words past the row end are zero-filled (zero-weighted taps)
                cp_async_4(sbase + (uint32_t)(i << 2), rp + i, w0 + i < row_words);
        }
        cp_async_commit();
    };

    issue(band0, 0);
    int buf = 0;
    for (int rb = band0; rb < band1; rb += kH5Rows, buf ^= 1) {
        const int nr = min(kH5Rows, band1 - rb);
        const bool more = rb + kH5Rows < band1;
        if (more) issue(rb + kH5Rows, buf ^ 1);              // its buffer was split into planes two barriers ago
        if (more) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();                                      // this batch's bytes have landed; the previous tap loop is done
        for (int r = 0; r < nr; ++r) {
            const uint32_t* rr = raw + (size_t)(buf * kH5Rows + r) * 3 * span;
            for (int gi = tid; gi < ngroups_cta; gi += kHCols) {
                const uint32_t a0 = rr[gi * 3], a1 = rr[gi * 3 + 1], a2 = rr[gi * 3 + 2];
                uint32_t* pl = planes + (size_t)r * 3 * span + gi;
                pl[0] = __byte_perm(__byte_perm(a0, a1, 0x0630), a2, 0x5210);          // b0 b3 b6 b9
                pl[span] = __byte_perm(__byte_perm(a0, a1, 0x0741), a2, 0x6210);       // b1 b4 b7 b10
                pl[2 * span] = __byte_perm(__byte_perm(a0, a1, 0x0052), a2, 0x7410);   // b2 b5 b8 b11
            }
        }
        __syncthreads();
        if (xx < out_w) {
            int acc[kH5Rows][9];
#pragma unroll
            for (int r = 0; r < kH5Rows; ++r)
#pragma unroll
                for (int a = 0; a < 9; ++a) acc[r][a] = 0;
            const uint32_t* kc = kk5 + xx;
            const uint32_t* pm = planes + gmine;
#pragma unroll
            for (int g = 0; g < (NG > 0 ? NG : groups); ++g) {
                const uint32_t c0 = __ldg(kc), c1 = __ldg(kc + out_w), c2 = __ldg(kc + 2 * out_w);
                kc += 3 * (size_t)out_w;
#pragma unroll
                for (int r = 0; r < kH5Rows; ++r) {
                    // rows past the batch read stale but in-bounds words; their sums are never stored
                    const uint32_t pr = pm[(r * 3 + 0) * span + g], pg = pm[(r * 3 + 1) * span + g], pb = pm[(r * 3 + 2) * span + g];
                    acc[r][0] = dp4a_uu(pr, c0, acc[r][0]); acc[r][1] = dp4a_uu(pr, c1, acc[r][1]); acc[r][2] = dp4a_us(pr, c2, acc[r][2]);
                    acc[r][3] = dp4a_uu(pg, c0, acc[r][3]); acc[r][4] = dp4a_uu(pg, c1, acc[r][4]); acc[r][5] = dp4a_us(pg, c2, acc[r][5]);
                    acc[r][6] = dp4a_uu(pb, c0, acc[r][6]); acc[r][7] = dp4a_uu(pb, c1, acc[r][7]); acc[r][8] = dp4a_us(pb, c2, acc[r][8]);
                }
            }
#pragma unroll
            for (int r = 0; r < kH5Rows; ++r) {
                if (r < nr) {
                    uint8_t* o = out + (size_t)(rb + r) * dst_stride + (size_t)xx * 3;
                    o[0] = clip8_planes(acc[r][0], acc[r][1], acc[r][2]);
                    o[1] = clip8_planes(acc[r][3], acc[r][4], acc[r][5]);
                    o[2] = clip8_planes(acc[r][6], acc[r][7], acc[r][8]);
                }
            }
        }
    }
}

// Aligned layouts only: src rows 4-byte aligned and padded to whole words, dst rows 4-byte aligned.
__global__ void __launch_bounds__(256)
resize_v4_kernel(const uint8_t* __restrict__ src, size_t src_image_stride, size_t src_stride, int src_rows,
                 uint8_t* __restrict__ dst, size_t dst_image_stride, size_t dst_stride, int row_bytes,
                 const int* __restrict__ bounds, const uint32_t* __restrict__ kk4, int groups) {
    const int j = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (j >= row_bytes) return;
    const int yy = blockIdx.y, img = blockIdx.z;
    const uint8_t* __restrict__ in = src + (size_t)img * src_image_stride + j;
    const int ymin = __ldg(bounds + yy * 2);
    const uint32_t* __restrict__ kc = kk4 + (size_t)yy * groups * 3;
    int acc[4][3];
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[b][0] = acc[b][1] = acc[b][2] = 0;
    for (int g = 0; g < groups; ++g) {
        const uint32_t c0 = __ldg(kc + g * 3), c1 = __ldg(kc + g * 3 + 1), c2 = __ldg(kc + g * 3 + 2);
        uint32_t w[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int row = min(ymin + g * 4 + t, src_rows - 1);     // taps past the window have zero coefficients
            w[t] = __ldg((const uint32_t*)(in + (size_t)row * src_stride));
        }
        // 4 x 4 byte transpose: column b of the four tap rows
        const uint32_t t0 = __byte_perm(w[0], w[1], 0x5140), t1 = __byte_perm(w[2], w[3], 0x5140);
        const uint32_t t2 = __byte_perm(w[0], w[1], 0x7362), t3 = __byte_perm(w[2], w[3], 0x7362);
        const uint32_t col[4] = {__byte_perm(t0, t1, 0x5410), __byte_perm(t0, t1, 0x7632), __byte_perm(t2, t3, 0x5410),
                                 __byte_perm(t2, t3, 0x7632)};
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            acc[b][0] = dp4a_uu(col[b], c0, acc[b][0]);
            acc[b][1] = dp4a_uu(col[b], c1, acc[b][1]);
            acc[b][2] = dp4a_us(col[b], c2, acc[b][2]);
        }
    }
    uint8_t* o = dst + (size_t)img * dst_image_stride + (size_t)yy * dst_stride + j;
    const uint32_t v = (uint32_t)clip8_planes(acc[0][0], acc[0][1], acc[0][2]) |
                       ((uint32_t)clip8_planes(acc[1][0], acc[1][1], acc[1][2]) << 8) |
                       ((uint32_t)clip8_planes(acc[2][0], acc[2][1], acc[2][2]) << 16) |
                       ((uint32_t)clip8_planes(acc[3][0], acc[3][1], acc[3][2]) << 24);
    if (j + 4 <= row_bytes) {
        *(uint32_t*)o = v;
    } else {
        for (int b = 0; b < row_bytes - j; ++b) o[b] = (uint8_t)(v >> (8 * b));
    }
}

template <bool ALIGNED_IN, bool ALIGNED_OUT>
__global__ void __launch_bounds__(256)
resize_v_kernel(const uint8_t* __restrict__ src, size_t src_stride, uint8_t* __restrict__ dst, size_t dst_stride,
                int row_bytes, const int* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize) {
    const int j = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (j >= row_bytes) return;
    const int yy = blockIdx.y;
    const int ymin = __ldg(bounds + yy * 2), cnt = __ldg(bounds + yy * 2 + 1);
    const int32_t* k = kk + (size_t)yy * ksize;
    const int nb = min(4, row_bytes - j);
    int a[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) a[b] = 1 << (kResizePrecisionBits - 1);
    for (int t = 0; t < cnt; ++t) {
        const int c = __ldg(k + t);
        const uint8_t* p = src + (size_t)(ymin + t) * src_stride + j;
        uint32_t wv;
        if (ALIGNED_IN) {
            wv = __ldg((const uint32_t*)p);               // padded rows: reading past row_bytes stays inside the row
        } else {
            wv = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (b < nb) wv |= (uint32_t)__ldg(p + b) << (8 * b);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) a[b] += (int)((wv >> (8 * b)) & 255u) * c;
    }
    uint8_t* o = dst + (size_t)yy * dst_stride + j;
    if (ALIGNED_OUT && nb == 4) {
        *(uint32_t*)o = (uint32_t)clip8(a[0]) | ((uint32_t)clip8(a[1]) << 8) | ((uint32_t)clip8(a[2]) << 16) |
                        ((uint32_t)clip8(a[3]) << 24);
    } else {
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (b < nb) o[b] = clip8(a[b]);
    }
}

static size_t tmp_stride(const ResizePlan* p) { return align_up((size_t)p->g.out_w * 3, 4); }

}  // namespace gdt

using namespace gdt;

extern "C" int gdt_thumbnail_geometry(int w, int h, double imsize, int* out_w, int* out_h, int* fx, int* fy) {
    if (w <= 0 || h <= 0 || !(imsize >= 1.0) || !out_w || !out_h) return GDT_ERR_INVALID_ARGUMENT;
    const ThumbGeom g = thumbnail_geometry(w, h, imsize);
    *out_w = g.out_w;
    *out_h = g.out_h;
    if (fx) *fx = g.fx;
    if (fy) *fy = g.fy;
    return g.resize;
}

extern "C" int gdt_debug_resize_coeffs(int in_size, float in0, float in1, int out_size, int* ksize, int* bounds,
                                       int32_t* kk, size_t kk_capacity) {
    if (in_size <= 0 || out_size <= 0 || !ksize || !bounds || !kk) return GDT_ERR_INVALID_ARGUMENT;
    std::vector<int> b;
    std::vector<int32_t> k;
    *ksize = resize_coeffs(in_size, in0, in1, out_size, b, k);
    if (k.size() > kk_capacity) return GDT_ERR_WORKSPACE_TOO_SMALL;
    memcpy(bounds, b.data(), b.size() * sizeof(int));
    memcpy(kk, k.data(), k.size() * sizeof(int32_t));
    return GDT_OK;
}

extern "C" void gdt_resize_plan_destroy(gdt_resize_plan* plan_) {
    ResizePlan* p = (ResizePlan*)plan_;
    if (!p) return;
    cudaFree(p->bounds_h);
    cudaFree(p->kk_h);
    cudaFree(p->bounds_v);
    cudaFree(p->kk_v);
    cudaFree(p->kk4_h);
    cudaFree(p->kk4_v);
    cudaFree(p->g0_h);
    cudaFree(p->kk5_h);
    delete p;
}

extern "C" int gdt_resize_plan_create(int in_w, int in_h, double imsize, gdt_resize_plan** plan_out) {
    if (!plan_out || in_w <= 0 || in_h <= 0 || !(imsize >= 1.0)) return GDT_ERR_INVALID_ARGUMENT;
    *plan_out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return GDT_ERR_NO_DEVICE;
    }
    ResizePlan* p = new (std::nothrow) ResizePlan();
    if (!p) return GDT_ERR_INVALID_ARGUMENT;
    GDT_CUDA(cudaGetDevice(&p->device));
    p->in_w = in_w;
    p->in_h = in_h;
    p->g = thumbnail_geometry(in_w, in_h, imsize);
    const ThumbGeom& g = p->g;
    p->red_w = (in_w + g.fx - 1) / g.fx;
    p->red_h = (in_h + g.fy - 1) / g.fy;
    p->need_h = p->need_v = 0;
    if (g.resize) {
        const int lw = in_w - (p->red_w - 1) * g.fx, lh = in_h - (p->red_h - 1) * g.fy;   // ragged last column / row
        p->red_mult[0] = reduce_multiplier(g.fx * g.fy);
        p->red_mult[1] = reduce_multiplier(lw * g.fy);
        p->red_mult[2] = reduce_multiplier(g.fx * lh);
        p->red_mult[3] = reduce_multiplier(lw * lh);
        // Image.resize after the reduction: box = (0, 0, w / fx, h / fy) as C floats
        const float bx1 = (float)((double)in_w / g.fx), by1 = (float)((double)in_h / g.fy);
        p->need_h = g.out_w != p->red_w || bx1 != (float)g.out_w;
        p->need_v = g.out_h != p->red_h || by1 != (float)g.out_h;
        std::vector<int> bh, bv;
        std::vector<int32_t> kh, kv;
        p->ksize_h = resize_coeffs(p->red_w, 0.0f, bx1, g.out_w, bh, kh);
        p->ksize_v = resize_coeffs(p->red_h, 0.0f, by1, g.out_h, bv, kv);
        p->ybox_first = bv[0];
        p->ybox_last = bv[(size_t)g.out_h * 2 - 2] + bv[(size_t)g.out_h * 2 - 1];
        if (p->need_h)
            for (int i = 0; i < g.out_h; ++i) bv[(size_t)i * 2] -= p->ybox_first;
        std::vector<int32_t> khT((size_t)p->ksize_h * g.out_w);
        for (int xx = 0; xx < g.out_w; ++xx)
            for (int k = 0; k < p->ksize_h; ++k) khT[(size_t)k * g.out_w + xx] = kh[(size_t)xx * p->ksize_h + k];
        int span = 0;
        for (int xx0 = 0; xx0 < g.out_w; xx0 += kHCols) {
            const int xl = xx0 + kHCols - 1 < g.out_w ? xx0 + kHCols - 1 : g.out_w - 1;
            const int s = bh[(size_t)xl * 2] + bh[(size_t)xl * 2 + 1] - bh[(size_t)xx0 * 2];
            if (s > span) span = s;
        }
        p->h_smem_bytes = (int)align_up((size_t)span * 3 + 8, 16);      // + up to 3 bytes of misalignment, whole words
        auto up = [&](void** d, const void* hsrc, size_t bytes) -> int {
            cudaError_t e = cudaMalloc(d, bytes);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(resize plan)", __FILE__, __LINE__);
            e = cudaMemcpy(*d, hsrc, bytes, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(resize plan)", __FILE__, __LINE__);
            return GDT_OK;
        };
        // dp4a form: taps in groups of 4, each 22-bit coefficient split into three byte planes
        // c = c2 * 65536 + c1 * 256 + c0 with c0, c1 in [0, 255] and c2 = c >> 16 (signed, |c2| <= 64)
        auto plane = [](int32_t c, int pl) -> uint32_t {
            return pl == 0 ? (uint32_t)(c & 255) : (pl == 1 ? (uint32_t)((c >> 8) & 255) : (uint32_t)((c >> 16) & 255));
        };
        p->groups_h = (p->ksize_h + 3) / 4;
        p->groups_v = (p->ksize_v + 3) / 4;
        std::vector<uint32_t> k4h((size_t)p->groups_h * 3 * g.out_w, 0u), k4v((size_t)g.out_h * p->groups_v * 3, 0u);
        for (int xx = 0; xx < g.out_w; ++xx)
            for (int k = 0; k < bh[(size_t)xx * 2 + 1]; ++k)
                for (int pl = 0; pl < 3; ++pl)
                    k4h[((size_t)(k / 4) * 3 + pl) * g.out_w + xx] |= plane(kh[(size_t)xx * p->ksize_h + k], pl) << (8 * (k & 3));
        for (int yy = 0; yy < g.out_h; ++yy)
            for (int k = 0; k < bv[(size_t)yy * 2 + 1]; ++k)
                for (int pl = 0; pl < 3; ++pl)
                    k4v[((size_t)yy * p->groups_v + k / 4) * 3 + pl] |= plane(kv[(size_t)yy * p->ksize_v + k], pl) << (8 * (k & 3));
        // planar form: tap groups aligned to 4 input pixels
        std::vector<int> g0v((size_t)g.out_w);
        p->groups_h5 = 1;
        for (int xx = 0; xx < g.out_w; ++xx) {
            const int xmin = bh[(size_t)xx * 2], cnt = bh[(size_t)xx * 2 + 1];
            g0v[xx] = xmin >> 2;
            const int ng = ((xmin + cnt + 3) >> 2) - (xmin >> 2);
            if (ng > p->groups_h5) p->groups_h5 = ng;
        }
        std::vector<uint32_t> k5h((size_t)p->groups_h5 * 3 * g.out_w, 0u);
        for (int xx = 0; xx < g.out_w; ++xx) {
            const int xmin = bh[(size_t)xx * 2], cnt = bh[(size_t)xx * 2 + 1];
            for (int k = 0; k < cnt; ++k) {
                const int pos = xmin + k - (g0v[xx] << 2);           // tap position inside the column's aligned groups
                for (int pl = 0; pl < 3; ++pl)
                    k5h[((size_t)(pos >> 2) * 3 + pl) * g.out_w + xx] |= plane(kh[(size_t)xx * p->ksize_h + k], pl) << (8 * (pos & 3));
            }
        }
        p->span5 = 1;
        for (int xx0 = 0; xx0 < g.out_w; xx0 += kHCols) {
            const int xl = xx0 + kHCols - 1 < g.out_w ? xx0 + kHCols - 1 : g.out_w - 1;
            const int sg = g0v[xl] + p->groups_h5 - g0v[xx0];
            if (sg > p->span5) p->span5 = sg;
        }
        // one staged row: the span, up to 3 bytes of misalignment, the over-read of the last (zero-weighted) tap groups
        p->row_smem = (int)align_up((size_t)span * 3 + 8 + (size_t)p->groups_h * 12 + 16, 16);
        int rc = up((void**)&p->bounds_h, bh.data(), bh.size() * sizeof(int));
        if (rc == GDT_OK) rc = up((void**)&p->kk_h, khT.data(), khT.size() * sizeof(int32_t));
        if (rc == GDT_OK) rc = up((void**)&p->bounds_v, bv.data(), bv.size() * sizeof(int));
        if (rc == GDT_OK) rc = up((void**)&p->kk_v, kv.data(), kv.size() * sizeof(int32_t));
        if (rc == GDT_OK) rc = up((void**)&p->kk4_h, k4h.data(), k4h.size() * sizeof(uint32_t));
        if (rc == GDT_OK) rc = up((void**)&p->kk4_v, k4v.data(), k4v.size() * sizeof(uint32_t));
        if (rc == GDT_OK) rc = up((void**)&p->g0_h, g0v.data(), g0v.size() * sizeof(int));
        if (rc == GDT_OK) rc = up((void**)&p->kk5_h, k5h.data(), k5h.size() * sizeof(uint32_t));
        if (rc != GDT_OK) {
            gdt_resize_plan_destroy((gdt_resize_plan*)p);
            return rc;
        }
    }
    *plan_out = (gdt_resize_plan*)p;
    return GDT_OK;
}

extern "C" int gdt_resize_plan_info(const gdt_resize_plan* plan_, int* out_w, int* out_h, int* fx, int* fy) {
    const ResizePlan* p = (const ResizePlan*)plan_;
    if (!p) return GDT_ERR_INVALID_ARGUMENT;
    if (out_w) *out_w = p->g.out_w;
    if (out_h) *out_h = p->g.out_h;
    if (fx) *fx = p->g.fx;
    if (fy) *fy = p->g.fy;
    return GDT_OK;
}

static size_t resize_image_ws(const ResizePlan* p) {
    size_t n = 0;
    if (p->g.fx > 1 || p->g.fy > 1) n += align_up((size_t)p->red_w * p->red_h * 3, 256);
    if (p->need_h && p->need_v) n += align_up(tmp_stride(p) * (size_t)(p->ybox_last - p->ybox_first), 256);
    return n;
}

extern "C" size_t gdt_resize_batch_workspace_bytes(const gdt_resize_plan* plan_, int n) {
    const ResizePlan* p = (const ResizePlan*)plan_;
    if (!p || n <= 0) return 0;
    const int chunk = n < kMaxBatch ? n : kMaxBatch;         // chunks of kMaxBatch images reuse the workspace
    return 512 + (size_t)chunk * resize_image_ws(p);
}

extern "C" size_t gdt_resize_workspace_bytes(const gdt_resize_plan* plan_) { return gdt_resize_batch_workspace_bytes(plan_, 1); }

static int g_k5_force_bytewise = 0;      // debug (gdt_debug_k5_bytewise): 1 = always the byte-wise kernels (A/B, parity)
static int g_k5_planar = 1;              // debug (gdt_debug_k5_planar): 0 = the interleaved dp4a kernel instead of the planar one

extern "C" int gdt_debug_k5_planar(int on) {
    g_k5_planar = on ? 1 : 0;
    return GDT_OK;
}

extern "C" int gdt_debug_k5_bytewise(int on) {
    g_k5_force_bytewise = on ? 1 : 0;
    return GDT_OK;
}

// One chunk (<= kMaxBatch images of one geometry): every pass is ONE launch over all images.
static int resize_chunk(const ResizePlan* p, const uint8_t* const* srcs, const size_t* strides, int n, uint8_t* dst, void* ws,
                        size_t ws_bytes, cudaStream_t stream) {
    const ThumbGeom& g = p->g;
    const size_t out_stride = (size_t)g.out_w * 3, out_image = out_stride * (size_t)g.out_h;
    if (!g.resize || (!p->need_h && !p->need_v && g.fx == 1 && g.fy == 1)) {   // crop only: strided copies
        for (int i = 0; i < n; ++i)
            GDT_CUDA(cudaMemcpy2DAsync(dst + (size_t)i * out_image, out_stride, srcs[i], strides[i], out_stride, (size_t)g.out_h,
                                       cudaMemcpyDeviceToDevice, stream));
        return GDT_OK;
    }
    Workspace W(ws, ws_bytes);
    BatchSrc cur;
    for (int i = 0; i < n; ++i) { cur.ptr[i] = srcs[i]; cur.stride[i] = strides[i]; }
    for (int i = n; i < kMaxBatch; ++i) { cur.ptr[i] = nullptr; cur.stride[i] = 0; }
    if (g.fx > 1 || g.fy > 1) {
        const size_t red_bytes = align_up((size_t)p->red_w * p->red_h * 3, 256);
        uint8_t* red = W.take<uint8_t>(red_bytes * n);
        if (!W.ok()) return GDT_ERR_WORKSPACE_TOO_SMALL;
        const bool last = !p->need_h && !p->need_v;
        for (int i = 0; i < n; ++i) {
            uint8_t* rdst = last ? dst + (size_t)i * out_image : red + (size_t)i * red_bytes;
            reduce_kernel<<<dim3(ceil_div(p->red_w, 256), p->red_h), 256, 0, stream>>>(
                srcs[i], strides[i], p->in_w, p->in_h, g.fx, g.fy, rdst, p->red_w, p->red_h, p->red_mult[0], p->red_mult[1],
                p->red_mult[2], p->red_mult[3]);
            cur.ptr[i] = rdst;
            cur.stride[i] = (size_t)p->red_w * 3;
        }
        GDT_LAUNCH_CHECK();
        if (last) return GDT_OK;
    }
    const uint8_t* vsrc = nullptr;       // input of the vertical pass when it comes from the horizontal one
    size_t vsrc_image = 0, vsrc_stride = 0;
    int vsrc_rows = 0;
    if (p->need_h) {
        const int nrows = p->ybox_last - p->ybox_first;
        uint8_t* hdst = dst;
        size_t hstride = out_stride, himage = out_image;
        if (p->need_v) {
            hstride = tmp_stride(p);
            himage = align_up(hstride * (size_t)nrows, 256);
            hdst = W.take<uint8_t>(himage * n);
            if (!W.ok()) return GDT_ERR_WORKSPACE_TOO_SMALL;
        }
        const size_t smem4 = (size_t)p->row_smem * kH4Rows;
        const size_t smem5 = (size_t)3 * kH5Rows * 3 * p->span5 * sizeof(uint32_t);      // planes + two raw buffers
        bool aligned_src = true;                         // planar form: rows are copied and read as aligned words
        for (int i = 0; i < n; ++i) aligned_src = aligned_src && (((uintptr_t)cur.ptr[i]) & 3) == 0 && (cur.stride[i] & 3) == 0;
        if (!g_k5_force_bytewise && g_k5_planar && aligned_src && smem5 <= 200 * 1024) {
            static size_t attr5[32] = {0};
            size_t& a5 = attr5[current_device_slot()];
            const dim3 grid5(ceil_div(g.out_w, kHCols), ceil_div(nrows, kH5Band), n);
#define GDT_H5(NG_)                                                                                                       \
    do {                                                                                                                  \
        if (smem5 > 48 * 1024 && smem5 > a5) {                                                                            \
            GDT_CUDA(cudaFuncSetAttribute(resize_h5_kernel<NG_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5)); \
        }                                                                                                                 \
        resize_h5_kernel<NG_><<<grid5, kHCols, smem5, stream>>>(cur, p->ybox_first, nrows, p->red_w, hdst, himage, hstride, \
                                                                g.out_w, p->g0_h, p->kk5_h, p->groups_h5, p->span5);       \
    } while (0)
            switch (p->groups_h5) {
                case 5: GDT_H5(5); break;
                case 6: GDT_H5(6); break;
                case 7: GDT_H5(7); break;
                case 8: GDT_H5(8); break;
                default: GDT_H5(0); break;
            }
#undef GDT_H5
            (void)a5;
        } else if (!g_k5_force_bytewise && smem4 <= 200 * 1024) {
            static size_t attr4[32] = {0};
            size_t& a4 = attr4[current_device_slot()];
            if (smem4 > 48 * 1024 && smem4 > a4) {
                GDT_CUDA(cudaFuncSetAttribute(resize_h4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
                a4 = smem4;
            }
            resize_h4_kernel<<<dim3(ceil_div(g.out_w, kHCols), ceil_div(nrows, kH4Band), n), kHCols, smem4, stream>>>(
                cur, p->ybox_first, nrows, hdst, himage, hstride, g.out_w, p->bounds_h, p->kk4_h, p->groups_h, p->row_smem);
        } else {
            if (p->h_smem_bytes > 200 * 1024) return GDT_ERR_UNSUPPORTED;
            if (p->h_smem_bytes > 48 * 1024)
                GDT_CUDA(cudaFuncSetAttribute(resize_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p->h_smem_bytes));
            for (int i = 0; i < n; ++i)
                resize_h_kernel<<<dim3(ceil_div(g.out_w, kHCols), ceil_div(nrows, kHRows)), kHCols, p->h_smem_bytes, stream>>>(
                    cur.ptr[i], (size_t)cur.stride[i], p->ybox_first, nrows, hdst + (size_t)i * himage, hstride, g.out_w,
                    p->bounds_h, p->kk_h);
        }
        GDT_LAUNCH_CHECK();
        vsrc = hdst; vsrc_image = himage; vsrc_stride = hstride; vsrc_rows = nrows;
        for (int i = 0; i < n; ++i) { cur.ptr[i] = hdst + (size_t)i * himage; cur.stride[i] = hstride; }
    }
    if (p->need_v) {
        const int row_bytes = g.out_w * 3;
        const bool aout = (((uintptr_t)dst) & 3) == 0 && (out_stride & 3) == 0;
        if (vsrc && aout && !g_k5_force_bytewise) {
            // tmp rows are 4-byte aligned and padded to whole words; the image bases are 256-byte aligned
            resize_v4_kernel<<<dim3(ceil_div(ceil_div(row_bytes, 4), 256), g.out_h, n), 256, 0, stream>>>(
                vsrc, vsrc_image, vsrc_stride, vsrc_rows, dst, out_image, out_stride, row_bytes, p->bounds_v, p->kk4_v, p->groups_v);
        } else {
            dim3 grid(ceil_div(ceil_div(row_bytes, 4), 256), g.out_h);
            for (int i = 0; i < n; ++i) {
                const uint8_t* c = cur.ptr[i];
                const size_t cs = (size_t)cur.stride[i];
                uint8_t* o = dst + (size_t)i * out_image;
                const bool ain = (((uintptr_t)c) & 3) == 0 && (cs & 3) == 0 && p->need_h;      // padded tmp rows only
                const bool ao = (((uintptr_t)o) & 3) == 0 && (out_stride & 3) == 0;
                if (ain && ao)
                    resize_v_kernel<true, true><<<grid, 256, 0, stream>>>(c, cs, o, out_stride, row_bytes, p->bounds_v, p->kk_v, p->ksize_v);
                else if (ain)
                    resize_v_kernel<true, false><<<grid, 256, 0, stream>>>(c, cs, o, out_stride, row_bytes, p->bounds_v, p->kk_v, p->ksize_v);
                else if (ao)
                    resize_v_kernel<false, true><<<grid, 256, 0, stream>>>(c, cs, o, out_stride, row_bytes, p->bounds_v, p->kk_v, p->ksize_v);
                else
                    resize_v_kernel<false, false><<<grid, 256, 0, stream>>>(c, cs, o, out_stride, row_bytes, p->bounds_v, p->kk_v, p->ksize_v);
            }
        }
        GDT_LAUNCH_CHECK();
    }
    return GDT_OK;
}

extern "C" int gdt_resize_u8_batch(const gdt_resize_plan* plan_, const uint8_t* const* host_srcs, const size_t* host_strides,
                                   int n, uint8_t* dst, void* ws, size_t ws_bytes, void* stream_) {
    const ResizePlan* p = (const ResizePlan*)plan_;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!p || !host_srcs || !host_strides || !dst || n <= 0) return GDT_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < n; ++i)
        if (!host_srcs[i] || host_strides[i] < (size_t)p->in_w * 3) return GDT_ERR_INVALID_ARGUMENT;
    int dev = -1;
    GDT_CUDA(cudaGetDevice(&dev));
    if (dev != p->device) return GDT_ERR_INVALID_ARGUMENT;
    const size_t need = gdt_resize_batch_workspace_bytes(plan_, n);
    if (ws_bytes < need || (!ws && need > 512)) return GDT_ERR_WORKSPACE_TOO_SMALL;
    const size_t out_image = (size_t)p->g.out_w * 3 * (size_t)p->g.out_h;
    for (int i0 = 0; i0 < n; i0 += kMaxBatch) {
        const int m = n - i0 < kMaxBatch ? n - i0 : kMaxBatch;
        const int rc = resize_chunk(p, host_srcs + i0, host_strides + i0, m, dst + (size_t)i0 * out_image, ws, ws_bytes, stream);
        if (rc != GDT_OK) return rc;
    }
    return GDT_OK;
}

extern "C" int gdt_resize_u8(const gdt_resize_plan* plan_, const uint8_t* src, size_t src_stride, uint8_t* dst, void* ws,
                             size_t ws_bytes, void* stream_) {
    if (!src) return GDT_ERR_INVALID_ARGUMENT;
    return gdt_resize_u8_batch(plan_, &src, &src_stride, 1, dst, ws, ws_bytes, stream_);
}
