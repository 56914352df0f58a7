// K5 -- dataset image geometry on the device for sm_100a: crop + `Image.thumbnail((imsize, imsize), LANCZOS)` of a decoded
// uint8 RGB image (mdir/external/cirtorch/datasets/genericdataset.py:86-97, datahelpers.py:75-82), bit-exact against
// Pillow's 8-bit resampler: optional integer box reduction (Reduce.c), then the two-pass fixed-point LANCZOS resample
// (Resample.c: horizontal pass over the rows the vertical pass needs, uint8 rounding in between, 22-bit coefficients).
// Integer / byte work, HBM- and L1-bound; the filter coefficients are computed on the host in double precision with the
// same libm calls Pillow makes (resize_math.h) when a plan is created, one plan per image geometry.
//   reduce_kernel    one thread per reduced pixel
//   resize_h_kernel  one CTA per (128 output columns, band of rows): the input span of the CTA's columns is staged in
//                    shared memory row by row, coefficients read transposed (coalesced)
//   resize_v_kernel  one thread per 4 consecutive bytes of an output row (rows are flat byte arrays: every channel of
//                    every column shares the row's coefficients), taps read as aligned 32-bit words
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
    int device = -1;
};

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
        int rc = up((void**)&p->bounds_h, bh.data(), bh.size() * sizeof(int));
        if (rc == GDT_OK) rc = up((void**)&p->kk_h, khT.data(), khT.size() * sizeof(int32_t));
        if (rc == GDT_OK) rc = up((void**)&p->bounds_v, bv.data(), bv.size() * sizeof(int));
        if (rc == GDT_OK) rc = up((void**)&p->kk_v, kv.data(), kv.size() * sizeof(int32_t));
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

extern "C" size_t gdt_resize_workspace_bytes(const gdt_resize_plan* plan_) {
    const ResizePlan* p = (const ResizePlan*)plan_;
    if (!p) return 0;
    size_t n = 512;
    if (p->g.fx > 1 || p->g.fy > 1) n += align_up((size_t)p->red_w * p->red_h * 3, 256);
    if (p->need_h && p->need_v) n += align_up(tmp_stride(p) * (size_t)(p->ybox_last - p->ybox_first), 256);
    return n;
}

extern "C" int gdt_resize_u8(const gdt_resize_plan* plan_, const uint8_t* src, size_t src_stride, uint8_t* dst, void* ws,
                             size_t ws_bytes, void* stream_) {
    const ResizePlan* p = (const ResizePlan*)plan_;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!p || !src || !dst || src_stride < (size_t)p->in_w * 3) return GDT_ERR_INVALID_ARGUMENT;
    int dev = -1;
    GDT_CUDA(cudaGetDevice(&dev));
    if (dev != p->device) return GDT_ERR_INVALID_ARGUMENT;
    const ThumbGeom& g = p->g;
    const size_t out_stride = (size_t)g.out_w * 3;
    if (!g.resize || (!p->need_h && !p->need_v && g.fx == 1 && g.fy == 1)) {   // crop only: strided copy
        GDT_CUDA(cudaMemcpy2DAsync(dst, out_stride, src, src_stride, out_stride, (size_t)g.out_h, cudaMemcpyDeviceToDevice, stream));
        return GDT_OK;
    }
    if (ws_bytes < gdt_resize_workspace_bytes(plan_) || (!ws && gdt_resize_workspace_bytes(plan_) > 512))
        return GDT_ERR_WORKSPACE_TOO_SMALL;
    Workspace W(ws, ws_bytes);
    const uint8_t* cur = src;
    size_t cur_stride = src_stride;
    if (g.fx > 1 || g.fy > 1) {
        uint8_t* red = W.take<uint8_t>((size_t)p->red_w * p->red_h * 3);
        if (!W.ok()) return GDT_ERR_WORKSPACE_TOO_SMALL;
        const bool last = !p->need_h && !p->need_v;
        uint8_t* rdst = last ? dst : red;
        reduce_kernel<<<dim3(ceil_div(p->red_w, 256), p->red_h), 256, 0, stream>>>(
            src, src_stride, p->in_w, p->in_h, g.fx, g.fy, rdst, p->red_w, p->red_h, p->red_mult[0], p->red_mult[1],
            p->red_mult[2], p->red_mult[3]);
        GDT_LAUNCH_CHECK();
        if (last) return GDT_OK;
        cur = red;
        cur_stride = (size_t)p->red_w * 3;
    }
    if (p->need_h) {
        const int nrows = p->ybox_last - p->ybox_first;
        uint8_t* hdst = dst;
        size_t hstride = out_stride;
        if (p->need_v) {
            hstride = tmp_stride(p);
            hdst = W.take<uint8_t>(hstride * (size_t)nrows);
            if (!W.ok()) return GDT_ERR_WORKSPACE_TOO_SMALL;
        }
        if (p->h_smem_bytes > 48 * 1024) {
            if (p->h_smem_bytes > 200 * 1024) return GDT_ERR_UNSUPPORTED;
            GDT_CUDA(cudaFuncSetAttribute(resize_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p->h_smem_bytes));
        }
        resize_h_kernel<<<dim3(ceil_div(g.out_w, kHCols), ceil_div(nrows, kHRows)), kHCols, p->h_smem_bytes, stream>>>(
            cur, cur_stride, p->ybox_first, nrows, hdst, hstride, g.out_w, p->bounds_h, p->kk_h);
        GDT_LAUNCH_CHECK();
        cur = hdst;
        cur_stride = hstride;
    }
    if (p->need_v) {
        const int row_bytes = g.out_w * 3;
        const bool ain = (((uintptr_t)cur) & 3) == 0 && (cur_stride & 3) == 0 && p->need_h;   // padded tmp rows only
        const bool aout = (((uintptr_t)dst) & 3) == 0 && (out_stride & 3) == 0;
        dim3 grid(ceil_div(ceil_div(row_bytes, 4), 256), g.out_h);
        if (ain && aout)
            resize_v_kernel<true, true><<<grid, 256, 0, stream>>>(cur, cur_stride, dst, out_stride, row_bytes, p->bounds_v, p->kk_v, p->ksize_v);
        else if (ain)
            resize_v_kernel<true, false><<<grid, 256, 0, stream>>>(cur, cur_stride, dst, out_stride, row_bytes, p->bounds_v, p->kk_v, p->ksize_v);
        else if (aout)
            resize_v_kernel<false, true><<<grid, 256, 0, stream>>>(cur, cur_stride, dst, out_stride, row_bytes, p->bounds_v, p->kk_v, p->ksize_v);
        else
            resize_v_kernel<false, false><<<grid, 256, 0, stream>>>(cur, cur_stride, dst, out_stride, row_bytes, p->bounds_v, p->kk_v, p->ksize_v);
        GDT_LAUNCH_CHECK();
    }
    return GDT_OK;
}
