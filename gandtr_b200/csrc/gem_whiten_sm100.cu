// K2 -- GeM pooling + L2N + multi-scale generalised mean + learned whitening for sm_100a.
// Replaces (reference, fp32 everywhere):
//   LF.gem   mdir/external/cirtorch/layers/functional.py:21-22   avg_pool2d(x.clamp(min=eps).pow(p)).pow(1/p)
//   LF.l2n   mdir/external/cirtorch/layers/functional.py:130-131 x / (||x||_2 + eps)
//   CirMultiscaleAggregation.aggregate_tensor  mdir/components/data/wrapper.py:235-245
//   CirtorchWhiten.postprocess                 mdir/components/data/wrapper.py:320-322
//
// gem_pool_kernel is the only kernel that touches the feature maps: one warp per (scale, image,
// channel) row, 128-bit coalesced streaming loads (L1 no-allocate), fp32 accumulation, one value out
// per row. It is HBM-bound: algorithmic bytes = 4*c*sum_s(h_s*w_s) per image. The remaining kernels
// work on [n][c] vectors (KBs per image): per-image L2N / aggregation, a SIMT fp32 GEMM for the
// whitening projection (P is read once per 64-image tile instead of once per image) and the final L2N.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gdt {

struct GemScales {
    const float* ptr[GDT_MAX_SCALES];
    int hw[GDT_MAX_SCALES];
    int nscales;
};

__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float rem = x - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(rem));
}
// x^p for x >= eps > 0. mode: 3 -> cube, 2 -> square, 1 -> identity, 0 -> generic exp2(p*log2 x)
template <int MODE>
__device__ __forceinline__ float gem_pow(float x, float p) {
    if (MODE == 3) return x * x * x;
    if (MODE == 2) return x * x;
    if (MODE == 1) return x;
    // x >= eps > 0: lg2.approx (abs. error 2^-22 on the logarithm) and ex2.approx (2 ulp) give x^p to ~5e-7 relative
    return exp2f(p * __log2f(x));
}

template <int MODE>
__device__ __forceinline__ float gem_row_sum(const float* __restrict__ row, int hw, float eps, float p, int lane) {
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    // peel to 16-byte alignment
    int head = (int)(((16u - ((uintptr_t)row & 15u)) & 15u) >> 2);
    if (head > hw) head = hw;
    if (lane < head) acc0 += gem_pow<MODE>(fmaxf(row[lane], eps), p);
    const float4* body = (const float4*)(row + head);
    const int nvec = (hw - head) >> 2;
    // up to 8 independent 16-byte loads per lane in flight (a 3 KB ResNet row is covered by one batch)
    for (int base = 0; base < nvec; base += 256) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = base + j * 32 + lane;
            v[j] = i < nvec ? ld_stream_f4(body + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (base + j * 32 + lane < nvec) {
                acc0 += gem_pow<MODE>(fmaxf(v[j].x, eps), p);
                acc1 += gem_pow<MODE>(fmaxf(v[j].y, eps), p);
                acc2 += gem_pow<MODE>(fmaxf(v[j].z, eps), p);
                acc3 += gem_pow<MODE>(fmaxf(v[j].w, eps), p);
            }
        }
    }
    const int tail0 = head + (nvec << 2);
    if (tail0 + lane < hw) acc2 += gem_pow<MODE>(fmaxf(row[tail0 + lane], eps), p);
    float acc = (acc0 + acc1) + (acc2 + acc3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}

// one warp per row; rows are ordered [scale][image][channel]; g has the same order
// Rows of one scale with G lanes per row. Loads are predicated, the loop is warp-uniform (full-mask shuffles).
template <int MODE, int G>
__device__ __noinline__ void gem_pool_scale(const float* __restrict__ base, int hw, long long rows, float* __restrict__ g,
                                               float p, float eps, int root, float inv_p, long long wid, long long nwarps) {
    constexpr int RPW = 32 / G;                       // rows per warp
    const int lane = threadIdx.x & 31, lg = lane & (G - 1), sub = lane / G;
    for (long long rb = wid * RPW; rb < rows; rb += nwarps * RPW) {
        const long long r = rb + sub;
        const bool valid = r < rows;
        const float* row = base + (valid ? r : 0) * hw;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        int head = (int)(((16u - ((uintptr_t)row & 15u)) & 15u) >> 2);     // peel to 16-byte alignment
        if (head > hw) head = hw;
        if (valid && lg < head) acc0 += gem_pow<MODE>(fmaxf(row[lg], eps), p);
        const float4* body = (const float4*)(row + head);
        const int nvec = (hw - head) >> 2;
        for (int b0 = 0; b0 < nvec; b0 += G * 8) {
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = b0 + j * G + lg;
                v[j] = (valid && i < nvec) ? ld_stream_f4(body + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (valid && b0 + j * G + lg < nvec) {
                    acc0 += gem_pow<MODE>(fmaxf(v[j].x, eps), p);
                    acc1 += gem_pow<MODE>(fmaxf(v[j].y, eps), p);
                    acc2 += gem_pow<MODE>(fmaxf(v[j].z, eps), p);
                    acc3 += gem_pow<MODE>(fmaxf(v[j].w, eps), p);
                }
            }
        }
        const int tail0 = head + (nvec << 2);
        if (valid && tail0 + lg < hw) acc2 += gem_pow<MODE>(fmaxf(row[tail0 + lg], eps), p);
        float acc = (acc0 + acc1) + (acc2 + acc3);
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (valid && lg == 0) {
            const float mean = acc / (float)hw;
            g[r] = root ? powf(mean, inv_p) : mean;
        }
    }
}

// All rows of one launch for a fixed exponent mode. Kept out of line so that each mode gets its own register
// allocation instead of the union of all four.
template <int MODE>
__device__ __noinline__ void gem_pool_rows(const GemScales& S, long long rows_per_scale, long long total_rows, float p,
                                           float eps, int root, float* __restrict__ g) {
    const float inv_p = 1.0f / p;
    const long long nwarps = (long long)gridDim.x * 8;
    const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    // scale by scale: G lanes per row (32 for long rows, 16 / 8 for the short rows of the small scales, so a
    // warp keeps 2 / 4 rows in flight), every warp takes part in every scale (grid-stride)
    for (int s = 0; s < S.nscales; ++s) {
        const int hw = S.hw[s];
        float* gs = g + (long long)s * rows_per_scale;
        if (hw > 256) gem_pool_scale<MODE, 32>(S.ptr[s], hw, rows_per_scale, gs, p, eps, root, inv_p, wid, nwarps);
        else if (hw > 96) gem_pool_scale<MODE, 16>(S.ptr[s], hw, rows_per_scale, gs, p, eps, root, inv_p, wid, nwarps);
        else gem_pool_scale<MODE, 8>(S.ptr[s], hw, rows_per_scale, gs, p, eps, root, inv_p, wid, nwarps);
    }
}

// G lanes per row; rows are ordered [scale][image][channel]; g has the same order.
// root != 0: g = mean^(1/p) (GeM proper); root == 0: g = mean, the caller applies the root (gem_finalize_kernel).
__global__ void __launch_bounds__(256, 3)
gem_pool_kernel(const __grid_constant__ GemScales S, long long rows_per_scale, long long total_rows,
                const float* __restrict__ p_dev, float eps, int root, float* __restrict__ g) {
    const float p = __ldg(p_dev);
    if (p == 3.0f) gem_pool_rows<3>(S, rows_per_scale, total_rows, p, eps, root, g);
    else if (p == 2.0f) gem_pool_rows<2>(S, rows_per_scale, total_rows, p, eps, root, g);
    else if (p == 1.0f) gem_pool_rows<1>(S, rows_per_scale, total_rows, p, eps, root, g);
    else gem_pool_rows<0>(S, rows_per_scale, total_rows, p, eps, root, g);
}

static unsigned gem_pool_grid(long long total_rows) {
    const long long want = ceil_div_ll(total_rows, 8);
    const long long cap = (long long)sm_count_current_device() * 12;   // 3 resident CTAs/SM x 4 for load balance
    return (unsigned)(want < cap ? want : cap);
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    return t;
}

struct DescScales {
    const float* ptr[GDT_MAX_SCALES];   // per-scale [n][c] pooled vectors / descriptors
};

// one CTA per image: per-scale L2N(+eps) (skipped for already-normalised descriptors), optional multi-scale
// power mean + eps-free renorm, optional centring for the whitening projection.
// msp = p_dev[0] when GDT_GEM_MSP_IS_P, else msp_host.
__global__ void __launch_bounds__(256)
gem_finalize_kernel(DescScales D, int n, int c, int scales, const float* __restrict__ p_dev, float msp_host, int flags,
                    const float* __restrict__ m, float* __restrict__ out /* [n][c] */) {
    __shared__ float red[8];
    __shared__ float inv_norm[GDT_MAX_SCALES];
    const int img = blockIdx.x, tid = threadIdx.x;
    if (c <= 2048) {
        // register-resident path: every global load of a phase is issued before the first dependent operation
        const bool normalised_ = (flags & GDT_DESC_NORMALISED) != 0;
        const bool raw_mean_ = (flags & GDT_POOLED_RAW_MEAN) != 0;
        const bool aggregate_ = (flags & GDT_GEM_AGGREGATE) != 0;
        const float p_ = p_dev ? __ldg(p_dev) : msp_host;
        const float inv_p_ = raw_mean_ ? 1.0f / p_ : 1.0f;
        const float msp_ = (flags & GDT_GEM_MSP_IS_P) ? p_ : msp_host;
        const float inv_msp_ = (float)(1.0 / (double)msp_);
        float mv[8], acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = tid + j * 256;
            mv[j] = (m && i < c) ? __ldg(m + i) : 0.f;
            acc[j] = 0.f;
        }
        for (int s = 0; s < scales; ++s) {
            const float* gs = D.ptr[s] + (size_t)img * c;
            float r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { const int i = tid + j * 256; r[j] = i < c ? gs[i] : 0.f; }
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (tid + j * 256 < c) {
                    if (raw_mean_) r[j] = powf(r[j], inv_p_);
                    ss += r[j] * r[j];
                }
            }
            float den = 1.0f;
            if (!normalised_) den = sqrtf(block_sum_256(ss, red)) + 1e-6f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = normalised_ ? r[j] : r[j] / den;
                if (!aggregate_) acc[j] = d;
                else acc[j] += (msp_ == 1.0f) ? d : powf(d, msp_);
            }
        }
        float* o = out + (size_t)img * c;
        if (aggregate_) {
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float v = acc[j] / (float)scales;
                if (msp_ != 1.0f) v = powf(v, inv_msp_);
                if (tid + j * 256 >= c) v = 0.f;
                acc[j] = v;
                ss += v * v;
            }
            const float nrm = sqrtf(block_sum_256(ss, red));
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = acc[j] / nrm;
        }
        const bool split_ = (flags & GDT_SPLIT_OUT) != 0;     // tcgen05 whitening: write TF32 hi / lo halves
        float* o_lo = out + (size_t)n * c + (size_t)img * c;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = tid + j * 256;
            if (i < c) {
                const float val = m ? acc[j] - mv[j] : acc[j];
                if (split_) {
                    uint32_t hi, lo;
                    tf32_split(val, hi, lo);
                    o[i] = __uint_as_float(hi);
                    o_lo[i] = __uint_as_float(lo);
                } else {
                    o[i] = val;
                }
            }
        }
        return;
    }
    const bool normalised = (flags & GDT_DESC_NORMALISED) != 0;
    const bool raw_mean = (flags & GDT_POOLED_RAW_MEAN) != 0;   // pooled values still need the 1/p root
    const float inv_p = raw_mean ? 1.0f / __ldg(p_dev) : 1.0f;
    for (int s = 0; s < scales; ++s) {
        if (normalised) {
            if (tid == 0) inv_norm[s] = 1.0f;
            continue;
        }
        float* gs = const_cast<float*>(D.ptr[s]) + (size_t)img * c;
        float ss = 0.f;
        for (int i = tid; i < c; i += 256) {
            float v = gs[i];
            if (raw_mean) { v = powf(v, inv_p); gs[i] = v; }   // same thread re-reads these elements below
            ss += v * v;
        }
        ss = block_sum_256(ss, red);
        if (tid == 0) inv_norm[s] = sqrtf(ss) + 1e-6f;
    }
    __syncthreads();
    const bool aggregate = (flags & GDT_GEM_AGGREGATE) != 0;
    const float msp = (flags & GDT_GEM_MSP_IS_P) ? __ldg(p_dev) : msp_host;
    float* o = out + (size_t)img * c;
    if (!aggregate) {
        const float* gs = D.ptr[0] + (size_t)img * c;
        for (int i = tid; i < c; i += 256) {
            float v = normalised ? gs[i] : gs[i] / inv_norm[0];
            if (m) v -= m[i];
            o[i] = v;
        }
        return;
    }
    // v = sum_s d_s^msp ; v = (v / S)^(1/msp) ; v /= ||v||   (wrapper.py:238-243)
    const float inv_msp = (float)(1.0 / (double)msp);
    float ss = 0.f;
    for (int i = tid; i < c; i += 256) {
        float v = 0.f;
        for (int s = 0; s < scales; ++s) {
            const float g = D.ptr[s][(size_t)img * c + i];
            const float d = normalised ? g : g / inv_norm[s];
            v += (msp == 1.0f) ? d : powf(d, msp);
        }
        v = v / (float)scales;
        if (msp != 1.0f) v = powf(v, inv_msp);
        o[i] = v;
        ss += v * v;
    }
    ss = block_sum_256(ss, red);
    const float nrm = sqrtf(ss);
    for (int i = tid; i < c; i += 256) {
        float v = o[i] / nrm;
        if (m) v -= m[i];
        o[i] = v;
    }
}

// Xpart[z][n][dim] = V[n][kz] . P[dim][kz]^T over the K-slice kz = [z*klen, (z+1)*klen)
// (fp32 SIMT, 64x64 tile, BK = 16, 4x4 outputs per thread). Split-K fills the machine when n is small: the
// slices are summed in a fixed order by whiten_reduce_l2n_kernel, so results are run-to-run deterministic.
// ---- whitening projection: 3xTF32 on the tensor cores --------------------------------------------
// Xpart[z][n][dim] = V[n][kz] . P[dim][kz]^T over the K-slice kz = [z*klen, (z+1)*klen). fp32 operands are split in
// registers into a TF32 "hi" part and a TF32 "lo" remainder (x = hi + lo up to 2^-22 |x|) and every product is
// formed as lo*hi + hi*lo + hi*hi with fp32 accumulation (mma.sync.m16n8k8.tf32): fp32-level accuracy (~1e-6
// relative) at a third of the TF32 rate, ~4x fewer instructions than the SIMT fp32 loop this replaces. The GEMM is
// tiny (n x dim x c = 128 x 2048 x 2048); split-K fills the machine and the slices are summed in a fixed order by
// whiten_reduce_l2n_kernel, so results are run-to-run deterministic.
constexpr int kWBK = 32;    // K-step
constexpr int kWLd = kWBK + 4;   // smem row stride (floats): fragment loads hit 32 distinct banks

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(128)
whiten_gemm_kernel(const float* __restrict__ V, const float* __restrict__ P, int ldP, int n, int c, int dim, int klen,
                   float* __restrict__ Xpart) {
    __shared__ __align__(16) float As[2][64][kWLd];   // [m][k]
    __shared__ __align__(16) float Bs[2][64][kWLd];   // [n][k]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = blockIdx.y * 64, col0 = blockIdx.x * 64;
    const int kbeg = blockIdx.z * klen, kend = min(c, kbeg + klen);
    float* X = Xpart + (size_t)blockIdx.z * n * dim;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;     // 2 x 2 warps, 32 x 32 outputs each
    const int g = lane >> 2, t = lane & 3;
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

    // global -> registers -> smem staging: 64 rows x 32 k per operand, 4 float4 per thread per operand
    const bool vec = (c & 3) == 0 && (ldP & 3) == 0 && ((((uintptr_t)V) | ((uintptr_t)P)) & 15) == 0;
    const int lr = tid >> 3, lk = (tid & 7) << 2;              // rows lr + 16*h, k offset lk
    float4 ra[4], rb[4];
    auto fetch = [&](int k0) {
        const int k = k0 + lk;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int ar = row0 + lr + 16 * h, br = col0 + lr + 16 * h;
            if (vec && k + 3 < kend) {
                ra[h] = ar < n ? __ldg((const float4*)(V + (size_t)ar * c + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
                rb[h] = br < dim ? __ldg((const float4*)(P + (size_t)br * ldP + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                float a[4], b[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    a[q] = (ar < n && k + q < kend) ? V[(size_t)ar * c + k + q] : 0.f;
                    b[q] = (br < dim && k + q < kend) ? P[(size_t)br * ldP + k + q] : 0.f;
                }
                ra[h] = make_float4(a[0], a[1], a[2], a[3]);
                rb[h] = make_float4(b[0], b[1], b[2], b[3]);
            }
        }
    };
    fetch(kbeg);
    int buf = 0;
    for (int k0 = kbeg; k0 < kend; k0 += kWBK, buf ^= 1) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            *(float4*)&As[buf][lr + 16 * h][lk] = ra[h];
            *(float4*)&Bs[buf][lr + 16 * h][lk] = rb[h];
        }
        __syncthreads();
        if (k0 + kWBK < kend) fetch(k0 + kWBK);
#pragma unroll
        for (int k8 = 0; k8 < kWBK; k8 += 8) {
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int m = wm + i * 16;
                tf32_split(As[buf][m + g][k8 + t], ah[i][0], al[i][0]);
                tf32_split(As[buf][m + g + 8][k8 + t], ah[i][1], al[i][1]);
                tf32_split(As[buf][m + g][k8 + t + 4], ah[i][2], al[i][2]);
                tf32_split(As[buf][m + g + 8][k8 + t + 4], ah[i][3], al[i][3]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t bh[2], bl[2];
                tf32_split(Bs[buf][wn + j * 8 + g][k8 + t], bh[0], bl[0]);
                tf32_split(Bs[buf][wn + j * 8 + g][k8 + t + 4], bh[1], bl[1]);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    mma_tf32(acc[i][j], al[i], bh);     // small terms first
                    mma_tf32(acc[i][j], ah[i], bl);
                    mma_tf32(acc[i][j], ah[i], bh);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = row0 + wm + i * 16 + g, cc = col0 + wn + j * 8 + 2 * t;
            if (r < n) {
                if (cc < dim) X[(size_t)r * dim + cc] = acc[i][j][0];
                if (cc + 1 < dim) X[(size_t)r * dim + cc + 1] = acc[i][j][1];
            }
            if (r + 8 < n) {
                if (cc < dim) X[(size_t)(r + 8) * dim + cc] = acc[i][j][2];
                if (cc + 1 < dim) X[(size_t)(r + 8) * dim + cc + 1] = acc[i][j][3];
            }
        }
}

// ---- whitening projection on tcgen05 (3xTF32) ----------------------------------------------------
// Same contraction as whiten_gemm_kernel, for callers that prepared the projection once with gdt_whiten_prepare
// (P split into TF32-exact hi / lo halves, [2][dim][c]). gem_finalize_kernel writes the centred descriptors already
// split ([2][n][c]), so all four operands are plain TMA tiles and the kernel is a TMA -> tcgen05.mma(kind::tf32) ->
// TMEM pipeline: per 32-wide K block three accumulating products lo*hi + hi*lo + hi*hi (12 MMAs of M=128, N=128, K=8).
// One CTA per (128-row image tile, 128-column output tile, K slice); warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 2..5 = epilogue (TMEM -> partial sums in global memory).
constexpr int kTcBM = 128, kTcBN = 128, kTcBK = 32;           // fp32 elements; 32 * 4 B = one 128-byte swizzle span
constexpr int kTcStages = 3;
constexpr int kTcTile = kTcBM * kTcBK * 4;                      // 16 KB per operand tile
constexpr int kTcStageBytes = 4 * kTcTile;                      // A_hi, A_lo, B_hi, B_lo
constexpr int kTcSmem = kTcStages * kTcStageBytes + 1024 + 256;
// kind::tf32 instruction descriptor: D = f32 (1 << 4), A = B = tf32 (format 2 at [7,10) and [10,13)), K-major both
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcBN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(192, 1)
whiten_tc_kernel(const __grid_constant__ CUtensorMap map_vh, const __grid_constant__ CUtensorMap map_vl,
                 const __grid_constant__ CUtensorMap map_ph, const __grid_constant__ CUtensorMap map_pl, int n, int dim,
                 int klen, int c, float* __restrict__ Xpart) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    uint64_t* bars = (uint64_t*)(smem_gen + kTcStages * kTcStageBytes);
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kTcStages, bar_done = bar_empty + 8 * kTcStages;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kTcStages + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col0 = blockIdx.x * kTcBN, row0 = blockIdx.y * kTcBM;
    const int kbeg = blockIdx.z * klen, kend = min(c, kbeg + klen);
    const int nkb = (kend - kbeg + kTcBK - 1) / kTcBK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_vh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_vl) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_ph) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_pl) : "memory");
        for (int s = 0; s < kTcStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), kTcBN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                const uint32_t st = smem_base + stage * kTcStageBytes;
                const int k = kbeg + kb * kTcBK;
                mbar_arrive_expect_tx(bar_full + 8 * stage, kTcStageBytes);
                tma_load_2d(st, &map_vh, bar_full + 8 * stage, k, row0);
                tma_load_2d(st + kTcTile, &map_vl, bar_full + 8 * stage, k, row0);
                tma_load_2d(st + 2 * kTcTile, &map_ph, bar_full + 8 * stage, k, col0);
                tma_load_2d(st + 3 * kTcTile, &map_pl, bar_full + 8 * stage, k, col0);
                if (++stage == kTcStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                const uint32_t st = smem_base + stage * kTcStageBytes;
                const uint64_t ah = umma_smem_desc(st), al = umma_smem_desc(st + kTcTile);
                const uint64_t bh = umma_smem_desc(st + 2 * kTcTile), bl = umma_smem_desc(st + 3 * kTcTile);
#pragma unroll
                for (int kk = 0; kk < kTcBK / 8; ++kk) {
                    const uint64_t o = (uint64_t)(kk * 2);      // +32 bytes along K inside the swizzle span
                    umma_tf32(tmem_base, al + o, bh + o, kTcIdesc, (kb | kk) != 0 ? 1u : 0u);   // small terms first
                    umma_tf32(tmem_base, ah + o, bl + o, kTcIdesc, 1u);
                    umma_tf32(tmem_base, ah + o, bh + o, kTcIdesc, 1u);
                }
                umma_commit(bar_empty + 8 * stage);
                if (++stage == kTcStages) { stage = 0; phase ^= 1; }
            }
            umma_commit(bar_done);
        }
    } else {
        // epilogue: thread == output row (TMEM lane)
        const int quarter = warp & 3;
        const int r = row0 + quarter * 32 + lane;
        mbar_wait(bar_done, 0);
        tc_fence_after();
        float* xrow = Xpart + ((size_t)blockIdx.z * n + (r < n ? r : 0)) * dim + col0;
#pragma unroll 1
        for (int cc = 0; cc < kTcBN / 32; ++cc) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + cc * 32, v);
            tmem_ld_wait();
            if (r < n) {
                if (col0 + cc * 32 + 32 <= dim && (dim & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *(float4*)(xrow + cc * 32 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                     __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + cc * 32 + j < dim) xrow[cc * 32 + j] = __uint_as_float(v[j]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTcBN);
    }
}

// split every element into TF32-exact hi and lo parts: out[0][i] = tf32(x), out[1][i] = tf32(x - tf32(x))
__global__ void __launch_bounds__(256)
tf32_split_kernel(const float* __restrict__ src, int rows, int cols, int ld, float* __restrict__ out) {
    const size_t total = (size_t)rows * cols;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const size_t r = i / cols, k = i - r * cols;
        uint32_t hi, lo;
        tf32_split(src[r * ld + k], hi, lo);
        out[i] = __uint_as_float(hi);
        out[total + i] = __uint_as_float(lo);
    }
}

// desc[r] = x / (||x||_2 + eps) with x = sum_z Xpart[z][r] (fixed order); one CTA per row, the row kept in registers
constexpr int kReduceMaxPerThread = 8;   // rows up to 256 * 4 * 8 = 8192 wide stay in registers
__global__ void __launch_bounds__(256)
whiten_reduce_l2n_kernel(const float* __restrict__ Xpart, int n, int dim, int splitk, float eps, float* __restrict__ desc) {
    __shared__ float red[8];
    const int r = blockIdx.x;
    const bool vec = (dim & 3) == 0 && ((((uintptr_t)Xpart) | ((uintptr_t)desc)) & 15) == 0 && dim <= 256 * 4 * kReduceMaxPerThread;
    float ss = 0.f;
    if (vec) {
        float4 acc[kReduceMaxPerThread];
        const int nv = dim >> 2;
#pragma unroll
        for (int j = 0; j < kReduceMaxPerThread; ++j) {
            acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int i = threadIdx.x + j * 256;
            if (i < nv) {
                for (int z = 0; z < splitk; ++z) {
                    const float4 v = __ldg((const float4*)(Xpart + ((size_t)z * n + r) * dim) + i);
                    acc[j].x += v.x; acc[j].y += v.y; acc[j].z += v.z; acc[j].w += v.w;
                }
                ss += acc[j].x * acc[j].x + acc[j].y * acc[j].y + acc[j].z * acc[j].z + acc[j].w * acc[j].w;
            }
        }
        ss = block_sum_256(ss, red);
        const float den = sqrtf(ss) + eps;
#pragma unroll
        for (int j = 0; j < kReduceMaxPerThread; ++j) {
            const int i = threadIdx.x + j * 256;
            if (i < nv)
                *((float4*)(desc + (size_t)r * dim) + i) = make_float4(acc[j].x / den, acc[j].y / den, acc[j].z / den, acc[j].w / den);
        }
        return;
    }
    for (int i = threadIdx.x; i < dim; i += 256) {
        float v = 0.f;
        for (int z = 0; z < splitk; ++z) v += Xpart[((size_t)z * n + r) * dim + i];
        desc[(size_t)r * dim + i] = v;
        ss += v * v;
    }
    ss = block_sum_256(ss, red);
    const float den = sqrtf(ss) + eps;
    for (int i = threadIdx.x; i < dim; i += 256) desc[(size_t)r * dim + i] = desc[(size_t)r * dim + i] / den;
}

static int whiten_splitk(int n, int c, int dim) {
    const int tiles = ceil_div(dim, 64) * ceil_div(n, 64);
    int z = ceil_div(4 * sm_count_current_device(), tiles);
    const int zmax = c / 128 > 1 ? c / 128 : 1;          // at least 128 of K per slice
    if (z > zmax) z = zmax;
    if (z > 32) z = 32;
    return z < 1 ? 1 : z;
}

// out = in / (||row||_2 + eps), one CTA per row (in == out allowed)
__global__ void __launch_bounds__(256) l2n_rows_kernel(const float* X, float* Y, int dim, float eps) {
    __shared__ float red[8];
    const float* x = X + (size_t)blockIdx.x * dim;
    float* y = Y + (size_t)blockIdx.x * dim;
    float ss = 0.f;
    for (int i = threadIdx.x; i < dim; i += 256) { const float v = x[i]; ss += v * v; }
    ss = block_sum_256(ss, red);
    const float den = sqrtf(ss) + eps;
    for (int i = threadIdx.x; i < dim; i += 256) y[i] = x[i] / den;
}

static int whiten_tc_slices(int n, int c, int dim) {
    const int tiles = ceil_div(dim, kTcBN) * ceil_div(n, kTcBM);
    int z = sm_count_current_device() / tiles;              // one CTA per SM (198 KB of smem each): stay within one wave
    const int zmax = c / 64 > 1 ? c / 64 : 1;              // at least two K blocks per slice
    if (z > zmax) z = zmax;
    if (z > 32) z = 32;
    return z < 1 ? 1 : z;
}
static int whiten_ws_slices(int n, int c, int dim) {
    const int a = whiten_splitk(n, c, dim), b = whiten_tc_slices(n, c, dim);
    return a > b ? a : b;
}

// finalize (+ whitening projection + final L2N) shared by gdt_gem_whiten and gdt_desc_post.
// V holds 2 * n * c floats (the second half is used by the tcgen05 path for the lo parts).
static int desc_tail(const DescScales& D, int n, int c, int scales, const float* p_dev, float msp_host, int flags,
                     const float* P, int ldP, const float* P_split, const float* m, int dim, float* desc, float* V,
                     float* Xpart, cudaStream_t stream) {
    const bool tc = P && P_split && (c & 3) == 0 && c <= 2048 && ((((uintptr_t)P_split) | ((uintptr_t)V)) & 15) == 0;
    float* fin_out = P ? V : desc;
    gem_finalize_kernel<<<n, 256, 0, stream>>>(D, n, c, scales, p_dev, msp_host, flags | (tc ? GDT_SPLIT_OUT : 0),
                                               P ? m : nullptr, fin_out);
    GDT_LAUNCH_CHECK();
    if (!P) return GDT_OK;
    int slices;
    if (tc) {
        CUtensorMap mvh, mvl, mph, mpl;
        int rc = make_tile_map(&mvh, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, V, n, c, kTcBM);
        if (rc == GDT_OK) rc = make_tile_map(&mvl, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, V + (size_t)n * c, n, c, kTcBM);
        if (rc == GDT_OK) rc = make_tile_map(&mph, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, P_split, dim, c, kTcBN);
        if (rc == GDT_OK) rc = make_tile_map(&mpl, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, P_split + (size_t)dim * c, dim, c, kTcBN);
        if (rc != GDT_OK) return rc;
        static bool attr_set[32] = {false};
        const int slot = current_device_slot();
        if (!attr_set[slot]) {
            GDT_CUDA(cudaFuncSetAttribute(whiten_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
            attr_set[slot] = true;
        }
        int z = whiten_tc_slices(n, c, dim);
        int klen = ceil_div(ceil_div(c, z), kTcBK) * kTcBK;
        while (z > 1 && ceil_div(c, klen) * ceil_div(dim, kTcBN) * ceil_div(n, kTcBM) > sm_count_current_device()) {
            --z;                                            // rounding klen up to whole K blocks must not add a wave
            klen = ceil_div(ceil_div(c, z), kTcBK) * kTcBK;
        }
        dim3 grid(ceil_div(dim, kTcBN), ceil_div(n, kTcBM), ceil_div(c, klen));
        whiten_tc_kernel<<<grid, 192, kTcSmem, stream>>>(mvh, mvl, mph, mpl, n, dim, klen, c, Xpart);
        GDT_LAUNCH_CHECK();
        slices = (int)grid.z;
    } else {
        const int splitk = whiten_splitk(n, c, dim);
        const int klen = ceil_div(ceil_div(c, splitk), kWBK) * kWBK;
        dim3 grid(ceil_div(dim, 64), ceil_div(n, 64), ceil_div(c, klen));
        whiten_gemm_kernel<<<grid, 128, 0, stream>>>(V, P, ldP, n, c, dim, klen, Xpart);
        GDT_LAUNCH_CHECK();
        slices = (int)grid.z;
    }
    whiten_reduce_l2n_kernel<<<n, 256, 0, stream>>>(Xpart, n, dim, slices, 1e-6f, desc);
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}

}  // namespace gdt

using namespace gdt;

extern "C" size_t gdt_gem_whiten_workspace_bytes(int n, int c, int scales, int dim) {
    if (n <= 0 || c <= 0 || scales <= 0) return 0;
    return align_up((size_t)scales * n * c * sizeof(float), 256) + align_up((size_t)2 * n * c * sizeof(float), 256) +
           align_up((size_t)whiten_ws_slices(n, c, dim > 0 ? dim : c) * n * (dim > 0 ? dim : c) * sizeof(float), 256) + 256;
}

extern "C" int gdt_gem_whiten(const float* const* host_fmaps, const int* host_h, const int* host_w, int n, int c,
                              int scales, const float* p_dev, float eps, int flags, const float* P, int ldP,
                              const float* P_split, const float* m, int dim, float* desc, void* ws, size_t ws_bytes,
                              void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!host_fmaps || !host_h || !host_w || !p_dev || !desc || !ws) return GDT_ERR_INVALID_ARGUMENT;
    if (n <= 0 || c <= 0 || scales <= 0 || scales > GDT_MAX_SCALES) return GDT_ERR_INVALID_ARGUMENT;
    if (!(flags & GDT_GEM_AGGREGATE) && scales != 1) return GDT_ERR_INVALID_ARGUMENT;
    if (P && (!m || dim <= 0 || ldP < c)) return GDT_ERR_INVALID_ARGUMENT;
    if (!P && dim != c) return GDT_ERR_INVALID_ARGUMENT;
    if (ws_bytes < gdt_gem_whiten_workspace_bytes(n, c, scales, dim)) return GDT_ERR_WORKSPACE_TOO_SMALL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return GDT_ERR_NO_DEVICE; }

    Workspace W(ws, ws_bytes);
    float* g = W.take<float>((size_t)scales * n * c);
    float* V = W.take<float>((size_t)2 * n * c);
    float* Xpart = W.take<float>((size_t)whiten_ws_slices(n, c, dim) * n * dim);
    if (!W.ok()) return GDT_ERR_WORKSPACE_TOO_SMALL;

    GemScales S;
    S.nscales = scales;
    for (int s = 0; s < GDT_MAX_SCALES; ++s) { S.ptr[s] = nullptr; S.hw[s] = 0; }
    for (int s = 0; s < scales; ++s) {
        if (!host_fmaps[s] || host_h[s] <= 0 || host_w[s] <= 0) return GDT_ERR_INVALID_ARGUMENT;
        if (((uintptr_t)host_fmaps[s]) & 3) return GDT_ERR_INVALID_ARGUMENT;
        S.ptr[s] = host_fmaps[s];
        S.hw[s] = host_h[s] * host_w[s];
    }
    const long long rows_per_scale = (long long)n * c;
    const long long total_rows = rows_per_scale * scales;
    gem_pool_kernel<<<gem_pool_grid(total_rows), 256, 0, stream>>>(S, rows_per_scale, total_rows, p_dev, eps, 0, g);
    GDT_LAUNCH_CHECK();
    DescScales D;
    for (int s = 0; s < GDT_MAX_SCALES; ++s) D.ptr[s] = s < scales ? g + (size_t)s * n * c : nullptr;
    return desc_tail(D, n, c, scales, p_dev, 1.0f, (flags & (GDT_GEM_AGGREGATE | GDT_GEM_MSP_IS_P)) | GDT_POOLED_RAW_MEAN, P, ldP, P_split, m, dim, desc, V,
                     Xpart, stream);
}

extern "C" int gdt_gem_pool(const float* fmap, int n, int c, int h, int w, const float* p_dev, float eps, float* pooled,
                            void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!fmap || !p_dev || !pooled || n <= 0 || c <= 0 || h <= 0 || w <= 0) return GDT_ERR_INVALID_ARGUMENT;
    if (((uintptr_t)fmap) & 3) return GDT_ERR_INVALID_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return GDT_ERR_NO_DEVICE; }
    GemScales S;
    S.nscales = 1;
    for (int s = 0; s < GDT_MAX_SCALES; ++s) { S.ptr[s] = nullptr; S.hw[s] = 0; }
    S.ptr[0] = fmap;
    S.hw[0] = h * w;
    const long long rows = (long long)n * c;
    gem_pool_kernel<<<gem_pool_grid(rows), 256, 0, stream>>>(S, rows, rows, p_dev, eps, 1, pooled);
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}

extern "C" int gdt_l2n_rows(const float* x, int n, int dim, float eps, float* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!x || !out || n <= 0 || dim <= 0) return GDT_ERR_INVALID_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return GDT_ERR_NO_DEVICE; }
    l2n_rows_kernel<<<n, 256, 0, stream>>>(x, out, dim, eps);
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}

extern "C" size_t gdt_desc_post_workspace_bytes(int n, int c, int dim) {
    if (n <= 0 || c <= 0) return 0;
    return align_up((size_t)2 * n * c * sizeof(float), 256) + align_up((size_t)whiten_ws_slices(n, c, dim > 0 ? dim : c) * n * (dim > 0 ? dim : c) * sizeof(float), 256) + 256;
}

extern "C" int gdt_desc_post(const float* const* host_descs, int n, int c, int scales, const float* msp_dev, float msp_host,
                             int flags, const float* P, int ldP, const float* P_split, const float* m, int dim, float* out,
                             void* ws, size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!host_descs || !out || !ws || n <= 0 || c <= 0 || scales <= 0 || scales > GDT_MAX_SCALES)
        return GDT_ERR_INVALID_ARGUMENT;
    if (!(flags & GDT_GEM_AGGREGATE) && scales != 1) return GDT_ERR_INVALID_ARGUMENT;
    if ((flags & GDT_GEM_MSP_IS_P) && !msp_dev) return GDT_ERR_INVALID_ARGUMENT;
    if (P && (!m || dim <= 0 || ldP < c)) return GDT_ERR_INVALID_ARGUMENT;
    if (!P && dim != c) return GDT_ERR_INVALID_ARGUMENT;
    if (ws_bytes < gdt_desc_post_workspace_bytes(n, c, dim)) return GDT_ERR_WORKSPACE_TOO_SMALL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return GDT_ERR_NO_DEVICE; }
    Workspace W(ws, ws_bytes);
    float* V = W.take<float>((size_t)2 * n * c);
    float* Xpart = W.take<float>((size_t)whiten_ws_slices(n, c, dim) * n * dim);
    if (!W.ok()) return GDT_ERR_WORKSPACE_TOO_SMALL;
    DescScales D;
    for (int s = 0; s < GDT_MAX_SCALES; ++s) {
        D.ptr[s] = s < scales ? host_descs[s] : nullptr;
        if (s < scales && !host_descs[s]) return GDT_ERR_INVALID_ARGUMENT;
    }
    return desc_tail(D, n, c, scales, msp_dev, msp_host,
                     (flags & (GDT_GEM_AGGREGATE | GDT_GEM_MSP_IS_P)) | GDT_DESC_NORMALISED, P, ldP, P_split, m, dim, out, V, Xpart,
                     stream);
}

extern "C" int gdt_whiten_prepare(const float* P, int ldP, int c, int dim, float* P_split, void* stream_) {
    if (!P || !P_split || c <= 0 || dim <= 0 || ldP < c) return GDT_ERR_INVALID_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return GDT_ERR_NO_DEVICE; }
    const size_t total = (size_t)dim * c;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)sm_count_current_device() * 16;
    if (blocks > cap) blocks = cap;
    tf32_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(P, dim, c, ldP, P_split);
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}
