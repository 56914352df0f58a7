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
#include "common.cuh"

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

// x^p for x >= eps > 0. mode: 3 -> cube, 2 -> square, 1 -> identity, 0 -> generic exp2(p*log2 x)
template <int MODE>
__device__ __forceinline__ float gem_pow(float x, float p) {
    if (MODE == 3) return x * x * x;
    if (MODE == 2) return x * x;
    if (MODE == 1) return x;
    return exp2f(p * log2f(x));
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
    int i = lane;
    for (; i + 96 < nvec; i += 128) {
        const float4 a = ld_stream_f4(body + i), b = ld_stream_f4(body + i + 32), c = ld_stream_f4(body + i + 64),
                     d = ld_stream_f4(body + i + 96);
        acc0 += gem_pow<MODE>(fmaxf(a.x, eps), p) + gem_pow<MODE>(fmaxf(a.y, eps), p);
        acc1 += gem_pow<MODE>(fmaxf(a.z, eps), p) + gem_pow<MODE>(fmaxf(a.w, eps), p);
        acc2 += gem_pow<MODE>(fmaxf(b.x, eps), p) + gem_pow<MODE>(fmaxf(b.y, eps), p);
        acc3 += gem_pow<MODE>(fmaxf(b.z, eps), p) + gem_pow<MODE>(fmaxf(b.w, eps), p);
        acc0 += gem_pow<MODE>(fmaxf(c.x, eps), p) + gem_pow<MODE>(fmaxf(c.y, eps), p);
        acc1 += gem_pow<MODE>(fmaxf(c.z, eps), p) + gem_pow<MODE>(fmaxf(c.w, eps), p);
        acc2 += gem_pow<MODE>(fmaxf(d.x, eps), p) + gem_pow<MODE>(fmaxf(d.y, eps), p);
        acc3 += gem_pow<MODE>(fmaxf(d.z, eps), p) + gem_pow<MODE>(fmaxf(d.w, eps), p);
    }
    for (; i < nvec; i += 32) {
        const float4 a = ld_stream_f4(body + i);
        acc0 += gem_pow<MODE>(fmaxf(a.x, eps), p) + gem_pow<MODE>(fmaxf(a.y, eps), p);
        acc1 += gem_pow<MODE>(fmaxf(a.z, eps), p) + gem_pow<MODE>(fmaxf(a.w, eps), p);
    }
    const int tail0 = head + (nvec << 2);
    if (tail0 + lane < hw) acc2 += gem_pow<MODE>(fmaxf(row[tail0 + lane], eps), p);
    float acc = (acc0 + acc1) + (acc2 + acc3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}

// one warp per row; rows are ordered [scale][image][channel]; g has the same order
__global__ void __launch_bounds__(256)
gem_pool_kernel(GemScales S, long long rows_per_scale, long long total_rows, const float* __restrict__ p_dev, float eps,
                float* __restrict__ g) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (warp >= total_rows) return;
    const int s = (int)(warp / rows_per_scale);
    const long long r = warp - (long long)s * rows_per_scale;
    const int hw = S.hw[s];
    const float* row = S.ptr[s] + r * hw;
    const float p = __ldg(p_dev);
    float sum;
    if (p == 3.0f) sum = gem_row_sum<3>(row, hw, eps, p, lane);
    else if (p == 2.0f) sum = gem_row_sum<2>(row, hw, eps, p, lane);
    else if (p == 1.0f) sum = gem_row_sum<1>(row, hw, eps, p, lane);
    else sum = gem_row_sum<0>(row, hw, eps, p, lane);
    if (lane == 0) {
        const float mean = sum / (float)hw;
        g[warp] = powf(mean, 1.0f / p);
    }
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

// one CTA per image: per-scale L2N(+eps), optional multi-scale power mean + eps-free renorm,
// optional centring for the whitening projection.
__global__ void __launch_bounds__(256)
gem_finalize_kernel(const float* __restrict__ g, int n, int c, int scales, const float* __restrict__ p_dev, int flags,
                    const float* __restrict__ m, float* __restrict__ out /* [n][c] */) {
    __shared__ float red[8];
    __shared__ float inv_norm[GDT_MAX_SCALES];
    const int img = blockIdx.x, tid = threadIdx.x;
    for (int s = 0; s < scales; ++s) {
        const float* gs = g + ((size_t)s * n + img) * c;
        float ss = 0.f;
        for (int i = tid; i < c; i += 256) { const float v = gs[i]; ss += v * v; }
        ss = block_sum_256(ss, red);
        if (tid == 0) inv_norm[s] = sqrtf(ss) + 1e-6f;
    }
    __syncthreads();
    const bool aggregate = (flags & GDT_GEM_AGGREGATE) != 0;
    const float msp = (flags & GDT_GEM_MSP_IS_P) ? __ldg(p_dev) : 1.0f;
    float* o = out + (size_t)img * c;
    if (!aggregate) {
        const float* gs = g + (size_t)img * c;
        for (int i = tid; i < c; i += 256) {
            float v = gs[i] / inv_norm[0];
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
            const float d = g[((size_t)s * n + img) * c + i] / inv_norm[s];
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

// X[n][dim] = V[n][c] . P[dim][c]^T   (fp32 SIMT, 64x64 tile, BK = 16, 4x4 outputs per thread)
__global__ void __launch_bounds__(256)
whiten_gemm_kernel(const float* __restrict__ V, const float* __restrict__ P, int ldP, int n, int c, int dim,
                   float* __restrict__ X) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tid = threadIdx.x;
    const int row0 = blockIdx.y * 64, col0 = blockIdx.x * 64;
    const int tr = (tid >> 4) << 2, tc = (tid & 15) << 2;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lr = tid >> 2, lk = (tid & 3) << 2;  // each thread loads 4 consecutive k of one row
    for (int k0 = 0; k0 < c; k0 += 16) {
        float a[4], b[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + lk + q;
            const int ar = row0 + lr, br = col0 + lr;
            a[q] = (ar < n && k < c) ? V[(size_t)ar * c + k] : 0.f;
            b[q] = (br < dim && k < c) ? P[(size_t)br * ldP + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) { As[lk + q][lr] = a[q]; Bs[lk + q][lr] = b[q]; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 av = *(const float4*)&As[k][tr];
            const float4 bv = *(const float4*)&Bs[k][tc];
            const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = row0 + tr + i, cc = col0 + tc + j;
            if (r < n && cc < dim) X[(size_t)r * dim + cc] = acc[i][j];
        }
}

// rows /= (||row||_2 + eps)
__global__ void __launch_bounds__(256) l2n_rows_kernel(float* __restrict__ X, int dim, float eps) {
    __shared__ float red[8];
    float* x = X + (size_t)blockIdx.x * dim;
    float ss = 0.f;
    for (int i = threadIdx.x; i < dim; i += 256) { const float v = x[i]; ss += v * v; }
    ss = block_sum_256(ss, red);
    const float den = sqrtf(ss) + eps;
    for (int i = threadIdx.x; i < dim; i += 256) x[i] = x[i] / den;
}

}  // namespace gdt

using namespace gdt;

extern "C" size_t gdt_gem_whiten_workspace_bytes(int n, int c, int scales, int dim) {
    if (n <= 0 || c <= 0 || scales <= 0) return 0;
    (void)dim;
    return align_up((size_t)scales * n * c * sizeof(float), 256) + align_up((size_t)n * c * sizeof(float), 256) + 256;
}

extern "C" int gdt_gem_whiten(const float* const* host_fmaps, const int* host_h, const int* host_w, int n, int c,
                              int scales, const float* p_dev, float eps, int flags, const float* P, int ldP,
                              const float* m, int dim, float* desc, void* ws, size_t ws_bytes, void* stream_) {
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
    float* V = W.take<float>((size_t)n * c);
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
    gem_pool_kernel<<<(unsigned)ceil_div_ll(total_rows, 8), 256, 0, stream>>>(S, rows_per_scale, total_rows, p_dev, eps, g);
    GDT_LAUNCH_CHECK();
    float* fin_out = P ? V : desc;
    gem_finalize_kernel<<<n, 256, 0, stream>>>(g, n, c, scales, p_dev, flags, P ? m : nullptr, fin_out);
    GDT_LAUNCH_CHECK();
    if (P) {
        dim3 grid(ceil_div(dim, 64), ceil_div(n, 64));
        whiten_gemm_kernel<<<grid, 256, 0, stream>>>(V, P, ldP, n, c, dim, desc);
        GDT_LAUNCH_CHECK();
        l2n_rows_kernel<<<n, 256, 0, stream>>>(desc, dim, 1e-6f);
        GDT_LAUNCH_CHECK();
    }
    return GDT_OK;
}
