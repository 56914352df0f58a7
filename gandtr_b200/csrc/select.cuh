// Device helpers shared by the scoring kernels: total order on (score, index), exact dot products,
// block-wide bitonic sort.
//
// Ranking rule of the whole library (replaces the unstable np.argsort(-scores) of
// mdir/components/optim/score/cirscore.py:72): score descending, ties -> lower index first.
// It is encoded as ONE unsigned 64-bit key so that "sorts before" == "key is larger":
//   high 32 bits: order-preserving map of the fp32 score, low 32 bits: ~index.
#pragma once
#include <stdint.h>

namespace gdt {

__device__ __forceinline__ uint32_t ordered_bits(float s) {
    s += 0.0f;  // -0 -> +0
    const uint32_t u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t rank_key_bits(uint32_t obits, uint32_t idx) {
    return ((uint64_t)obits << 32) | (uint64_t)(~idx);
}
__device__ __forceinline__ uint64_t rank_key(float score, uint32_t idx) { return rank_key_bits(ordered_bits(score), idx); }
__device__ __forceinline__ float key_score(uint64_t k) { return from_ordered_bits((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_index(uint64_t k) { return ~(uint32_t)k; }

// The library's definition of an exact fp32 score: the dot product accumulated in fp64 (products of
// two fp32 values are exact in fp64) and rounded once to fp32. One warp per dot, fixed summation
// order (lane-strided 16-byte groups, then an xor-shuffle tree) so every kernel that scores the same
// (query, row) pair returns the same bits. `q` may live in shared memory.
__device__ __forceinline__ bool dot_vec_ok(const float* q, const float* x, int d) {
    return (d & 3) == 0 && ((((uintptr_t)x) | ((uintptr_t)q)) & 15) == 0;
}
__device__ __forceinline__ double warp_reduce_f64(double acc) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}
__device__ __forceinline__ float warp_exact_dot(const float* __restrict__ q, const float* __restrict__ x, int d, int lane) {
    double acc = 0.0;
    if (dot_vec_ok(q, x, d)) {
        const float4* x4 = (const float4*)x;
        const float4* q4 = (const float4*)q;
        for (int i = lane; i < (d >> 2); i += 32) {
            const float4 a = q4[i];
            const float4 b = __ldg(x4 + i);
            acc += (double)a.x * (double)b.x;
            acc += (double)a.y * (double)b.y;
            acc += (double)a.z * (double)b.z;
            acc += (double)a.w * (double)b.w;
        }
    } else {
        for (int i = lane; i < d; i += 32) acc += (double)q[i] * (double)__ldg(x + i);
    }
    return (float)warp_reduce_f64(acc);
}

// NQ queries (rows of `q`, row stride qs floats, shared memory) against one database row: the row is
// read once. Same summation order per query as warp_exact_dot.
template <int NQ>
__device__ __forceinline__ void warp_exact_dot_multi(const float* __restrict__ q, int qs, const float* __restrict__ x, int d,
                                                     int lane, float* out) {
    double acc[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) acc[j] = 0.0;
    if (dot_vec_ok(q, x, d) && (qs & 3) == 0) {
        const float4* x4 = (const float4*)x;
        for (int i = lane; i < (d >> 2); i += 32) {
            const float4 b = __ldg(x4 + i);
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                const float4 a = *(const float4*)(q + (size_t)j * qs + (i << 2));
                acc[j] += (double)a.x * (double)b.x;
                acc[j] += (double)a.y * (double)b.y;
                acc[j] += (double)a.z * (double)b.z;
                acc[j] += (double)a.w * (double)b.w;
            }
        }
    } else {
        for (int i = lane; i < d; i += 32) {
            const float b = __ldg(x + i);
#pragma unroll
            for (int j = 0; j < NQ; ++j) acc[j] += (double)q[(size_t)j * qs + i] * (double)b;
        }
    }
#pragma unroll
    for (int j = 0; j < NQ; ++j) out[j] = (float)warp_reduce_f64(acc[j]);
}

// In-place descending bitonic sort of n (power of two) keys in shared memory by all `nthreads`
// threads of the CTA. Ends with a __syncthreads().
__device__ __forceinline__ void block_bitonic_sort_desc(uint64_t* keys, int n, int tid, int nthreads) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int i = tid; i < n; i += nthreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint64_t a = keys[i], b = keys[ixj];
                    const bool desc = ((i & k) == 0);
                    if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
                }
            }
        }
    }
    __syncthreads();
}

__host__ __device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace gdt
