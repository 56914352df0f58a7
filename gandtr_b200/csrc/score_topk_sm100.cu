// K3 -- query x database scoring fused with a streaming per-query top-k, for sm_100a.
// Replaces `scores = np.dot(vecs.T, qvecs); ranks = np.argsort(-scores, axis=0)` restricted to the
// first k ranks (mdir/components/optim/score/cirscore.py:71-72). The score matrix never reaches HBM.
//
// Pipeline (all on the caller's stream):
//   db_prepare_kernel      once per database shard: bf16 shadow copy + max row norm.
//   q_prepare_kernel       per call: bf16 copy of the queries, per-query scale and error bound, state reset.
//   score_filter_kernel    the hot kernel. Persistent, warp-specialised, one CTA per SM:
//        warp 0      TMA producer   cp.async.bulk.tensor (128B-swizzled [128 x 64] query and
//                                   [256 x 64] database boxes) into a 4-stage mbarrier ring
//        warp 1      MMA issuer     one elected thread issues tcgen05.mma.cta_group::1.kind::f16
//                                   (bf16 x bf16 -> fp32, M=128, N=256, K=16) into a double-buffered
//                                   TMEM accumulator (2 x 256 of the 512 columns)
//        warps 2..5  epilogue       tcgen05.ld 32 lanes x 32 columns; thread == query row, so the running
//                                   threshold of a query lives in one register and filtering needs no
//                                   cross-thread traffic. Survivors (rare after warm-up) are appended to a
//                                   per-query candidate list in global memory and counted in a per-query
//                                   256-bin score histogram from which the threshold is tightened.
//   topk_finalize_kernel   one CTA per query: final threshold from the histogram, exact fp32 re-scoring
//                          (fp64-accumulated) of the few survivors, bitonic sort by (score desc, index asc).
//
// Exactness. The bf16 pass only *filters*. With d_q = ||q||*max||x|| * (2^-8 + d*2^-22) (operand rounding
// plus fp32 accumulation, worst case) every row whose exact score could reach the top k has a coarse score
// >= t - 2*d_q where t is any value that >= k coarse scores are known to reach. Survivors are re-scored
// exactly, so returned scores and ranking are those of the exact kernel (score_exact_sm100.cu).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "select.cuh"

namespace gdt {

// ---- geometry -----------------------------------------------------------------------------------
constexpr int kBlockM = 128;   // queries per tile (TMEM lanes)
constexpr int kBlockN = 256;   // database rows per tile (TMEM columns)
constexpr int kBlockK = 64;    // bf16 elements per k-block == one 128-byte swizzle span
constexpr int kUmmaK = 16;
constexpr int kStages = 4;
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr int kBBytes = kBlockN * kBlockK * 2;   // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int kThreads = 192;
constexpr int kHistBins = 256;
constexpr uint32_t kTmemCols = 512;

struct QMeta {
    float scale;      // ||q|| * max||x||  (0 -> degenerate query)
    float inv_scale;
    float margin;     // 2 * d_q
    float pad;
};

// Per-query candidate capacity. Every database stripe of a query tile starts from the threshold published
// so far (initially -inf), so its first tile can insert all kBlockN rows and a few multiples of k follow
// before its histogram threshold bites: capacity grows with the number of concurrently started stripes.
__host__ __device__ inline int cand_capacity(int k, int n_stripes) {
    long long c = 16LL * k;
    const long long per_stripe = (long long)n_stripes * (kBlockN + 4LL * k);
    if (c < per_stripe) c = per_stripe;
    if (c < 4096) c = 4096;
    if (c > (1 << 22)) c = 1 << 22;
    return next_pow2((int)c);
}
__host__ __device__ inline int survivor_capacity(int k) {
    int c = 4 * k;
    if (c < 2048) c = 2048;
    return next_pow2(c);
}

// histogram bin of a coarse score relative to the query's scale: 32 bins per octave over [2^-8, 1)
__device__ __forceinline__ int score_bin(float v, float inv_scale) {
    const float x = v * inv_scale;
    if (!(x >= 0.00390625f)) return 0;
    const int b = (int)(__float_as_uint(x) >> 18) - (119 << 5);
    return b > kHistBins - 1 ? kHistBins - 1 : b;
}
// pass threshold implied by "at least k scores fell into bins >= b"
__device__ __forceinline__ float bin_threshold(int b, const QMeta& m) {
    if (b <= 0) return __int_as_float(0xff800000);
    const float edge = __uint_as_float((uint32_t)(b + (119 << 5)) << 18);
    return edge * m.scale * 0.999999f - m.margin;
}

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a pipeline bug must not hang the GPU. ~2 s at 2 GHz, then trap.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t it = 0;; ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if ((it & 1023u) == 1023u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000LL) __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row (1024 B) swizzle atoms stacked along M/N.
// start address >> 4 in [0,14), LBO (unused for swizzled K-major) = 1 in [16,30), SBO = 1024 B >> 4 in [32,46),
// descriptor version 1 in [46,48), layout SWIZZLE_128B (2) in [61,64).
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// kind::f16 instruction descriptor: D=f32 (1<<4), A=B=bf16 (1<<7, 1<<10), K-major both, N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t kInstrDesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBlockN >> 3) << 17) |
                                ((uint32_t)(kBlockM >> 4) << 24);

// ---- preparation kernels ------------------------------------------------------------------------

// one warp per row: bf16 (round-to-nearest-even) copy + atomic max of the row norm (positive floats
// order like their bit patterns)
__global__ void __launch_bounds__(256)
db_prepare_kernel(const float* __restrict__ db, long long ndb, int d, __nv_bfloat16* __restrict__ out,
                  float* __restrict__ norm_max) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * 8;
    float wmax = 0.f;
    for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < ndb; r += warps) {
        const float* x = db + (size_t)r * d;
        __nv_bfloat16* o = out + (size_t)r * d;
        float ss = 0.f;
        if ((d & 3) == 0 && ((((uintptr_t)x) | ((uintptr_t)o)) & 15) == 0) {
            for (int i = lane; i < (d >> 2); i += 32) {
                const float4 v = __ldg((const float4*)x + i);
                ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
                __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                uint2 pk;
                pk.x = *(uint32_t*)&lo;
                pk.y = *(uint32_t*)&hi;
                *((uint2*)o + i) = pk;
            }
        } else {
            for (int i = lane; i < d; i += 32) {
                const float v = x[i];
                ss += v * v;
                o[i] = __float2bfloat16_rn(v);
            }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
        wmax = fmaxf(wmax, sqrtf(ss));
    }
    if (lane == 0 && wmax > 0.f) atomicMax((int*)norm_max, __float_as_int(wmax));
}

// one warp per query: bf16 copy, scale / margin, state reset
__global__ void __launch_bounds__(256)
q_prepare_kernel(const float* __restrict__ q, int nq, int d, const float* __restrict__ db_norm_max,
                 __nv_bfloat16* __restrict__ qb, QMeta* __restrict__ meta, uint32_t* __restrict__ tau,
                 uint32_t* __restrict__ cnt, uint32_t* __restrict__ hist, int32_t* __restrict__ status) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (blockIdx.x == 0 && threadIdx.x < 4) status[threadIdx.x] = 0;
    if (r >= nq) return;
    const float* x = q + (size_t)r * d;
    float ss = 0.f;
    for (int i = lane; i < d; i += 32) {
        const float v = x[i];
        ss += v * v;
        qb[(size_t)r * d + i] = __float2bfloat16_rn(v);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
    for (int i = lane; i < kHistBins; i += 32) hist[(size_t)r * kHistBins + i] = 0;
    if (lane == 0) {
        QMeta m;
        m.scale = sqrtf(ss) * __ldg(db_norm_max) * 1.00001f;
        m.inv_scale = m.scale > 0.f ? 1.0f / m.scale : 0.f;
        m.margin = 2.0f * m.scale * (0.00390625f + (float)d * 2.384185791015625e-07f) * 1.001f;
        m.pad = 0.f;
        meta[r] = m;
        tau[r] = ordered_bits(__int_as_float(0xff800000));
        cnt[r] = 0;
    }
}

// ---- the hot kernel -----------------------------------------------------------------------------

struct FilterParams {
    int nq, d, k, cap;
    long long ndb;
    int n_qtiles, n_dtiles, n_stripes, stripe_len, n_items, n_kblocks;
    const QMeta* meta;
    uint32_t* tau;
    uint32_t* cnt;
    uint32_t* hist;
    uint64_t* cand;
};

// per-query threshold from the global histogram (scanned from the top, 128-bit loads through L2)
__device__ __forceinline__ float scan_threshold(const uint32_t* hist_row, int k, const QMeta& m) {
    const uint4* h4 = (const uint4*)hist_row;
    uint32_t cum = 0;
    for (int g = kHistBins / 32 - 1; g >= 0; --g) {
        uint4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldcg(h4 + g * 8 + j);
        uint32_t gs = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) gs += v[j].x + v[j].y + v[j].z + v[j].w;
        if (cum + gs >= (uint32_t)k) {
#pragma unroll
            for (int j = 7; j >= 0; --j) {
                const uint32_t e[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
                for (int c = 3; c >= 0; --c) {
                    cum += e[c];
                    if (cum >= (uint32_t)k) return bin_threshold(g * 32 + j * 4 + c, m);
                }
            }
        }
        cum += gs;
    }
    return __int_as_float(0xff800000);
}

__global__ void __launch_bounds__(kThreads, 1)
score_filter_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_db,
                    const FilterParams P) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte aligned operand ring, then the barriers
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    uint64_t* bars = (uint64_t*)(smem_gen + kStages * kStageBytes);
    // bars: [0,S) full, [S,2S) empty, [2S,2S+2) tmem_full, [2S+2,2S+4) tmem_empty, then tmem base slot
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kStages;
    const uint32_t bar_tfull = bar_empty + 8 * kStages, bar_tempty = bar_tfull + 16;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_db) : "memory");
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
                const int stripe = item / P.n_qtiles, qt = item - stripe * P.n_qtiles;
                const int t0 = stripe * P.stripe_len, t1 = min(t0 + P.stripe_len, P.n_dtiles);
                for (int t = t0; t < t1; ++t) {
                    for (int kb = 0; kb < P.n_kblocks; ++kb) {
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
                        mbar_arrive_expect_tx(bar_full + 8 * stage, kStageBytes);
                        tma_load_2d(sa, &map_q, bar_full + 8 * stage, kb * kBlockK, qt * kBlockM);
                        tma_load_2d(sb, &map_db, bar_full + 8 * stage, kb * kBlockK, t * kBlockN);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
                const int stripe = item / P.n_qtiles;
                const int t0 = stripe * P.stripe_len, t1 = min(t0 + P.stripe_len, P.n_dtiles);
                for (int t = t0; t < t1; ++t) {
                    mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * kBlockN;
                    for (int kb = 0; kb < P.n_kblocks; ++kb) {
                        mbar_wait(bar_full + 8 * stage, phase);
                        tc_fence_after();
                        const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
                        const uint64_t da = umma_smem_desc(sa), db = umma_smem_desc(sb);
#pragma unroll
                        for (int kk = 0; kk < kBlockK / kUmmaK; ++kk) {
                            // +32 bytes along K inside the swizzle span == +2 in the (>>4) address field
                            umma_bf16(tmem_d, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), kInstrDesc,
                                      (kb | kk) != 0 ? 1u : 0u);
                        }
                        umma_commit(bar_empty + 8 * stage);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(bar_tfull + 8 * acc);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else {
        // ===== epilogue: thread == query row =====
        const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter + 32)
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const float neg_inf = __int_as_float(0xff800000), pos_inf = __int_as_float(0x7f800000);
        uint32_t acc = 0, acc_phase = 0;
        for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
            const int stripe = item / P.n_qtiles, qt = item - stripe * P.n_qtiles;
            const int t0 = stripe * P.stripe_len, t1 = min(t0 + P.stripe_len, P.n_dtiles);
            const int qrow = qt * kBlockM + quarter * 32 + lane;
            const bool valid = qrow < P.nq;
            QMeta m;
            m.scale = 0.f; m.inv_scale = 0.f; m.margin = 0.f; m.pad = 0.f;
            float tau = pos_inf;
            if (valid) {
                m = P.meta[qrow];
                tau = from_ordered_bits(__ldcg(P.tau + qrow));
            }
            uint32_t* hist_row = P.hist + (size_t)(valid ? qrow : 0) * kHistBins;
            uint64_t* cand_row = P.cand + (size_t)(valid ? qrow : 0) * P.cap;
            uint32_t inserted = 0;
            int tiles_done = 0;
            for (int t = t0; t < t1; ++t) {
                if (valid && t != t0) tau = fmaxf(tau, from_ordered_bits(__ldcg(P.tau + qrow)));
                mbar_wait(bar_tfull + 8 * acc, acc_phase);
                tc_fence_after();
                const long long col0 = (long long)t * kBlockN;
                const int ncols = (int)min((long long)kBlockN, P.ndb - col0);
#pragma unroll 1
                for (int c = 0; c < kBlockN / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + lane_addr + acc * kBlockN + c * 32, v);
                    tmem_ld_wait();
                    float mx = __uint_as_float(v[0]);
#pragma unroll
                    for (int j = 1; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
                    if (mx >= tau) {
                        uint32_t mask = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (__uint_as_float(v[j]) >= tau && c * 32 + j < ncols) mask |= 1u << j;
                        if (mask) {
                            const uint32_t n = __popc(mask);
                            uint32_t slot = atomicAdd(P.cnt + qrow, n);
                            inserted += n;
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                if (mask & (1u << j)) {
                                    if (slot < (uint32_t)P.cap)
                                        cand_row[slot] = ((uint64_t)v[j] << 32) | (uint32_t)(col0 + c * 32 + j);
                                    ++slot;
                                    atomicAdd(hist_row + score_bin(__uint_as_float(v[j]), m.inv_scale), 1u);
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                ++tiles_done;
                // tighten the threshold: on a doubling schedule while warming up, whenever many candidates
                // went in, and when leaving the stripe
                if (valid && inserted != 0 &&
                    ((tiles_done & (tiles_done - 1)) == 0 || inserted >= 64 || t + 1 == t1)) {
                    __threadfence();
                    const float nt = scan_threshold(hist_row, P.k, m);
                    if (nt > tau) tau = nt;
                    atomicMax(P.tau + qrow, ordered_bits(tau));
                    inserted = 0;
                }
            }
        }
        (void)neg_inf;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- finalisation -------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
topk_finalize_kernel(const float* __restrict__ q, const float* __restrict__ db, int d, int dpad, int k, int cap, int scap,
                     long long index_base, const QMeta* __restrict__ meta, const uint32_t* __restrict__ cnt,
                     const uint32_t* __restrict__ hist, const uint64_t* __restrict__ cand, float* __restrict__ out_s,
                     int64_t* __restrict__ out_i, int32_t* __restrict__ status) {
    extern __shared__ __align__(16) uint8_t fsm[];
    uint64_t* keys = (uint64_t*)fsm;                      // [scap]
    float* qrow = (float*)(fsm + (size_t)scap * 8);       // [dpad]
    __shared__ uint32_t wsum[8];
    __shared__ int s_bin;
    __shared__ uint32_t s_ns;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int qi = blockIdx.x;
    const QMeta m = meta[qi];
    for (int i = tid; i < dpad; i += 256) qrow[i] = i < d ? q[(size_t)qi * d + i] : 0.f;
    if (tid == 0) { s_bin = 0; s_ns = 0; }
    // suffix counts over the 256 bins (thread == bin), largest bin whose suffix count reaches k
    const uint32_t h = hist[(size_t)qi * kHistBins + tid];
    uint32_t inc = h;  // inclusive suffix scan: reverse the lane order inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_down_sync(0xffffffffu, inc, o);
        if (lane + o < 32) inc += t;
    }
    if (lane == 0) wsum[wid] = inc;
    __syncthreads();
    uint32_t above = 0;
    for (int w = wid + 1; w < 8; ++w) above += wsum[w];
    const uint32_t suffix = inc + above;
    if (suffix >= (uint32_t)k) atomicMax(&s_bin, tid);
    __syncthreads();
    const float thr = bin_threshold(s_bin, m);

    const uint32_t total = cnt[qi];
    bool overflow = total > (uint32_t)cap;
    const uint32_t n = overflow ? (uint32_t)cap : total;
    const uint64_t* crow = cand + (size_t)qi * cap;
    for (uint32_t i = tid; i < n; i += 256) {
        const uint64_t c = __ldg(crow + i);
        if (__uint_as_float((uint32_t)(c >> 32)) >= thr) {
            const uint32_t pos = atomicAdd(&s_ns, 1u);
            if (pos < (uint32_t)scap) keys[pos] = c;
        }
    }
    __syncthreads();
    uint32_t ns = s_ns;
    if (ns > (uint32_t)scap) { overflow = true; ns = scap; }
    // exact re-scoring, one warp per survivor
    for (uint32_t j = wid; j < ns; j += 8) {
        const uint32_t row = (uint32_t)keys[j];
        const float s = warp_exact_dot(qrow, db + (size_t)row * d, d, lane);
        __syncwarp();
        if (lane == 0) keys[j] = rank_key(s, row);
    }
    int np = next_pow2((int)(ns > (uint32_t)k ? ns : (uint32_t)k));
    if (np > scap) np = scap;
    __syncthreads();
    for (int i = ns + tid; i < np; i += 256) keys[i] = 0ull;
    block_bitonic_sort_desc(keys, np, tid, 256);
    for (int i = tid; i < k; i += 256) {
        const bool ok = (uint32_t)i < ns;
        const uint64_t key = ok ? keys[i] : 0ull;
        out_s[(size_t)qi * k + i] = ok ? key_score(key) : __int_as_float(0xff800000);
        int64_t id = ok ? (int64_t)(index_base + (long long)key_index(key)) : (int64_t)-1;
        if (overflow && i == 0) id = -2;  // marks the query for the caller's exact fallback
        out_i[(size_t)qi * k + i] = id;
    }
    if (tid == 0) {
        atomicMax(status + 1, (int)ns);
        atomicMax(status + 3, (int)(total > 0x7fffffffu ? 0x7fffffff : total));
        if (overflow) {
            status[0] = GDT_ERR_CANDIDATE_OVERFLOW;
            atomicAdd(status + 2, 1);
        }
    }
}

// ---- host ---------------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static int make_bf16_map(CUtensorMap* map, const void* base, long long rows, int d, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) {
        snprintf(tls_error_buf(), 512, "cuTensorMapEncodeTiled entry point not found");
        return GDT_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)d * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(tls_error_buf(), 512, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return GDT_ERR_CUDA;
    }
    return GDT_OK;
}

struct TopkLayout {
    size_t qb, meta, tau, cnt, hist, cand, total;
    int cap;
};

static TopkLayout topk_layout(int nq, int d, int k, int n_stripes) {
    TopkLayout L;
    size_t off = 0;
    auto take = [&](size_t bytes) { off = align_up(off, 256); const size_t r = off; off += bytes; return r; };
    L.cap = cand_capacity(k, n_stripes);
    L.qb = take((size_t)nq * d * 2);
    L.meta = take((size_t)nq * sizeof(QMeta));
    L.tau = take((size_t)nq * 4);
    L.cnt = take((size_t)nq * 4);
    L.hist = take((size_t)nq * kHistBins * 4);
    L.cand = take((size_t)nq * L.cap * 8);
    L.total = align_up(off, 256);
    return L;
}

// pick the number of database stripes so that items fill whole waves of the persistent grid
static void plan_items(int n_qtiles, int n_dtiles, int sms, int& n_stripes, int& stripe_len) {
    double best = -1.0;
    n_stripes = 1;
    for (int s = 1; s <= n_dtiles && s <= 4096; ++s) {
        const int len = ceil_div(n_dtiles, s);
        const int s_eff = ceil_div(n_dtiles, len);
        const long long items = (long long)s_eff * n_qtiles;
        const long long waves = (items + sms - 1) / sms;
        // work per CTA is proportional to waves * len (+1 tile of pipeline fill per item)
        const double cost = (double)waves * (len + 0.5);
        const double ideal = (double)n_qtiles * n_dtiles / sms;
        const double eff = ideal / cost;
        if (eff > best + 1e-9) { best = eff; n_stripes = s_eff; }
        if (len <= 2) break;
    }
    stripe_len = ceil_div(n_dtiles, n_stripes);
    n_stripes = ceil_div(n_dtiles, stripe_len);
}

}  // namespace gdt

using namespace gdt;

extern "C" size_t gdt_db_prepare_workspace_bytes(long long ndb, int d) {
    (void)ndb; (void)d;
    return 256;
}

extern "C" int gdt_db_prepare(const float* db, long long ndb, int d, void* db_bf16, float* db_norm_max, void* ws,
                              size_t ws_bytes, void* stream_) {
    (void)ws; (void)ws_bytes;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!db || !db_bf16 || !db_norm_max || ndb <= 0 || d <= 0) return GDT_ERR_INVALID_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return GDT_ERR_NO_DEVICE; }
    GDT_CUDA(cudaMemsetAsync(db_norm_max, 0, sizeof(float), stream));
    const int sms = sm_count_current_device();
    long long blocks = ceil_div_ll(ndb, 8);
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    db_prepare_kernel<<<(unsigned)blocks, 256, 0, stream>>>(db, ndb, d, (__nv_bfloat16*)db_bf16, db_norm_max);
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}

extern "C" size_t gdt_score_topk_workspace_bytes(int nq, long long ndb, int d, int k) {
    if (nq <= 0 || ndb <= 0 || d <= 0 || k <= 0) return 0;
    int n_stripes, stripe_len;
    plan_items(ceil_div(nq, kBlockM), (int)ceil_div_ll(ndb, kBlockN), sm_count_current_device(), n_stripes, stripe_len);
    return topk_layout(nq, d, k, n_stripes).total + 256;
}

extern "C" int gdt_score_topk(const float* q, const float* db, const void* db_bf16, const float* db_norm_max, int nq,
                              long long ndb, int d, int k, long long index_base, float* top_scores, int64_t* top_idx,
                              int32_t* status_dev, void* ws, size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!q || !db || !db_bf16 || !db_norm_max || !top_scores || !top_idx || !status_dev || !ws)
        return GDT_ERR_INVALID_ARGUMENT;
    if (nq <= 0 || ndb <= 0 || d <= 0 || k <= 0) return GDT_ERR_INVALID_ARGUMENT;
    if ((d & 7) != 0 || d > 8192 || k > 1024 || ndb > 0x7fffffffLL || index_base < 0 ||
        index_base + ndb > 0xffffffffLL)
        return GDT_ERR_UNSUPPORTED;
    if ((((uintptr_t)db_bf16) & 15) != 0) return GDT_ERR_INVALID_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return GDT_ERR_NO_DEVICE; }
    if (ws_bytes < gdt_score_topk_workspace_bytes(nq, ndb, d, k) || (((uintptr_t)ws) & 255)) return GDT_ERR_WORKSPACE_TOO_SMALL;

    FilterParams P;
    P.n_qtiles = ceil_div(nq, kBlockM);
    P.n_dtiles = (int)ceil_div_ll(ndb, kBlockN);
    const int sms = sm_count_current_device();
    plan_items(P.n_qtiles, P.n_dtiles, sms, P.n_stripes, P.stripe_len);
    const TopkLayout L = topk_layout(nq, d, k, P.n_stripes);
    char* base = (char*)ws;
    __nv_bfloat16* qb = (__nv_bfloat16*)(base + L.qb);
    QMeta* meta = (QMeta*)(base + L.meta);
    uint32_t* tau = (uint32_t*)(base + L.tau);
    uint32_t* cnt = (uint32_t*)(base + L.cnt);
    uint32_t* hist = (uint32_t*)(base + L.hist);
    uint64_t* cand = (uint64_t*)(base + L.cand);

    q_prepare_kernel<<<ceil_div(nq, 8), 256, 0, stream>>>(q, nq, d, db_norm_max, qb, meta, tau, cnt, hist, status_dev);
    GDT_LAUNCH_CHECK();

    CUtensorMap map_q, map_db;
    int rc = make_bf16_map(&map_q, qb, nq, d, kBlockM);
    if (rc != GDT_OK) return rc;
    rc = make_bf16_map(&map_db, db_bf16, ndb, d, kBlockN);
    if (rc != GDT_OK) return rc;

    P.nq = nq; P.d = d; P.k = k; P.cap = L.cap; P.ndb = ndb;
    P.n_items = P.n_stripes * P.n_qtiles;
    P.n_kblocks = ceil_div(d, kBlockK);
    P.meta = meta; P.tau = tau; P.cnt = cnt; P.hist = hist; P.cand = cand;

    static bool attr_set = false;
    if (!attr_set) {
        GDT_CUDA(cudaFuncSetAttribute(score_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_set = true;
    }
    const int grid = P.n_items < sms ? P.n_items : sms;
    score_filter_kernel<<<grid, kThreads, kSmemBytes, stream>>>(map_q, map_db, P);
    GDT_LAUNCH_CHECK();

    const int scap = survivor_capacity(k);
    const int dpad = (d + 3) & ~3;
    const size_t fsmem = (size_t)scap * 8 + (size_t)dpad * 4;
    static size_t fattr = 0;
    if (fsmem > 48 * 1024 && fsmem > fattr) {
        GDT_CUDA(cudaFuncSetAttribute(topk_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
        fattr = fsmem;
    }
    topk_finalize_kernel<<<nq, 256, fsmem, stream>>>(q, db, d, dpad, k, L.cap, scap, index_base, meta, cnt, hist, cand,
                                                     top_scores, top_idx, status_dev);
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}
