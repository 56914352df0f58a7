// K3 -- query x database scoring fused with a streaming per-query top-k, for sm_100a.
// Replaces `scores = np.dot(vecs.T, qvecs); ranks = np.argsort(-scores, axis=0)` restricted to the
// first k ranks (mdir/components/optim/score/cirscore.py:71-72). The score matrix never reaches HBM.
//
// Pipeline (all on the caller's stream):
//   db_norm_kernel / db_convert_kernel   once per database shard: max row norm, then the fp16 shadow copy (rows scaled by
//                          an exact power of two so that the largest norm is in (0.5, 1]) and the largest norm of the
//                          rounding residual x*s - fp16(x*s).
//   q_prepare_kernel       per call: fp16 copy of the (power-of-two scaled) queries, per-query error bound, state reset.
//   score_filter_kernel    the hot kernel, launched twice (a short "seed" range of database tiles that establishes
//                          per-query thresholds, then the rest striped over the machine). Persistent, warp-specialised,
//                          one CTA per SM:
//        warp 0      TMA producer   cp.async.bulk.tensor (128B-swizzled [128 x 64] query and [256 x 64] database boxes)
//                                   into a 4-stage mbarrier ring
//        warp 1      MMA issuer     one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (fp16 x fp16 -> fp32,
//                                   M=128, N=256, K=16) into a double-buffered TMEM accumulator (2 x 256 of 512 columns)
//        warps 2..5  epilogue       tcgen05.ld 32 lanes x 32 columns; thread == query row, so the running threshold of a
//                                   query lives in one register and filtering needs no cross-thread traffic. Survivors
//                                   (rare after warm-up) go to a candidate segment PRIVATE to (query, stripe) -- the slot
//                                   counter is a register, no returning atomics -- and are counted in a per-query 256-bin
//                                   score histogram (fire-and-forget reductions) from which the threshold is tightened.
//   topk_finalize_kernel   one CTA per query: final threshold from the histogram, exact fp32 re-scoring
//                          (fp64-accumulated) of the few survivors, bitonic sort by (score desc, index asc).
//
// Exactness. The fp16 pass only *filters*. With q~ = fp16(q*sq), x~ = fp16(x*sx) (sq, sx exact powers of two) the coarse
// score c = <q~, x~> (fp32-accumulated on the tensor cores) differs from the exact scaled score sq*sx*<q, x> by at most
//     E_q = ||q*sq - q~|| * max||x*sx||  +  ||q~|| * max||x*sx - x~||  +  acc(d) * ||q~|| * max||x~||
// (Cauchy-Schwarz on the two MEASURED rounding residuals; acc(d) bounds the fp32 accumulation of d products, one rounding
// per K=16 MMA plus the alignment error inside it). Every row whose exact score can reach the top k therefore has a coarse
// score >= t - 2*E_q, where t is any value that >= k coarse scores are known to reach. Survivors are re-scored exactly
// from the fp32 rows, so the returned scores and ranking are those of the exact kernel (score_exact_sm100.cu).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "select.cuh"
#include "tc_common.cuh"

namespace gdt {

// ---- geometry -----------------------------------------------------------------------------------
constexpr int kBlockM = 128;   // queries per tile (TMEM lanes)
constexpr int kBlockN = 256;   // database rows per tile (TMEM columns)
constexpr int kBlockK = 64;    // fp16 elements per k-block == one 128-byte swizzle span
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

// Candidate segments. Segment 0 belongs to the seed pass (its capacity covers every row of the seed range, so it
// cannot overflow); segments 1..n_stripes belong to the stripes of the main pass, which start from the seed threshold.
__host__ __device__ inline int seed_tiles_for(int k) {
    int t = (16 * k + kBlockN - 1) / kBlockN;
    return t < 32 ? 32 : t;
}
constexpr int kSeedDiv = 0;      // see topk_plan
__host__ __device__ inline int stripe_capacity(int k) {
    int c = 8 * k;
    if (c < 1024) c = 1024;
    return next_pow2(c);
}
__host__ __device__ inline int survivor_capacity(int k) {
    int c = 8 * k;
    if (c < 2048) c = 2048;
    return next_pow2(c);       // <= 8192 keys (64 KB of shared memory) at k = 1024
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

// kind::f16 instruction descriptor: D=f32 (1<<4), A=B=fp16 (format 0 at [7,10) and [10,13)), K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t kInstrDesc = (1u << 4) | ((uint32_t)(kBlockN >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);

// ---- preparation kernels ------------------------------------------------------------------------

// db_stats layout (device float[4]): [0] max row norm, [1] power-of-two scale sx, [2] max ||x*sx - fp16(x*sx)||,
// [3] max ||fp16(x*sx)||
__device__ __forceinline__ float pow2_scale_for(float norm_max) {
    // largest power of two s with norm_max * s <= 1 (1 for empty / zero data)
    if (!(norm_max > 0.f)) return 1.0f;
    int e;
    const float m = frexpf(norm_max, &e);      // norm_max = m * 2^e, m in [0.5, 1)
    return ldexpf(1.0f, (m == 0.5f) ? -(e - 1) : -e);
}

// one warp per row: max of the row norms (positive floats order like their bit patterns)
__global__ void __launch_bounds__(256)
db_norm_kernel(const float* __restrict__ db, long long ndb, int d, float* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * 8;
    float wmax = 0.f;
    for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < ndb; r += warps) {
        const float* x = db + (size_t)r * d;
        float ss = 0.f;
        if ((d & 3) == 0 && (((uintptr_t)x) & 15) == 0) {
            for (int i = lane; i < (d >> 2); i += 32) {
                const float4 v = __ldg((const float4*)x + i);
                ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            }
        } else {
            for (int i = lane; i < d; i += 32) { const float v = x[i]; ss += v * v; }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
        wmax = fmaxf(wmax, sqrtf(ss));
    }
    if (lane == 0 && wmax > 0.f) atomicMax((int*)stats, __float_as_int(wmax));
}

// one warp per row: fp16 (round-to-nearest-even) copy of x*sx + maxima of the residual and shadow norms
__global__ void __launch_bounds__(256)
db_convert_kernel(const float* __restrict__ db, long long ndb, int d, __half* __restrict__ out, float* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * 8;
    const float sx = pow2_scale_for(__ldcg(stats));
    float rmax = 0.f, hmax = 0.f;
    for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < ndb; r += warps) {
        const float* x = db + (size_t)r * d;
        __half* o = out + (size_t)r * d;
        float rr = 0.f, hh = 0.f;
        if ((d & 3) == 0 && ((((uintptr_t)x) | ((uintptr_t)o)) & 15) == 0) {
            for (int i = lane; i < (d >> 2); i += 32) {
                const float4 v = __ldg((const float4*)x + i);
                const float a[4] = {v.x * sx, v.y * sx, v.z * sx, v.w * sx};
                __half hv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    hv[j] = __float2half_rn(a[j]);
                    const float f = __half2float(hv[j]), e = a[j] - f;
                    rr += e * e;
                    hh += f * f;
                }
                uint2 pk;
                pk.x = (uint32_t)__half_as_ushort(hv[0]) | ((uint32_t)__half_as_ushort(hv[1]) << 16);
                pk.y = (uint32_t)__half_as_ushort(hv[2]) | ((uint32_t)__half_as_ushort(hv[3]) << 16);
                *((uint2*)o + i) = pk;
            }
        } else {
            for (int i = lane; i < d; i += 32) {
                const float a = x[i] * sx;
                const __half hv = __float2half_rn(a);
                const float f = __half2float(hv), e = a - f;
                rr += e * e;
                hh += f * f;
                o[i] = hv;
            }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            rr += __shfl_xor_sync(0xffffffffu, rr, s);
            hh += __shfl_xor_sync(0xffffffffu, hh, s);
        }
        rmax = fmaxf(rmax, sqrtf(rr));
        hmax = fmaxf(hmax, sqrtf(hh));
    }
    if (lane == 0) {
        if (rmax > 0.f) atomicMax((int*)stats + 2, __float_as_int(rmax));
        if (hmax > 0.f) atomicMax((int*)stats + 3, __float_as_int(hmax));
        if (blockIdx.x == 0 && threadIdx.x == 0) stats[1] = sx;
    }
}

// one warp per query: fp16 copy of q*sq, scale / error bound, state reset
__global__ void __launch_bounds__(256)
q_prepare_kernel(const float* __restrict__ q, int nq, int d, const float* __restrict__ db_stats,
                 __half* __restrict__ qb, QMeta* __restrict__ meta, uint32_t* __restrict__ tau,
                 uint32_t* __restrict__ cnt, int n_segs, uint32_t* __restrict__ hist, int32_t* __restrict__ status) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (blockIdx.x == 0 && threadIdx.x < 4) status[threadIdx.x] = 0;
    if (r >= nq) return;
    const float* x = q + (size_t)r * d;
    float ss = 0.f;
    for (int i = lane; i < d; i += 32) { const float v = x[i]; ss += v * v; }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
    const float sq = pow2_scale_for(sqrtf(ss) * 1.0001f);
    float rr = 0.f, hh = 0.f;
    for (int i = lane; i < d; i += 32) {
        const float a = x[i] * sq;
        const __half hv = __float2half_rn(a);
        const float f = __half2float(hv), e = a - f;
        rr += e * e;
        hh += f * f;
        qb[(size_t)r * d + i] = hv;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        rr += __shfl_xor_sync(0xffffffffu, rr, s);
        hh += __shfl_xor_sync(0xffffffffu, hh, s);
    }
    for (int i = lane; i < kHistBins; i += 32) hist[(size_t)r * kHistBins + i] = 0;
    for (int i = lane; i < n_segs; i += 32) cnt[(size_t)r * n_segs + i] = 0;
    if (lane == 0) {
        const float xs_max = __ldg(db_stats) * __ldg(db_stats + 1);          // max ||x * sx||  (<= 1)
        const float dx_max = __ldg(db_stats + 2), hx_max = __ldg(db_stats + 3);
        const float dq = sqrtf(rr), hq = sqrtf(hh);
        // fp32 accumulation on the tensor cores: one rounding per K=16 MMA plus the alignment error inside it
        const float acc = ((float)(d / 16 + 16)) * 9.5367431640625e-07f;     // * 2^-20: 2x the bound, the accumulator's
                                                                             // internal alignment width is undocumented
        const float eq = (dq * xs_max + hq * dx_max + acc * hq * hx_max) * 1.001f;
        QMeta m;
        m.scale = hq * hx_max * 1.0001f;                                     // |coarse score| <= scale
        m.inv_scale = m.scale > 0.f ? 1.0f / m.scale : 0.f;
        m.margin = 2.0f * eq;
        m.pad = sq;
        meta[r] = m;
        tau[r] = ordered_bits(__int_as_float(0xff800000));
    }
}

// ---- the hot kernel -----------------------------------------------------------------------------

struct FilterParams {
    int nq, d, k;
    long long ndb;
    int n_qtiles, n_qgroups, n_kblocks;   // n_qgroups = ceil(n_qtiles / CLUSTER): query tiles handled together by one cluster
    int tile_begin, tile_end;          // database tile range of this launch
    int n_stripes, stripe_len, n_items;
    int seg_first;                     // candidate segment of stripe 0 of this launch
    int n_segs, n_seed, cap0, cap1;    // segments per query; the first n_seed (seed pass) hold cap0 entries, the rest cap1
    const QMeta* meta;
    uint32_t* tau;
    uint32_t* cnt;                     // [nq][n_segs]
    uint32_t* hist;
    uint64_t* cand;                    // [nq][n_seed * cap0 + (n_segs - n_seed) * cap1]
    float* dump;                       // debug (gdt_debug_k3_coarse_scores): every coarse score, [nq][dump_ld]; else null
    long long dump_ld;
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

// CLUSTER > 1: CLUSTER CTAs of a thread-block cluster work on CLUSTER consecutive query tiles against the SAME stripe of
// database tiles; each loads 1/CLUSTER of every database tile and multicasts it to all of them, so the L2 -> SM operand
// traffic per CTA drops from 48 KB to 16 + 32/CLUSTER KB per K block (the operand feed, not the tensor pipe, limits the
// single-CTA version). A stage is released to the producers only when the MMA warps of all CTAs have consumed it.
template <int CLUSTER>
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
            mbar_init(bar_empty + 8 * s, CLUSTER);
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
    if (CLUSTER > 1) cluster_sync_all();      // peers' barriers are initialised before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int crank = CLUSTER > 1 ? (int)cluster_ctarank() : 0;
    const int cluster_id = blockIdx.x / CLUSTER, n_clusters = gridDim.x / CLUSTER;
    constexpr uint16_t kMask = (uint16_t)((1u << CLUSTER) - 1u);
    constexpr int kSliceRows = kBlockN / CLUSTER, kSliceBytes = kBBytes / CLUSTER;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int item = cluster_id; item < P.n_items; item += n_clusters) {
                const int stripe = item / P.n_qgroups, qt = (item - stripe * P.n_qgroups) * CLUSTER + crank;
                const int t0 = P.tile_begin + stripe * P.stripe_len, t1 = min(t0 + P.stripe_len, P.tile_end);
                for (int t = t0; t < t1; ++t) {
                    for (int kb = 0; kb < P.n_kblocks; ++kb) {
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
                        mbar_arrive_expect_tx(bar_full + 8 * stage, kStageBytes);
                        tma_load_2d(sa, &map_q, bar_full + 8 * stage, kb * kBlockK, qt * kBlockM);
                        if (CLUSTER == 1)
                            tma_load_2d(sb, &map_db, bar_full + 8 * stage, kb * kBlockK, t * kBlockN);
                        else        // my slice of the database tile, delivered to every CTA of the cluster
                            tma_load_2d_mc(sb + crank * kSliceBytes, &map_db, bar_full + 8 * stage, kb * kBlockK,
                                           t * kBlockN + crank * kSliceRows, kMask);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int item = cluster_id; item < P.n_items; item += n_clusters) {
                const int stripe = item / P.n_qgroups;
                const int t0 = P.tile_begin + stripe * P.stripe_len, t1 = min(t0 + P.stripe_len, P.tile_end);
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
                            umma_f16(tmem_d, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), kInstrDesc,
                                     (kb | kk) != 0 ? 1u : 0u);
                        }
                        if (CLUSTER == 1) umma_commit(bar_empty + 8 * stage);
                        else umma_commit_mc(bar_empty + 8 * stage, kMask);    // frees the stage in every CTA's producer
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
        for (int item = cluster_id; item < P.n_items; item += n_clusters) {
            const int stripe = item / P.n_qgroups, qt = (item - stripe * P.n_qgroups) * CLUSTER + crank;
            const int t0 = P.tile_begin + stripe * P.stripe_len, t1 = min(t0 + P.stripe_len, P.tile_end);
            const int qrow = qt * kBlockM + quarter * 32 + lane;
            const bool valid = qrow < P.nq;
            const int seg = P.seg_first + stripe;
            const uint32_t segcap = seg < P.n_seed ? (uint32_t)P.cap0 : (uint32_t)P.cap1;
            QMeta m;
            m.scale = 0.f; m.inv_scale = 0.f; m.margin = 0.f; m.pad = 0.f;
            float tau = pos_inf;
            if (valid) {
                m = P.meta[qrow];
                tau = from_ordered_bits(__ldcg(P.tau + qrow));
            }
            const size_t qv = valid ? (size_t)qrow : 0;
            uint32_t* hist_row = P.hist + qv * kHistBins;
            uint64_t* seg_row = P.cand + qv * ((size_t)P.n_seed * P.cap0 + (size_t)(P.n_segs - P.n_seed) * P.cap1) +
                                (seg < P.n_seed ? (size_t)seg * P.cap0
                                                : (size_t)P.n_seed * P.cap0 + (size_t)(seg - P.n_seed) * P.cap1);
            uint32_t slot = 0, inserted = 0;
            int tiles_done = 0;
            uint32_t tau_pub = ordered_bits(tau);     // threshold published by all stripes of this query, one tile stale:
            for (int t = t0; t < t1; ++t) {           // the load is issued a tile ahead so its latency is never exposed
                tau = fmaxf(tau, from_ordered_bits(tau_pub));
                if (valid) tau_pub = __ldcg(P.tau + qrow);
                mbar_wait(bar_tfull + 8 * acc, acc_phase);
                tc_fence_after();
                const long long col0 = (long long)t * kBlockN;
                const int ncols = (int)min((long long)kBlockN, P.ndb - col0);
#pragma unroll 1
                for (int c = 0; c < kBlockN / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + lane_addr + acc * kBlockN + c * 32, v);
                    tmem_ld_wait();
                    if (c * 32 + 32 > ncols) {      // ragged last tile: rows past the end of the shard never qualify --
#pragma unroll                                  // NaN fails every `>= tau` test, even while tau is still -inf
                        for (int j = 0; j < 32; ++j)
                            if (c * 32 + j >= ncols) v[j] = 0x7fc00000u;
                    }
                    if (P.dump != nullptr && valid) {      // debug only: the raw tensor-core scores, for the error-bound tests
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (c * 32 + j < ncols) P.dump[(size_t)qrow * P.dump_ld + col0 + c * 32 + j] = __uint_as_float(v[j]);
                    }
                    float mg[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float x = __uint_as_float(v[g * 8]);
#pragma unroll
                        for (int j = 1; j < 8; ++j) x = fmaxf(x, __uint_as_float(v[g * 8 + j]));
                        mg[g] = x;
                    }
                    if (fmaxf(fmaxf(mg[0], mg[1]), fmaxf(mg[2], mg[3])) >= tau) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (mg[g] >= tau) {
#pragma unroll
                                for (int jj = 0; jj < 8; ++jj) {
                                    const int j = g * 8 + jj;
                                    if (__uint_as_float(v[j]) >= tau) {
                                        if (slot < segcap)
                                            seg_row[slot] = ((uint64_t)v[j] << 32) | (uint32_t)(col0 + c * 32 + j);
                                        ++slot;
                                        // bin 0 (everything below 2^-8 of the scale) never yields a threshold:
                                        // not counting it keeps the same-address reductions off the L2 atomic units
                                        const int bin = score_bin(__uint_as_float(v[j]), m.inv_scale);
                                        if (bin) red_add_u32(hist_row + bin, 1u);
                                    }
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
                const uint32_t fresh = slot - inserted;
                if (valid && fresh != 0 && ((tiles_done & (tiles_done - 1)) == 0 || fresh >= 32 || t + 1 == t1)) {
                    __threadfence();
                    const float nt = scan_threshold(hist_row, P.k, m);
                    if (nt > tau) tau = nt;
                    red_max_u32(P.tau + qrow, ordered_bits(tau));
                    inserted = slot;
                }
            }
            if (valid) P.cnt[(size_t)qrow * P.n_segs + seg] = slot;
        }
        (void)neg_inf;
    }

    tc_fence_before();
    __syncthreads();
    if (CLUSTER > 1) cluster_sync_all();      // nobody leaves while a peer may still multicast into / signal this CTA
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- finalisation -------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
topk_finalize_kernel(const float* __restrict__ q, const float* __restrict__ db, int d, int dpad, int k, int n_segs, int n_seed,
                     int cap0, int cap1, int scap, long long index_base, const QMeta* __restrict__ meta,
                     const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ hist, const uint64_t* __restrict__ cand,
                     float* __restrict__ out_s, int64_t* __restrict__ out_i, int32_t* __restrict__ status) {
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

    // gather the candidates that pass the final threshold from every segment of this query
    bool overflow = false;
    uint32_t total = 0;
    const uint64_t* qcand = cand + (size_t)qi * ((size_t)n_seed * cap0 + (size_t)(n_segs - n_seed) * cap1);
    for (int sgm = 0; sgm < n_segs; ++sgm) {
        const uint32_t cap = sgm < n_seed ? (uint32_t)cap0 : (uint32_t)cap1;
        const uint32_t have = cnt[(size_t)qi * n_segs + sgm];
        total += have;
        if (have > cap) overflow = true;
        const uint32_t n = have > cap ? cap : have;
        const uint64_t* crow = qcand + (sgm < n_seed ? (size_t)sgm * cap0 : (size_t)n_seed * cap0 + (size_t)(sgm - n_seed) * cap1);
        for (uint32_t i = tid; i < n; i += 256) {
            const uint64_t c = __ldg(crow + i);
            if (__uint_as_float((uint32_t)(c >> 32)) >= thr) {
                const uint32_t pos = atomicAdd(&s_ns, 1u);
                if (pos < (uint32_t)scap) keys[pos] = c;
            }
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

static int make_f16_map(CUtensorMap* map, const void* base, long long rows, int d, int box_rows) {
    return make_tile_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, rows, d, box_rows);
}

// pick the number of database stripes so that items fill whole waves of the persistent grid
static void plan_items(int n_qtiles, int n_dtiles, int sms, int& n_stripes, int& stripe_len) {
    double best = -1.0;
    n_stripes = 1;
    if (n_dtiles <= 0) { n_stripes = 0; stripe_len = 1; return; }
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

// Cluster size of the filter kernel: 2 when there are at least two query tiles to pair (GDT_DEBUG_K3_CLUSTER = 1 | 2 | 4
// overrides, for A/B timing). Decided from (nq) alone so that every entry point derives the same workspace layout.
static int topk_cluster_size(int n_qtiles) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("GDT_DEBUG_K3_CLUSTER");
        forced = e ? atoi(e) : 0;
        if (forced != 1 && forced != 2 && forced != 4) forced = 0;
    }
    int c = forced ? forced : 2;
    while (c > 1 && n_qtiles < c) c >>= 1;
    return c;
}

template <int CLUSTER>
static int max_active_clusters_t() {
    static int cached_dev[32] = {0};
    int& cached = cached_dev[current_device_slot()];
    if (cached) return cached;
    int n = 0;
    if (cudaFuncSetAttribute(score_filter_kernel<CLUSTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) == cudaSuccess) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(sm_count_current_device() / CLUSTER * CLUSTER), 1, 1);
        cfg.blockDim = dim3(kThreads, 1, 1);
        cfg.dynamicSmemBytes = kSmemBytes;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CLUSTER;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&n, score_filter_kernel<CLUSTER>, &cfg) != cudaSuccess) n = 0;
    }
    cudaGetLastError();
    if (n <= 0) n = sm_count_current_device() / CLUSTER - 2;   // conservative guess (odd-sized GPCs strand an SM each)
    if (n < 1) n = 1;
    cached = n;
    return n;
}
// clusters of `cluster` CTAs (one CTA per SM) that can be resident at once
static int max_active_clusters(int cluster) {
    if (cluster == 4) return max_active_clusters_t<4>();
    if (cluster == 2) return max_active_clusters_t<2>();
    return sm_count_current_device();
}

struct TopkPlan {
    int cluster, n_qgroups, n_units;   // CTAs per cluster, query-tile groups, clusters the grid can hold
    int n_qtiles, n_dtiles, seed_tiles, n_seed, seed_len, n_stripes, stripe_len, n_segs, cap0, cap1;
    size_t qb, meta, tau, cnt, hist, cand, total;
};

static TopkPlan topk_plan(int nq, long long ndb, int d, int k) {
    TopkPlan L;
    L.n_qtiles = ceil_div(nq, kBlockM);
    L.n_dtiles = (int)ceil_div_ll(ndb, kBlockN);
    L.seed_tiles = seed_tiles_for(k) < L.n_dtiles ? seed_tiles_for(k) : L.n_dtiles;
    {
        // Small shards (a row-sharded database on many GPUs): the cold seed pass runs at about half the rate of the main
        // pass, so it is capped at 1 / kSeedDiv of the shard's tiles (never below 8 tiles = 2048 rows); the main pass
        // starts from a looser threshold and tightens it from the running histogram. GDT_DEBUG_K3_SEED_DIV overrides
        // (0 = no cap), for A/B timing.
        static int div = -1;
        if (div < 0) {
            const char* e = getenv("GDT_DEBUG_K3_SEED_DIV");
            div = e ? atoi(e) : kSeedDiv;
            if (div < 0) div = 0;
        }
        if (div > 0) {
            int cap = L.n_dtiles / div;
            if (cap < 8) cap = 8;
            if (cap < L.seed_tiles) L.seed_tiles = cap;
        }
    }
    const int sms = sm_count_current_device();
    L.cluster = topk_cluster_size(L.n_qtiles);
    L.n_qgroups = ceil_div(L.n_qtiles, L.cluster);
    L.n_units = max_active_clusters(L.cluster);
    (void)sms;
    plan_items(L.n_qgroups, L.n_dtiles - L.seed_tiles, L.n_units, L.n_stripes, L.stripe_len);
    // The seed range itself is striped over the machine: every seed stripe starts cold and its segment can hold every row
    // of the stripe, so it cannot overflow however cold the threshold is. The stripe count minimises the makespan
    // waves x (stripe length + pipeline fill) -- e.g. 40 query groups on 74 clusters: 1 stripe of 32 tiles leaves 34
    // clusters idle for 32 tile times, 3 stripes of 11 run in 2 waves of 11 -- with at least 8 tiles (2048 rows) per
    // stripe so that a cold stripe still reaches a useful threshold.
    {
        int best_s = 1;
        double best_cost = 1e300;
        const int max_s = L.seed_tiles / 8 > 1 ? L.seed_tiles / 8 : 1;
        for (int sd = 1; sd <= max_s; ++sd) {
            const int len = ceil_div(L.seed_tiles, sd);
            const int s_eff = ceil_div(L.seed_tiles, len);
            const long long items = (long long)s_eff * L.n_qgroups;
            const long long waves = (items + L.n_units - 1) / L.n_units;
            const double cost = (double)waves * (len + 0.5);
            if (cost < best_cost - 1e-9) { best_cost = cost; best_s = s_eff; }
        }
        L.n_seed = best_s;
    }
    L.seed_len = ceil_div(L.seed_tiles, L.n_seed);
    L.n_seed = ceil_div(L.seed_tiles, L.seed_len);
    L.n_segs = L.n_seed + L.n_stripes;
    L.cap0 = L.seed_len * kBlockN;
    L.cap1 = stripe_capacity(k);
    size_t off = 0;
    auto take = [&](size_t bytes) { off = align_up(off, 256); const size_t r = off; off += bytes; return r; };
    L.qb = take((size_t)nq * d * 2);
    L.meta = take((size_t)nq * sizeof(QMeta));
    L.tau = take((size_t)nq * 4);
    L.cnt = take((size_t)nq * L.n_segs * 4);
    L.hist = take((size_t)nq * kHistBins * 4);
    L.cand = take((size_t)nq * ((size_t)L.n_seed * L.cap0 + (size_t)L.n_stripes * L.cap1) * 8);
    L.total = align_up(off, 256);
    return L;
}

}  // namespace gdt

using namespace gdt;

extern "C" size_t gdt_db_prepare_workspace_bytes(long long ndb, int d) {
    (void)ndb; (void)d;
    return 256;
}

static int db_prepare_impl(const float* db, long long ndb, int d, void* db_f16, float* db_stats, bool norm_phase,
                           bool convert_phase, cudaStream_t stream) {
    if (!db || !db_stats || ndb <= 0 || d <= 0 || (convert_phase && !db_f16)) return GDT_ERR_INVALID_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return GDT_ERR_NO_DEVICE; }
    const int sms = sm_count_current_device();
    long long blocks = ceil_div_ll(ndb, 8);
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    if (norm_phase) {
        GDT_CUDA(cudaMemsetAsync(db_stats, 0, 4 * sizeof(float), stream));
        db_norm_kernel<<<(unsigned)blocks, 256, 0, stream>>>(db, ndb, d, db_stats);
        GDT_LAUNCH_CHECK();
    }
    if (convert_phase) {
        GDT_CUDA(cudaMemsetAsync(db_stats + 1, 0, 3 * sizeof(float), stream));
        db_convert_kernel<<<(unsigned)blocks, 256, 0, stream>>>(db, ndb, d, (__half*)db_f16, db_stats);
        GDT_LAUNCH_CHECK();
    }
    return GDT_OK;
}

extern "C" int gdt_db_prepare(const float* db, long long ndb, int d, void* db_f16, float* db_stats, void* ws,
                              size_t ws_bytes, void* stream_) {
    (void)ws; (void)ws_bytes;
    return db_prepare_impl(db, ndb, d, db_f16, db_stats, true, true, (cudaStream_t)stream_);
}

extern "C" int gdt_db_prepare_norm(const float* db, long long ndb, int d, float* db_stats, void* stream_) {
    return db_prepare_impl(db, ndb, d, nullptr, db_stats, true, false, (cudaStream_t)stream_);
}

extern "C" int gdt_db_prepare_convert(const float* db, long long ndb, int d, void* db_f16, float* db_stats, void* stream_) {
    return db_prepare_impl(db, ndb, d, db_f16, db_stats, false, true, (cudaStream_t)stream_);
}

extern "C" size_t gdt_score_topk_workspace_bytes(int nq, long long ndb, int d, int k) {
    if (nq <= 0 || ndb <= 0 || d <= 0 || k <= 0) return 0;
    return topk_plan(nq, ndb, d, k).total + 256;
}

template <int CLUSTER>
static int launch_filter_t(int n_units, const CUtensorMap& map_q, const CUtensorMap& map_db, const FilterParams& P,
                           cudaStream_t stream) {
    static bool attr_set[32] = {false};
    const int slot = current_device_slot();
    if (!attr_set[slot]) {
        GDT_CUDA(cudaFuncSetAttribute(score_filter_kernel<CLUSTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_set[slot] = true;
    }
    const int units = P.n_items < n_units ? P.n_items : n_units;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(units * CLUSTER), 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CLUSTER > 1 ? 1 : 0;
    GDT_CUDA(cudaLaunchKernelEx(&cfg, score_filter_kernel<CLUSTER>, map_q, map_db, P));
    return GDT_OK;
}

static int launch_filter(int cluster, int n_units, const CUtensorMap& map_q, const CUtensorMap& map_db, const FilterParams& P,
                         cudaStream_t stream) {
    if (cluster == 4) return launch_filter_t<4>(n_units, map_q, map_db, P, stream);
    if (cluster == 2) return launch_filter_t<2>(n_units, map_q, map_db, P, stream);
    return launch_filter_t<1>(n_units, map_q, map_db, P, stream);
}

static int topk_check(const void* q, int nq, long long ndb, int d, int k, long long index_base, const void* ws,
                      size_t ws_bytes) {
    if (!q || !ws) return GDT_ERR_INVALID_ARGUMENT;
    if (nq <= 0 || ndb <= 0 || d <= 0 || k <= 0) return GDT_ERR_INVALID_ARGUMENT;
    if ((d & 7) != 0 || d > 8192 || k > 1024 || ndb > 0x7fffffffLL || index_base < 0 ||
        index_base + ndb > 0xffffffffLL)
        return GDT_ERR_UNSUPPORTED;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return GDT_ERR_NO_DEVICE; }
    if (ws_bytes < gdt_score_topk_workspace_bytes(nq, ndb, d, k) || (((uintptr_t)ws) & 255)) return GDT_ERR_WORKSPACE_TOO_SMALL;
    return GDT_OK;
}

extern "C" int gdt_score_topk_exchange_layout(int nq, long long ndb, int d, int k, size_t* hist_offset, size_t* hist_bytes) {
    if (nq <= 0 || ndb <= 0 || d <= 0 || k <= 0 || !hist_offset || !hist_bytes) return GDT_ERR_INVALID_ARGUMENT;
    const TopkPlan L = topk_plan(nq, ndb, d, k);
    *hist_offset = L.hist;
    *hist_bytes = (size_t)nq * kHistBins * 4;
    return GDT_OK;
}

static int score_topk_filter_impl(const float* q, const void* db_f16, const float* db_stats, int nq, long long ndb,
                                  int d, int k, int32_t* status_dev, void* ws, size_t ws_bytes, float* dump,
                                  void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!db_f16 || !db_stats || !status_dev) return GDT_ERR_INVALID_ARGUMENT;
    if ((((uintptr_t)db_f16) & 15) != 0) return GDT_ERR_INVALID_ARGUMENT;
    int rc = topk_check(q, nq, ndb, d, k, 0, ws, ws_bytes);
    if (rc != GDT_OK) return rc;

    const TopkPlan L = topk_plan(nq, ndb, d, k);
    char* base = (char*)ws;
    __half* qb = (__half*)(base + L.qb);
    QMeta* meta = (QMeta*)(base + L.meta);
    uint32_t* tau = (uint32_t*)(base + L.tau);
    uint32_t* cnt = (uint32_t*)(base + L.cnt);
    uint32_t* hist = (uint32_t*)(base + L.hist);
    uint64_t* cand = (uint64_t*)(base + L.cand);

    q_prepare_kernel<<<ceil_div(nq, 8), 256, 0, stream>>>(q, nq, d, db_stats, qb, meta, tau, cnt, L.n_segs, hist, status_dev);
    GDT_LAUNCH_CHECK();

    CUtensorMap map_q, map_db;
    rc = make_f16_map(&map_q, qb, nq, d, kBlockM);
    if (rc != GDT_OK) return rc;
    rc = make_f16_map(&map_db, db_f16, ndb, d, kBlockN / L.cluster);     // one multicast slice per CTA of a cluster
    if (rc != GDT_OK) return rc;

    const int sms = sm_count_current_device();
    FilterParams P;
    P.nq = nq; P.d = d; P.k = k; P.ndb = ndb;
    P.n_qtiles = L.n_qtiles; P.n_qgroups = L.n_qgroups;
    P.n_kblocks = ceil_div(d, kBlockK);
    P.n_segs = L.n_segs; P.n_seed = L.n_seed; P.cap0 = L.cap0; P.cap1 = L.cap1;
    P.meta = meta; P.tau = tau; P.cnt = cnt; P.hist = hist; P.cand = cand;
    P.dump = dump; P.dump_ld = ndb;
    (void)sms;
    // seed pass: the first tiles of the shard against every query tile establish the thresholds
    P.tile_begin = 0; P.tile_end = L.seed_tiles;
    P.n_stripes = L.n_seed; P.stripe_len = L.seed_len; P.n_items = L.n_seed * L.n_qgroups; P.seg_first = 0;
    rc = launch_filter(L.cluster, L.n_units, map_q, map_db, P, stream);
    if (rc != GDT_OK) return rc;
    if (L.n_stripes > 0) {
        P.tile_begin = L.seed_tiles; P.tile_end = L.n_dtiles;
        P.n_stripes = L.n_stripes; P.stripe_len = L.stripe_len; P.n_items = L.n_stripes * L.n_qgroups; P.seg_first = L.n_seed;
        rc = launch_filter(L.cluster, L.n_units, map_q, map_db, P, stream);
        if (rc != GDT_OK) return rc;
    }
    return GDT_OK;
}

extern "C" int gdt_score_topk_filter(const float* q, const void* db_f16, const float* db_stats, int nq, long long ndb,
                                     int d, int k, int32_t* status_dev, void* ws, size_t ws_bytes, void* stream) {
    return score_topk_filter_impl(q, db_f16, db_stats, nq, ndb, d, k, status_dev, ws, ws_bytes, nullptr, stream);
}

// Debug / test hook: the filter pass with every raw tensor-core score written to coarse[nq][ndb], plus the per-query
// bound the filter relies on: meta[q] = {scale, 1/scale, margin = 2 * E_q, sq}. tests/ compare
// |coarse - sq * sx * <q, x>| (exact, fp64) with E_q = margin / 2.
extern "C" int gdt_debug_k3_coarse_scores(const float* q, const void* db_f16, const float* db_stats, int nq, long long ndb,
                                          int d, int k, float* coarse, float* meta_out, int32_t* status_dev, void* ws,
                                          size_t ws_bytes, void* stream) {
    if (!coarse || !meta_out) return GDT_ERR_INVALID_ARGUMENT;
    int rc = score_topk_filter_impl(q, db_f16, db_stats, nq, ndb, d, k, status_dev, ws, ws_bytes, coarse, stream);
    if (rc != GDT_OK) return rc;
    const TopkPlan L = topk_plan(nq, ndb, d, k);
    GDT_CUDA(cudaMemcpyAsync(meta_out, (const char*)ws + L.meta, (size_t)nq * sizeof(QMeta), cudaMemcpyDeviceToDevice,
                             (cudaStream_t)stream));
    return GDT_OK;
}

extern "C" int gdt_score_topk_finalize(const float* q, const float* db, int nq, long long ndb, int d, int k,
                                       long long index_base, float* top_scores, int64_t* top_idx, int32_t* status_dev,
                                       void* ws, size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!db || !top_scores || !top_idx || !status_dev) return GDT_ERR_INVALID_ARGUMENT;
    int rc = topk_check(q, nq, ndb, d, k, index_base, ws, ws_bytes);
    if (rc != GDT_OK) return rc;
    const TopkPlan L = topk_plan(nq, ndb, d, k);
    char* base = (char*)ws;
    const int scap = survivor_capacity(k);
    const int dpad = (d + 3) & ~3;
    const size_t fsmem = (size_t)scap * 8 + (size_t)dpad * 4;
    static size_t fattr[32] = {0};
    const int slot = current_device_slot();
    if (fsmem > 48 * 1024 && fsmem > fattr[slot]) {
        GDT_CUDA(cudaFuncSetAttribute(topk_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
        fattr[slot] = fsmem;
    }
    topk_finalize_kernel<<<nq, 256, fsmem, stream>>>(q, db, d, dpad, k, L.n_segs, L.n_seed, L.cap0, L.cap1, scap, index_base,
                                                     (const QMeta*)(base + L.meta), (const uint32_t*)(base + L.cnt),
                                                     (const uint32_t*)(base + L.hist), (const uint64_t*)(base + L.cand),
                                                     top_scores, top_idx, status_dev);
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}

extern "C" int gdt_score_topk(const float* q, const float* db, const void* db_f16, const float* db_stats, int nq,
                              long long ndb, int d, int k, long long index_base, float* top_scores, int64_t* top_idx,
                              int32_t* status_dev, void* ws, size_t ws_bytes, void* stream) {
    int rc = gdt_score_topk_filter(q, db_f16, db_stats, nq, ndb, d, k, status_dev, ws, ws_bytes, stream);
    if (rc != GDT_OK) return rc;
    return gdt_score_topk_finalize(q, db, nq, ndb, d, k, index_base, top_scores, top_idx, status_dev, ws, ws_bytes, stream);
}
