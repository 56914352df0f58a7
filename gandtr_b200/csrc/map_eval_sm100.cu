// K4 -- ranks of ground-truth ids and mAP on the GPU, for sm_100a.
// Replaces the consumers of the full `ranks = np.argsort(-scores, axis=0)` matrix
// (mdir/components/optim/score/cirscore.py:72-73) inside compute_ap / compute_map
// (mdir/external/cirtorch/utils/evaluate.py:3-37,39-111): only the positions of a query's positive and
// junk ids are ever read from `ranks`, so we compute exactly those positions -- the number of database
// rows that sort before each probe under the library's total order (score desc, index asc) -- and
// never materialise the ndb x nq rank matrix.
//
//   probe_scores_kernel  exact score of every (query, probe id) pair owned by this shard
//   probe_sort_kernel    per query: the probes' rank keys sorted once (descending)
//   rank_counts_kernel   8 queries per CTA in shared memory, one warp per database row (read once): exact scores,
//                        then ONE binary search per (row, query) into the sorted probe keys and ONE bucket increment
//                        ("this row sorts before the probes from slot m on") -- O(log pmax) instead of pmax compares and
//                        up to pmax shared-memory atomics per (row, query)
//   rank_finish_kernel   per query: prefix sum over the buckets, scattered back to the probes' columns
//   map_eval_kernel      junk shift + trapezoidal AP + precision@k in fp64, same operation order as
//                        the reference so the per-query numbers are bit-identical
#include "common.cuh"
#include "select.cuh"

namespace gdt {

constexpr int kRankQB = 8;

__global__ void __launch_bounds__(256)
probe_scores_kernel(const float* __restrict__ q, const float* __restrict__ db, int d, int dpad, long long ndb,
                    long long index_base, const int64_t* __restrict__ probe_idx, int pmax, float* __restrict__ probe_score) {
    extern __shared__ __align__(16) float qrow[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int qi = blockIdx.x;
    for (int i = tid; i < dpad; i += 256) qrow[i] = i < d ? q[(size_t)qi * d + i] : 0.f;
    __syncthreads();
    for (int p = wid; p < pmax; p += 8) {
        const long long id = probe_idx[(size_t)qi * pmax + p];
        const long long local = id - index_base;
        if (id < 0 || local < 0 || local >= ndb) continue;
        const float s = warp_exact_dot(qrow, db + (size_t)local * d, d, lane);
        if (lane == 0) probe_score[(size_t)qi * pmax + p] = s;
    }
}

// rank key of a probe: 0 for padding / ids this library cannot address
__device__ __forceinline__ uint64_t probe_key(long long id, float score) {
    return (id >= 0 && id <= 0xffffffffLL) ? rank_key(score, (uint32_t)id) : 0ull;
}

__global__ void __launch_bounds__(256)
probe_sort_kernel(const int64_t* __restrict__ probe_idx, const float* __restrict__ probe_score, int pmax, int pp,
                  uint64_t* __restrict__ skeys, unsigned long long* __restrict__ gbucket) {
    extern __shared__ __align__(16) uint64_t keys[];   // [pp]
    const int tid = threadIdx.x, qi = blockIdx.x;
    for (int i = tid; i < pp; i += 256)
        keys[i] = i < pmax ? probe_key(probe_idx[(size_t)qi * pmax + i], probe_score[(size_t)qi * pmax + i]) : 0ull;
    block_bitonic_sort_desc(keys, pp, tid, 256);
    for (int i = tid; i < pp; i += 256) skeys[(size_t)qi * pp + i] = keys[i];
    for (int i = tid; i <= pp; i += 256) gbucket[(size_t)qi * (pp + 1) + i] = 0ull;
}

// number of sorted (descending) keys that are >= key: the row sorts before the probes in the slots from there on
__device__ __forceinline__ int slots_not_after(const uint64_t* sk, int pp, uint64_t key) {
    int lo = 0, n = pp;                       // first index with sk[i] < key
    while (n > 0) {
        const int half = n >> 1;
        if (sk[lo + half] >= key) { lo += half + 1; n -= half + 1; }
        else n = half;
    }
    return lo;
}

// Rows per warp and iteration: the query float4 read from shared memory is reused for 4 database rows (16 FMAs per
// 16-byte shared-memory load), which balances the FMA pipe against the shared-memory bandwidth.
constexpr int kRankRows = 4;

// Scores are computed in fp32 (FMA, 128 lanes / clk / SM -- the fp64 pipe of this part is >10x slower) together with a
// rigorous bound on their distance to the library's exact score (fp64-accumulated, rounded once):
//     |s32 - s_exact| <= tol = 1.1 * (d / 32 + 8) * 2^-24 * ||q|| * ||x||
// (every lane adds d / 32 products with FMAs, five butterfly levels follow: a summation tree of depth h = d / 32 + 5 has
// error <= h u / (1 - h u) * sum |q_i x_i| <= ... * ||q|| ||x|| by Cauchy-Schwarz; + 1 for the exact score's own rounding
// to fp32; the factor 1.1 and the + 2 cover 1 / (1 - h u) and the fp32 norms.)
// A row is ranked from s32 alone when no probe key lies between the keys of (s32 - tol) and (s32 + tol); only the rare
// rows that come that close to a probe are re-scored exactly (warp_exact_dot), so the counts are those of the exact
// kernel. One lane per (row, query) pair does the binary search: all 32 lanes of a warp are busy.
__global__ void __launch_bounds__(256)
rank_counts_kernel(const float* __restrict__ q, const float* __restrict__ db, int nq, long long ndb, int d, int dpad,
                   long long index_base, const uint64_t* __restrict__ skeys, int pp,
                   unsigned long long* __restrict__ gbucket, int rows_per_cta, int force_exact, int keys_in_smem) {
    extern __shared__ __align__(16) uint8_t rsm[];
    float* qs = (float*)rsm;                                              // [kRankQB][dpad]
    uint64_t* sk_s = (uint64_t*)(qs + (size_t)kRankQB * dpad);            // [kRankQB][pp] sorted probe keys (keys_in_smem)
    uint32_t* bucket = (uint32_t*)(sk_s + (keys_in_smem ? (size_t)kRankQB * pp : 0));   // [kRankQB][pp + 1]
    __shared__ float qnorm[kRankQB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qb0 = blockIdx.y * kRankQB;
    const int nqb = min(kRankQB, nq - qb0);
    for (int i = tid; i < kRankQB * dpad; i += 256) {
        const int j = i / dpad, c = i - j * dpad;
        qs[i] = (j < nqb && c < d) ? q[(size_t)(qb0 + j) * d + c] : 0.0f;
    }
    for (int i = tid; i < kRankQB * (pp + 1); i += 256) bucket[i] = 0;
    if (keys_in_smem)      // the binary searches are chains of dependent loads: from shared memory, not from L2
        for (int i = tid; i < kRankQB * pp; i += 256) sk_s[i] = i < nqb * pp ? skeys[(size_t)qb0 * pp + i] : 0ull;
    __syncthreads();
    {   // ||q_j||, one warp per query
        float ss = 0.f;
        for (int i = lane; i < dpad; i += 32) { const float v = qs[warp * dpad + i]; ss = fmaf(v, v, ss); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) qnorm[warp] = sqrtf(ss) * 1.0001f;
    }
    __syncthreads();
    const float cd = 1.1f * (float)((d + 31) / 32 + 8) * 5.9604644775390625e-08f;
    const bool vec = (d & 3) == 0 && (((uintptr_t)db) & 15) == 0;
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    const long long r1 = min(r0 + (long long)rows_per_cta, ndb);
    const int my_row = lane >> 3, my_q = lane & 7;                          // the (row, query) pair this lane ranks
    for (long long rb = r0 + warp * kRankRows; rb < r1; rb += 8 * kRankRows) {
        float acc[kRankRows * kRankQB];
        float xx[kRankRows];
#pragma unroll
        for (int a = 0; a < kRankRows * kRankQB; ++a) acc[a] = 0.f;
#pragma unroll
        for (int r = 0; r < kRankRows; ++r) xx[r] = 0.f;
        const float* xrow[kRankRows];
#pragma unroll
        for (int r = 0; r < kRankRows; ++r) xrow[r] = db + (size_t)min(rb + r, r1 - 1) * d;   // ragged tail: clamped
        if (vec) {
            const int nv = d >> 2;
            float4 xnext[kRankRows];
#pragma unroll
            for (int r = 0; r < kRankRows; ++r) xnext[r] = lane < nv ? __ldg((const float4*)xrow[r] + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = lane; i < nv; i += 32) {
                float4 x[kRankRows];
#pragma unroll
                for (int r = 0; r < kRankRows; ++r) x[r] = xnext[r];
                if (i + 32 < nv) {                      // next step's rows are in flight while this step's FMAs issue
#pragma unroll
                    for (int r = 0; r < kRankRows; ++r) xnext[r] = __ldg((const float4*)xrow[r] + i + 32);
                }
#pragma unroll
                for (int r = 0; r < kRankRows; ++r)
                    xx[r] = fmaf(x[r].x, x[r].x, fmaf(x[r].y, x[r].y, fmaf(x[r].z, x[r].z, fmaf(x[r].w, x[r].w, xx[r]))));
#pragma unroll
                for (int j = 0; j < kRankQB; ++j) {
                    const float4 a = *(const float4*)(qs + (size_t)j * dpad + (i << 2));
#pragma unroll
                    for (int r = 0; r < kRankRows; ++r)
                        acc[r * kRankQB + j] = fmaf(a.x, x[r].x, fmaf(a.y, x[r].y, fmaf(a.z, x[r].z, fmaf(a.w, x[r].w, acc[r * kRankQB + j]))));
                }
            }
        } else {
            for (int i = lane; i < d; i += 32) {
#pragma unroll
                for (int r = 0; r < kRankRows; ++r) {
                    const float x = __ldg(xrow[r] + i);
                    xx[r] = fmaf(x, x, xx[r]);
#pragma unroll
                    for (int j = 0; j < kRankQB; ++j) acc[r * kRankQB + j] = fmaf(qs[(size_t)j * dpad + i], x, acc[r * kRankQB + j]);
                }
            }
        }
        // transposing butterfly: 32 partial sums per lane -> lane L holds the total of pair L (31 shuffles instead of 160)
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            const bool upper = (lane & step) != 0;
#pragma unroll
            for (int i = 0; i < step; ++i) {
                const float send = upper ? acc[i] : acc[i + step];
                const float keep = upper ? acc[i + step] : acc[i];
                acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
            }
        }
        float xn = 0.f;
#pragma unroll
        for (int r = 0; r < kRankRows; ++r) {
            float v = xx[r];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (r == my_row) xn = v;
        }
        const long long row = rb + my_row;
        const bool live = row < r1 && my_q < nqb;
        const uint64_t* sk = keys_in_smem ? sk_s + (size_t)my_q * pp : skeys + (size_t)(qb0 + (my_q < nqb ? my_q : 0)) * pp;
        const uint32_t gid = (uint32_t)(index_base + row);
        float s = acc[0];
        int m = 0;
        bool need_exact = false;
        if (live) {
            const float tol = cd * qnorm[my_q] * (sqrtf(xn) * 1.0001f);
            // keys of the interval ends: (s + tol, best index) is the largest key the row can have, (s - tol, worst
            // index) the smallest; m(K) = number of probe keys >= K is non-increasing in K
            // ONE search: m_lo counts the probe keys >= the largest key the row can have; the row is ambiguous iff the
            // next probe key (the largest one below that) is still >= the smallest key the row can have
            const int m_lo = slots_not_after(sk, pp, rank_key(__fadd_ru(s, tol), 0u));
            const uint64_t k_min = rank_key(__fadd_rd(s, -tol), 0xffffffffu);
            m = m_lo;
            need_exact = (m_lo < pp && sk[m_lo] >= k_min) || force_exact || !(tol >= 0.f);
        }
        unsigned todo = __ballot_sync(0xffffffffu, need_exact);
        while (todo) {                                    // rare: the whole warp re-scores one pair exactly
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            const long long er = rb + (l >> 3);
            const float es = warp_exact_dot(qs + (size_t)(l & 7) * dpad, db + (size_t)er * d, d, lane);
            if (lane == l) m = slots_not_after(sk, pp, rank_key(es, gid));
        }
        if (live) atomicAdd(&bucket[my_q * (pp + 1) + m], 1u);
    }
    __syncthreads();
    for (int i = tid; i < nqb * (pp + 1); i += 256) {
        const uint32_t c = bucket[i];
        if (c) atomicAdd(gbucket + (size_t)qb0 * (pp + 1) + i, (unsigned long long)c);
    }
}

// before[q][p] += rows of this shard that sort before probe p = inclusive prefix sum of the buckets up to the probe's slot
__global__ void __launch_bounds__(256)
rank_finish_kernel(const int64_t* __restrict__ probe_idx, const float* __restrict__ probe_score, int pmax, int pp,
                   const uint64_t* __restrict__ skeys, const unsigned long long* __restrict__ gbucket,
                   unsigned long long* __restrict__ before) {
    extern __shared__ __align__(16) unsigned long long pre[];   // [pp + 1] inclusive prefix sums
    __shared__ unsigned long long carry;
    const int tid = threadIdx.x, qi = blockIdx.x;
    if (tid == 0) carry = 0ull;
    __syncthreads();
    // chunked scan: 256 buckets at a time (pp + 1 <= 2049)
    for (int base = 0; base <= pp; base += 256) {
        const int i = base + tid;
        unsigned long long v = i <= pp ? gbucket[(size_t)qi * (pp + 1) + i] : 0ull;
        const int lane = tid & 31, wid = tid >> 5;
        __shared__ unsigned long long wtot[8];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) wtot[wid] = v;
        __syncthreads();
        unsigned long long add = carry;
        for (int w = 0; w < wid; ++w) add += wtot[w];
        if (i <= pp) pre[i] = v + add;
        __syncthreads();
        if (tid == 255) carry = v + add;
        __syncthreads();
    }
    for (int p = tid; p < pmax; p += 256) {
        const uint64_t key = probe_key(probe_idx[(size_t)qi * pmax + p], probe_score[(size_t)qi * pmax + p]);
        if (key == 0ull) continue;
        // slot of the probe = number of sorted keys strictly greater than its own
        int lo = 0, n = pp;
        const uint64_t* sk = skeys + (size_t)qi * pp;
        while (n > 0) {
            const int half = n >> 1;
            if (sk[lo + half] > key) { lo += half + 1; n -= half + 1; }
            else n = half;
        }
        before[(size_t)qi * pmax + p] += pre[lo];
    }
}

// one CTA per query; sorts the two rank lists in shared memory, thread 0 runs the reference's loops
__global__ void __launch_bounds__(128)
map_eval_kernel(const int64_t* __restrict__ pos_rank, int pmax_pos, const int64_t* __restrict__ junk_rank, int pmax_junk,
                const int32_t* __restrict__ npos, const int32_t* __restrict__ njunk, const int32_t* __restrict__ nres,
                const int32_t* __restrict__ kappas, int nk, int np_pos, int np_junk, double* __restrict__ ap,
                double* __restrict__ prk) {
    extern __shared__ __align__(16) uint64_t lists[];  // [np_pos] ascending positives, then [np_junk] ascending junk
    uint64_t* pos = lists;
    uint64_t* junk = lists + np_pos;
    const int tid = threadIdx.x, qi = blockIdx.x;
    const int n_p = npos[qi], n_j = njunk[qi];
    // evaluate.py:75-80: positions come from the ids FOUND in the ranking (np.in1d: set semantics), the recall step from
    // the length of the ground-truth list as given (nres = len(qgnd)); the two differ for duplicated / foreign ids
    const int n_res = nres ? nres[qi] : n_p;
    // descending bitonic sort on (~rank) == ascending on rank; padding key 0 sorts last
    for (int i = tid; i < np_pos; i += 128) pos[i] = i < n_p ? ~(uint64_t)pos_rank[(size_t)qi * pmax_pos + i] : 0ull;
    for (int i = tid; i < np_junk; i += 128) junk[i] = i < n_j ? ~(uint64_t)junk_rank[(size_t)qi * pmax_junk + i] : 0ull;
    block_bitonic_sort_desc(pos, np_pos, tid, 128);
    block_bitonic_sort_desc(junk, np_junk, tid, 128);
    if (tid != 0) return;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (n_res == 0 || n_p == 0) {  // evaluate.py:68-72 (empty ground truth); no id found: nothing to average either
        ap[qi] = nan;
        for (int j = 0; j < nk; ++j) prk[(size_t)qi * nk + j] = nan;
        return;
    }
    // junk shift (evaluate.py:83-94) fused with compute_ap (evaluate.py:20-35)
    long long kshift = 0;
    int ij = 0;
    double acc = 0.0;
    const double recall_step = 1.0 / (double)n_res;
    long long maxpos = 0;
    for (int ip = 0; ip < n_p; ++ip) {
        long long r = (long long)(~pos[ip]);
        while (ij < n_j && r > (long long)(~junk[ij])) { ++kshift; ++ij; }
        r -= kshift;
        pos[ip] = (uint64_t)r;  // keep the shifted rank for precision@k
        const double precision_0 = r == 0 ? 1.0 : (double)ip / (double)r;
        const double precision_1 = (double)(ip + 1) / (double)(r + 1);
        acc += (precision_0 + precision_1) * recall_step / 2.0;
        if (r + 1 > maxpos) maxpos = r + 1;
    }
    ap[qi] = acc;
    // precision@k on 1-based positions (evaluate.py:101-105)
    for (int j = 0; j < nk; ++j) {
        const long long kq = maxpos < (long long)kappas[j] ? maxpos : (long long)kappas[j];
        long long hits = 0;
        for (int ip = 0; ip < n_p; ++ip) hits += ((long long)pos[ip] + 1 <= kq) ? 1 : 0;
        prk[(size_t)qi * nk + j] = (double)hits / (double)kq;
    }
}

}  // namespace gdt

using namespace gdt;

static bool have_device_k4() {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return false;
    }
    return true;
}

extern "C" int gdt_probe_scores(const float* q, const float* db, int nq, long long ndb, int d, long long index_base,
                                const int64_t* probe_idx, int pmax, float* probe_score, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!q || !db || !probe_idx || !probe_score) return GDT_ERR_INVALID_ARGUMENT;
    if (nq <= 0 || ndb <= 0 || d <= 0 || pmax <= 0) return GDT_ERR_INVALID_ARGUMENT;
    if (!have_device_k4()) return GDT_ERR_NO_DEVICE;
    const int dpad = (d + 3) & ~3;
    const size_t smem = (size_t)dpad * 4;
    if (smem > 48 * 1024) return GDT_ERR_UNSUPPORTED;
    probe_scores_kernel<<<nq, 256, smem, stream>>>(q, db, d, dpad, ndb, index_base, probe_idx, pmax, probe_score);
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}

static int g_k4_force_exact = 0;     // debug (gdt_debug_k4_exact): 1 = every (row, query) pair is re-scored exactly

extern "C" int gdt_debug_k4_exact(int on) {
    g_k4_force_exact = on ? 1 : 0;
    return GDT_OK;
}

extern "C" size_t gdt_rank_counts_workspace_bytes(int nq, int pmax) {
    if (nq <= 0 || pmax <= 0) return 0;
    const size_t pp = (size_t)next_pow2(pmax);
    return align_up((size_t)nq * pp * 8, 256) + align_up((size_t)nq * (pp + 1) * 8, 256) + 256;
}

extern "C" int gdt_rank_counts(const float* q, const float* db, int nq, long long ndb, int d, long long index_base,
                               const int64_t* probe_idx, const float* probe_score, int pmax, int64_t* before, void* ws,
                               size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!q || !db || !probe_idx || !probe_score || !before || !ws) return GDT_ERR_INVALID_ARGUMENT;
    if (nq <= 0 || ndb <= 0 || d <= 0 || pmax <= 0) return GDT_ERR_INVALID_ARGUMENT;
    if (pmax > 2048 || index_base < 0 || index_base + ndb > 0xffffffffLL) return GDT_ERR_UNSUPPORTED;
    if (!have_device_k4()) return GDT_ERR_NO_DEVICE;
    if (ws_bytes < gdt_rank_counts_workspace_bytes(nq, pmax) || (((uintptr_t)ws) & 255)) return GDT_ERR_WORKSPACE_TOO_SMALL;
    const int pp = next_pow2(pmax);
    Workspace W(ws, ws_bytes);
    uint64_t* skeys = W.take<uint64_t>((size_t)nq * pp);
    unsigned long long* gbucket = W.take<unsigned long long>((size_t)nq * (pp + 1));
    if (!W.ok()) return GDT_ERR_WORKSPACE_TOO_SMALL;
    probe_sort_kernel<<<nq, 256, (size_t)pp * 8, stream>>>(probe_idx, probe_score, pmax, pp, skeys, gbucket);
    GDT_LAUNCH_CHECK();
    const int dpad = (d + 3) & ~3;
    size_t smem = (size_t)kRankQB * dpad * 4 + (size_t)kRankQB * (pp + 1) * 4;
    const int keys_in_smem = smem + (size_t)kRankQB * pp * 8 <= 96 * 1024;
    if (keys_in_smem) smem += (size_t)kRankQB * pp * 8;
    if (smem > 200 * 1024) return GDT_ERR_UNSUPPORTED;
    static size_t attr_bytes_dev[32] = {0};
    size_t& attr_bytes = attr_bytes_dev[current_device_slot()];
    if (smem > 48 * 1024 && smem > attr_bytes) {
        GDT_CUDA(cudaFuncSetAttribute(rank_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_bytes = smem;
    }
    const int qblocks = ceil_div(nq, kRankQB);
    const int sms = sm_count_current_device();
    long long want = ceil_div_ll(4LL * sms, qblocks);
    if (want < 1) want = 1;
    long long rows = ceil_div_ll(ndb, want);
    if (rows < 256) rows = 256;
    rows = (rows + 31) / 32 * 32;
    dim3 grid((unsigned)ceil_div_ll(ndb, rows), (unsigned)qblocks);
    rank_counts_kernel<<<grid, 256, smem, stream>>>(q, db, nq, ndb, d, dpad, index_base, skeys, pp, gbucket, (int)rows,
                                                     g_k4_force_exact, keys_in_smem);
    GDT_LAUNCH_CHECK();
    rank_finish_kernel<<<nq, 256, (size_t)(pp + 1) * 8, stream>>>(probe_idx, probe_score, pmax, pp, skeys, gbucket,
                                                                  (unsigned long long*)before);
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}

extern "C" int gdt_map_eval(const int64_t* pos_rank, int pmax_pos, const int64_t* junk_rank, int pmax_junk,
                            const int32_t* npos, const int32_t* njunk, const int32_t* nres, int nq, const int32_t* kappas,
                            int nk, double* ap, double* prk, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!pos_rank || !junk_rank || !npos || !njunk || !ap) return GDT_ERR_INVALID_ARGUMENT;
    if (nq <= 0 || pmax_pos <= 0 || pmax_junk <= 0 || nk < 0 || (nk > 0 && (!kappas || !prk))) return GDT_ERR_INVALID_ARGUMENT;
    if (!have_device_k4()) return GDT_ERR_NO_DEVICE;
    const int np_pos = next_pow2(pmax_pos), np_junk = next_pow2(pmax_junk);
    const size_t smem = (size_t)(np_pos + np_junk) * 8;
    if (smem > 200 * 1024) return GDT_ERR_UNSUPPORTED;
    static size_t attr_bytes_dev[32] = {0};
    size_t& attr_bytes = attr_bytes_dev[current_device_slot()];
    if (smem > 48 * 1024 && smem > attr_bytes) {
        GDT_CUDA(cudaFuncSetAttribute(map_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_bytes = smem;
    }
    map_eval_kernel<<<nq, 128, smem, stream>>>(pos_rank, pmax_pos, junk_rank, pmax_junk, npos, njunk, nres, kappas, nk,
                                               np_pos, np_junk, ap, prk);
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}
