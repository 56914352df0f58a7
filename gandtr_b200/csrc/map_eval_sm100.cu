// K4 -- ranks of ground-truth ids and mAP on the GPU, for sm_100a.
// Replaces the consumers of the full `ranks = np.argsort(-scores, axis=0)` matrix
// (mdir/components/optim/score/cirscore.py:72-73) inside compute_ap / compute_map
// (mdir/external/cirtorch/utils/evaluate.py:3-37,39-111): only the positions of a query's positive and
// junk ids are ever read from `ranks`, so we compute exactly those positions -- the number of database
// rows that sort before each probe under the library's total order (score desc, index asc) -- and
// never materialise the ndb x nq rank matrix.
//
//   probe_scores_kernel  exact score of every (query, probe id) pair owned by this shard
//   rank_counts_kernel   8 queries per CTA in shared memory, one warp per database row (read once),
//                        exact scores compared against the probes of each query; per-CTA shared-memory
//                        counters, one global atomicAdd per (query, probe) per CTA
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

__global__ void __launch_bounds__(256)
rank_counts_kernel(const float* __restrict__ q, const float* __restrict__ db, int nq, long long ndb, int d, int dpad,
                   long long index_base, const int64_t* __restrict__ probe_idx, const float* __restrict__ probe_score,
                   int pmax, unsigned long long* __restrict__ before, int rows_per_cta) {
    extern __shared__ __align__(16) uint8_t rsm[];
    float* qs = (float*)rsm;                                              // [kRankQB][dpad]
    float* ps = qs + (size_t)kRankQB * dpad;                              // [kRankQB][pmax]
    long long* pi = (long long*)(ps + (size_t)kRankQB * ((pmax + 1) & ~1));  // [kRankQB][pmax]
    uint32_t* cnt = (uint32_t*)(pi + (size_t)kRankQB * pmax);             // [kRankQB][pmax]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qb0 = blockIdx.y * kRankQB;
    const int nqb = min(kRankQB, nq - qb0);
    for (int i = tid; i < kRankQB * dpad; i += 256) {
        const int j = i / dpad, c = i - j * dpad;
        qs[i] = (j < nqb && c < d) ? q[(size_t)(qb0 + j) * d + c] : 0.0f;
    }
    for (int i = tid; i < kRankQB * pmax; i += 256) {
        const int j = i / pmax, p = i - j * pmax;
        long long id = -1;
        float s = 0.f;
        if (j < nqb) {
            id = probe_idx[(size_t)(qb0 + j) * pmax + p];
            s = probe_score[(size_t)(qb0 + j) * pmax + p];
        }
        pi[i] = id;
        ps[i] = s;
        cnt[i] = 0;
    }
    __syncthreads();
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    const long long r1 = min(r0 + (long long)rows_per_cta, ndb);
    for (long long r = r0 + warp; r < r1; r += 8) {
        float out[kRankQB];
        warp_exact_dot_multi<kRankQB>(qs, dpad, db + (size_t)r * d, d, lane, out);
        const long long gid = index_base + r;
#pragma unroll
        for (int j = 0; j < kRankQB; ++j) {
            if (j < nqb) {
                const float s = out[j] + 0.0f;
                for (int p = lane; p < pmax; p += 32) {
                    const long long id = pi[j * pmax + p];
                    if (id < 0) continue;
                    const float t = ps[j * pmax + p] + 0.0f;
                    if (s > t || (s == t && gid < id)) atomicAdd(&cnt[j * pmax + p], 1u);
                }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < nqb * pmax; i += 256) {
        const uint32_t c = cnt[i];
        if (c) atomicAdd(before + (size_t)qb0 * pmax + i, (unsigned long long)c);
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

extern "C" int gdt_rank_counts(const float* q, const float* db, int nq, long long ndb, int d, long long index_base,
                               const int64_t* probe_idx, const float* probe_score, int pmax, int64_t* before,
                               void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!q || !db || !probe_idx || !probe_score || !before) return GDT_ERR_INVALID_ARGUMENT;
    if (nq <= 0 || ndb <= 0 || d <= 0 || pmax <= 0) return GDT_ERR_INVALID_ARGUMENT;
    if (!have_device_k4()) return GDT_ERR_NO_DEVICE;
    const int dpad = (d + 3) & ~3;
    const size_t smem = (size_t)kRankQB * dpad * 4 + (size_t)kRankQB * ((pmax + 1) & ~1) * 4 + (size_t)kRankQB * pmax * 8 +
                        (size_t)kRankQB * pmax * 4;
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
    rows = (rows + 7) / 8 * 8;
    dim3 grid((unsigned)ceil_div_ll(ndb, rows), (unsigned)qblocks);
    rank_counts_kernel<<<grid, 256, smem, stream>>>(q, db, nq, ndb, d, dpad, index_base, probe_idx, probe_score, pmax,
                                                     (unsigned long long*)before, (int)rows);
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
