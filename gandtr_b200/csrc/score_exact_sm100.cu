// K3 (exact variant) -- query x database scoring + top-k on CUDA cores, fp64-accumulated.
// Replaces `scores = np.dot(vecs.T, qvecs); ranks = np.argsort(-scores, axis=0)[:k]`
// (mdir/components/optim/score/cirscore.py:71-72) for small problems, serves as the on-device
// cross-check of the tcgen05 path (score_topk_sm100.cu) and as its overflow fallback.
//
//   exact_scores_kernel  8 queries staged in shared memory per CTA, one warp per database row: the row
//                        is read once (16-byte coalesced), 8 fp64 dot products, rounded once to fp32.
//   select_rows_kernel   one CTA per query: 3-pass radix select (11+11+10 bits of the order-preserving
//                        key) for the k-th score, ordered compaction of ties (lower index first),
//                        bitonic sort of the k winners.
//   topk_merge_kernel    one CTA per query: merges the per-shard lists gathered by the NCCL allgather.
#include "common.cuh"
#include "select.cuh"

namespace gdt {

constexpr int kExactQB = 8;

__global__ void __launch_bounds__(256)
exact_scores_kernel(const float* __restrict__ q, const float* __restrict__ db, int nq, long long ndb, int d, int dpad,
                    float* __restrict__ scores, long long ld, int rows_per_cta) {
    extern __shared__ __align__(16) float qs[];  // [kExactQB][dpad]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qb0 = blockIdx.y * kExactQB;
    const int nqb = min(kExactQB, nq - qb0);
    for (int i = tid; i < kExactQB * dpad; i += 256) {
        const int j = i / dpad, c = i - j * dpad;
        qs[i] = (j < nqb && c < d) ? q[(size_t)(qb0 + j) * d + c] : 0.0f;
    }
    __syncthreads();
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    const long long r1 = min(r0 + (long long)rows_per_cta, ndb);
    for (long long r = r0 + warp; r < r1; r += 8) {
        float out[kExactQB];
        warp_exact_dot_multi<kExactQB>(qs, dpad, db + (size_t)r * d, d, lane, out);
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < kExactQB; ++j)
            if (lane == j) v = out[j];
        if (lane < nqb) scores[(size_t)(qb0 + lane) * ld + r] = v;
    }
}

// ---- radix select helpers -----------------------------------------------------------------------

// Scanning bins from the top, find the bin where the running count first reaches `need`.
// Thread t owns bins [nb - (t+1)*per, nb - t*per). Results in *s_bin / *s_above. All 256 threads call.
__device__ __forceinline__ void find_bin_from_top(const uint32_t* hist, int nb, uint32_t need, int tid, uint32_t* wsum,
                                                  int* s_bin, uint32_t* s_above) {
    const int per = nb >> 8;
    const int hi = nb - tid * per, lo = hi - per;
    uint32_t local = 0;
    for (int b = lo; b < hi; ++b) local += hist[b];
    // exclusive scan over threads (thread order == from the top)
    const int lane = tid & 31, wid = tid >> 5;
    uint32_t inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < wid; ++w) base += wsum[w];
    const uint32_t excl = base + inc - local;
    if (excl < need && need <= excl + local) {
        uint32_t running = excl;
        for (int b = hi - 1; b >= lo; --b) {
            const uint32_t h = hist[b];
            if (running + h >= need) { *s_bin = b; *s_above = running; break; }
            running += h;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
select_rows_kernel(const float* __restrict__ scores, long long ld, int n, int k, int kp, long long index_base,
                   float* __restrict__ out_s, int64_t* __restrict__ out_i) {
    extern __shared__ __align__(16) uint64_t keys[];  // [kp]
    __shared__ uint32_t hist[2048];
    __shared__ uint32_t wsum[8];
    __shared__ int s_bin;
    __shared__ uint32_t s_above, s_cgt, s_ceq;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int q = blockIdx.x;
    const float* row = scores + (size_t)q * ld;
    const int kk = min(k, n);
    for (int i = tid; i < kp; i += 256) keys[i] = 0ull;
    if (tid == 0) { s_cgt = 0; s_ceq = 0; s_bin = 0; s_above = 0; }
    __syncthreads();

    if (kk > 0) {
        uint32_t prefix = 0, mask = 0, need = (uint32_t)kk;
        const int shifts[3] = {21, 10, 0};
        const int nbs[3] = {2048, 2048, 1024};
        for (int pass = 0; pass < 3; ++pass) {
            const int shift = shifts[pass], nb = nbs[pass];
            for (int i = tid; i < nb; i += 256) hist[i] = 0;
            __syncthreads();
            for (int i = tid; i < n; i += 256) {
                const uint32_t u = ordered_bits(__ldg(row + i));
                if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & (uint32_t)(nb - 1)], 1u);
            }
            __syncthreads();
            find_bin_from_top(hist, nb, need, tid, wsum, &s_bin, &s_above);
            prefix |= (uint32_t)s_bin << shift;
            mask |= (uint32_t)(nb - 1) << shift;
            need -= s_above;
            __syncthreads();
        }
        const uint32_t T = prefix;               // key of the k-th largest score
        const int n_gt = kk - (int)need;         // scores strictly above it; `need` ties are taken by index
        for (int base = 0; base < n; base += 256) {
            const int i = base + tid;
            uint32_t u = 0;
            bool gt = false, eq = false;
            if (i < n) {
                u = ordered_bits(__ldg(row + i));
                gt = u > T;
                eq = u == T;
            }
            if (__syncthreads_or(gt || eq)) {
                if (gt) {
                    const uint32_t pos = atomicAdd(&s_cgt, 1u);
                    keys[pos] = rank_key_bits(u, (uint32_t)i);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, eq);
                if (lane == 0) wsum[wid] = __popc(bal);
                __syncthreads();
                uint32_t before = s_ceq, total = 0;
                for (int w = 0; w < 8; ++w) {
                    if (w < wid) before += wsum[w];
                    total += wsum[w];
                }
                if (eq) {
                    const uint32_t p = before + __popc(bal & ((1u << lane) - 1u));
                    if (p < need) keys[n_gt + p] = rank_key_bits(u, (uint32_t)i);
                }
                __syncthreads();
                if (tid == 0) s_ceq += total;
            }
        }
    }
    block_bitonic_sort_desc(keys, kp, tid, 256);
    for (int i = tid; i < k; i += 256) {
        const uint64_t key = keys[i];
        const bool valid = i < kk;
        out_s[(size_t)q * k + i] = valid ? key_score(key) : __int_as_float(0xff800000);
        out_i[(size_t)q * k + i] = valid ? (int64_t)(index_base + (long long)key_index(key)) : (int64_t)-1;
    }
}

__global__ void __launch_bounds__(256)
topk_merge_kernel(const float* __restrict__ scores, const int64_t* __restrict__ idx, int g, int nq, int k, int np,
                  float* __restrict__ out_s, int64_t* __restrict__ out_i) {
    extern __shared__ __align__(16) uint64_t keys[];  // [np]
    __shared__ int s_overflow;
    const int tid = threadIdx.x, q = blockIdx.x;
    const int total = g * k;
    if (tid == 0) s_overflow = 0;
    __syncthreads();
    for (int i = tid; i < np; i += 256) {
        uint64_t key = 0ull;
        if (i < total) {
            const int s = i / k, j = i - s * k;
            const size_t off = ((size_t)s * nq + q) * k + j;
            const int64_t id = idx[off];
            if (id >= 0) key = rank_key(scores[off], (uint32_t)id);
            else if (id == -2) s_overflow = 1;      // a shard's list overflowed (topk_finalize_kernel's marker)
        }
        keys[i] = key;
    }
    block_bitonic_sort_desc(keys, np, tid, 256);
    const bool overflow = s_overflow != 0;
    for (int i = tid; i < k; i += 256) {
        const uint64_t key = keys[i];
        const bool valid = key != 0ull;
        out_s[(size_t)q * k + i] = valid ? key_score(key) : __int_as_float(0xff800000);
        int64_t id = valid ? (int64_t)key_index(key) : (int64_t)-1;
        if (overflow && i == 0) id = -2;
        out_i[(size_t)q * k + i] = id;
    }
}

// ---- packed exchange format ---------------------------------------------------------------------
// One 64-bit word per list entry instead of (fp32 score, int64 index): the library's rank key (select.cuh) of
// (score, GLOBAL index) -- larger key == sorts first -- so per-shard lists travel as 8 bytes per entry in ONE
// all-gather and merge without unpacking. Keys with a zero high word cannot be rank keys (ordered_bits() is zero only
// for a negative NaN): 0 = padding, kKeyOverflow = "this shard's list overflowed, the query needs the exact fallback".
constexpr uint64_t kKeyOverflow = 2ull;

__global__ void __launch_bounds__(256)
topk_pack_kernel(const float* __restrict__ scores, const int64_t* __restrict__ idx, long long n, uint64_t* __restrict__ keys) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const int64_t id = idx[i];
        keys[i] = id >= 0 ? rank_key(scores[i], (uint32_t)id) : (id == -2 ? kKeyOverflow : 0ull);
    }
}

// The per-shard lists arrive sorted (key descending, padding at the tail -- what topk_finalize_kernel and
// gdt_score_topk_exact write), and rank keys are unique across shards (they embed the global row index). So the merged
// position of an entry is its position in its own list plus, for every other list, the number of larger keys there:
// g - 1 binary searches per entry, no sort, one barrier. Unsorted input (a caller's own lists) is detected and takes the
// bitonic sort instead. The merged order is the same either way.
__global__ void __launch_bounds__(256)
topk_merge_packed_kernel(const uint64_t* __restrict__ in, int g, int nq, int k, int np, float* __restrict__ out_s,
                         int64_t* __restrict__ out_i, uint64_t* __restrict__ out_k) {
    extern __shared__ __align__(16) uint64_t keys[];  // [np], list s at [s * k, (s + 1) * k)
    __shared__ int s_overflow;
    const int tid = threadIdx.x, q = blockIdx.x;
    const int total = g * k;
    if (tid == 0) s_overflow = 0;
    __syncthreads();
    for (int i = tid; i < np; i += 256) {
        uint64_t key = 0ull;
        if (i < total) {
            const int s = i / k, j = i - s * k;
            key = in[((size_t)s * nq + q) * k + j];
            if ((key >> 32) == 0ull) {
                if (key == kKeyOverflow) s_overflow = 1;
                key = 0ull;
            }
        }
        keys[i] = key;
    }
    __syncthreads();
    const bool overflow = s_overflow != 0;
    bool unsorted = false;
    {
        int s = tid / k, j = tid - s * k;             // (list, position) of entry i, advanced incrementally
        const int ds = 256 / k, dj = 256 - ds * k;
        for (int i = tid; i < total; i += 256) {
            if (j + 1 < k && keys[i] < keys[i + 1]) unsorted = true;
            s += ds; j += dj;
            if (j >= k) { j -= k; ++s; }
        }
    }
    if (__syncthreads_or(unsorted || overflow)) {
        block_bitonic_sort_desc(keys, np, tid, 256);
        for (int i = tid; i < k; i += 256) {
            const uint64_t key = keys[i];
            const bool valid = key != 0ull;
            if (out_k) {                          // keys out (query-sharded merge): the overflow marker travels on
                out_k[(size_t)q * k + i] = (overflow && i == 0) ? kKeyOverflow : key;
                continue;
            }
            out_s[(size_t)q * k + i] = valid ? key_score(key) : __int_as_float(0xff800000);
            int64_t id = valid ? (int64_t)key_index(key) : (int64_t)-1;
            if (overflow && i == 0) id = -2;      // same marker as topk_finalize_kernel: the caller repairs this query
            out_i[(size_t)q * k + i] = id;
        }
        return;
    }
    // Entry j of a list has at least j entries before it: only the first k positions of the merged order are wanted,
    // so the walk over the other lists stops as soon as the rank reaches k (most entries stop after one or two lists).
    int nvalid = 0;
    {
        int s = tid / k, j = tid - s * k;
        const int ds = 256 / k, dj = 256 - ds * k;
        for (int i = tid; i < total; i += 256) {
            const uint64_t key = keys[i];
            if (key != 0ull) {
                ++nvalid;
                int rank = j;
                for (int t = 0; t < g && rank < k; ++t) {
                    if (t == s) continue;
                    const uint64_t* lst = keys + t * k;
                    // number of entries of list t that sort before `key`: larger keys, and equal ones in earlier lists
                    // (a caller may hand in the same entry twice; the merge stays a stable one). Entries at or beyond
                    // position k - rank cannot change the outcome: the search is confined to the first k - rank.
                    int lo = 0, hi = k - rank;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        const uint64_t m = lst[mid];
                        if (m > key || (t < s && m == key)) lo = mid + 1; else hi = mid;
                    }
                    rank += lo;
                }
                if (rank < k) {
                    if (out_k) {
                        out_k[(size_t)q * k + rank] = key;
                    } else {
                        out_s[(size_t)q * k + rank] = key_score(key);
                        out_i[(size_t)q * k + rank] = (int64_t)key_index(key);
                    }
                }
            }
            s += ds; j += dj;
            if (j >= k) { j -= k; ++s; }
        }
    }
    // fewer than k valid entries in total: pad the tail
    __shared__ int s_valid;
    if (tid == 0) s_valid = 0;
    __syncthreads();
    if (nvalid) atomicAdd(&s_valid, nvalid);
    __syncthreads();
    for (int i = s_valid + tid; i < k; i += 256) {
        if (out_k) {
            out_k[(size_t)q * k + i] = 0ull;
        } else {
            out_s[(size_t)q * k + i] = __int_as_float(0xff800000);
            out_i[(size_t)q * k + i] = (int64_t)-1;
        }
    }
}

// keys -> (score, index): 0 -> (-inf, -1) padding, the overflow key -> (-inf, -2)
__global__ void __launch_bounds__(256)
topk_unpack_kernel(const uint64_t* __restrict__ keys, long long n, float* __restrict__ scores, int64_t* __restrict__ idx) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const uint64_t key = keys[i];
        const bool valid = (key >> 32) != 0ull;
        scores[i] = valid ? key_score(key) : __int_as_float(0xff800000);
        idx[i] = valid ? (int64_t)key_index(key) : (key == kKeyOverflow ? (int64_t)-2 : (int64_t)-1);
    }
}

int launch_exact_scores(const float* q, const float* db, int nq, long long ndb, int d, float* scores, long long ld,
                        cudaStream_t stream) {
    const int dpad = (d + 3) & ~3;
    const size_t smem = (size_t)kExactQB * dpad * sizeof(float);
    if (smem > 200 * 1024) return GDT_ERR_UNSUPPORTED;
    static size_t attr_bytes_dev[32] = {0};
    size_t& attr_bytes = attr_bytes_dev[current_device_slot()];
    if (smem > 48 * 1024 && smem > attr_bytes) {
        GDT_CUDA(cudaFuncSetAttribute(exact_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_bytes = smem;
    }
    const int qblocks = ceil_div(nq, kExactQB);
    const int sms = sm_count_current_device();
    long long want = ceil_div_ll(8LL * sms, qblocks);  // row chunks for ~8 CTAs per SM overall
    if (want < 1) want = 1;
    long long rows = ceil_div_ll(ndb, want);
    if (rows < 64) rows = 64;
    if (rows > (1 << 20)) rows = 1 << 20;
    rows = (rows + 7) / 8 * 8;
    dim3 grid((unsigned)ceil_div_ll(ndb, rows), (unsigned)qblocks);
    exact_scores_kernel<<<grid, 256, smem, stream>>>(q, db, nq, ndb, d, dpad, scores, ld, (int)rows);
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}

int launch_select_rows(const float* scores, long long ld, int nq, int n, int k, long long index_base, float* out_s,
                       int64_t* out_i, cudaStream_t stream) {
    const int kp = next_pow2(k < 1 ? 1 : k);
    select_rows_kernel<<<nq, 256, (size_t)kp * sizeof(uint64_t), stream>>>(scores, ld, n, k, kp, index_base, out_s, out_i);
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}

}  // namespace gdt

using namespace gdt;

static bool have_device() {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return false;
    }
    return true;
}

extern "C" size_t gdt_score_topk_exact_workspace_bytes(int nq, long long ndb, int d, int k) {
    (void)d; (void)k;
    if (nq <= 0 || ndb <= 0) return 0;
    return align_up((size_t)nq * (size_t)ndb * sizeof(float), 256) + 256;
}

extern "C" int gdt_score_topk_exact(const float* q, const float* db, int nq, long long ndb, int d, int k,
                                    long long index_base, float* top_scores, int64_t* top_idx, void* ws,
                                    size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!q || !db || !top_scores || !top_idx || !ws) return GDT_ERR_INVALID_ARGUMENT;
    if (nq <= 0 || ndb <= 0 || d <= 0 || k <= 0) return GDT_ERR_INVALID_ARGUMENT;
    if (k > 4096 || ndb > 0x7fffffffLL || index_base < 0 || index_base + ndb > 0xffffffffLL) return GDT_ERR_UNSUPPORTED;
    if (!have_device()) return GDT_ERR_NO_DEVICE;
    if (ws_bytes < gdt_score_topk_exact_workspace_bytes(nq, ndb, d, k) || (((uintptr_t)ws) & 255)) return GDT_ERR_WORKSPACE_TOO_SMALL;
    float* scores = (float*)ws;
    int rc = launch_exact_scores(q, db, nq, ndb, d, scores, ndb, stream);
    if (rc != GDT_OK) return rc;
    return launch_select_rows(scores, ndb, nq, (int)ndb, k, index_base, top_scores, top_idx, stream);
}

extern "C" int gdt_topk_merge(const float* scores, const int64_t* idx, int g, int nq, int k, float* out_scores,
                              int64_t* out_idx, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!scores || !idx || !out_scores || !out_idx) return GDT_ERR_INVALID_ARGUMENT;
    if (g <= 0 || nq <= 0 || k <= 0) return GDT_ERR_INVALID_ARGUMENT;
    const long long total = (long long)g * k;
    if (total > 16384) return GDT_ERR_UNSUPPORTED;
    if (!have_device()) return GDT_ERR_NO_DEVICE;
    const int np = next_pow2((int)total);
    const size_t smem = (size_t)np * sizeof(uint64_t);
    static size_t attr_bytes_dev[32] = {0};
    size_t& attr_bytes = attr_bytes_dev[current_device_slot()];
    if (smem > 48 * 1024 && smem > attr_bytes) {
        GDT_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_bytes = smem;
    }
    topk_merge_kernel<<<nq, 256, smem, stream>>>(scores, idx, g, nq, k, np, out_scores, out_idx);
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}

extern "C" int gdt_topk_pack(const float* scores, const int64_t* idx, long long n, uint64_t* keys, void* stream_) {
    if (!scores || !idx || !keys || n < 0) return GDT_ERR_INVALID_ARGUMENT;
    if (!have_device()) return GDT_ERR_NO_DEVICE;
    if (n == 0) return GDT_OK;
    long long blocks = ceil_div_ll(n, 256);
    const long long cap = (long long)sm_count_current_device() * 8;
    if (blocks > cap) blocks = cap;
    topk_pack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(scores, idx, n, keys);
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}

static int topk_merge_packed_impl(const uint64_t* keys, int g, int nq, int k, float* out_scores, int64_t* out_idx,
                                  uint64_t* out_keys, cudaStream_t stream) {
    if (!keys || (!out_keys && (!out_scores || !out_idx))) return GDT_ERR_INVALID_ARGUMENT;
    if (g <= 0 || nq <= 0 || k <= 0) return GDT_ERR_INVALID_ARGUMENT;
    const long long total = (long long)g * k;
    if (total > 16384) return GDT_ERR_UNSUPPORTED;
    if (!have_device()) return GDT_ERR_NO_DEVICE;
    const int np = next_pow2((int)total);
    const size_t smem = (size_t)np * sizeof(uint64_t);
    static size_t attr_bytes_dev[32] = {0};
    size_t& attr_bytes = attr_bytes_dev[current_device_slot()];
    if (smem > 48 * 1024 && smem > attr_bytes) {
        GDT_CUDA(cudaFuncSetAttribute(topk_merge_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_bytes = smem;
    }
    topk_merge_packed_kernel<<<nq, 256, smem, stream>>>(keys, g, nq, k, np, out_scores, out_idx, out_keys);
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}

extern "C" int gdt_topk_merge_packed(const uint64_t* keys, int g, int nq, int k, float* out_scores, int64_t* out_idx,
                                     void* stream_) {
    return topk_merge_packed_impl(keys, g, nq, k, out_scores, out_idx, nullptr, (cudaStream_t)stream_);
}

extern "C" int gdt_topk_merge_packed_keys(const uint64_t* keys, int g, int nq, int k, uint64_t* out_keys, void* stream_) {
    return topk_merge_packed_impl(keys, g, nq, k, nullptr, nullptr, out_keys, (cudaStream_t)stream_);
}

extern "C" int gdt_topk_unpack(const uint64_t* keys, long long n, float* scores, int64_t* idx, void* stream_) {
    if (!keys || !scores || !idx || n < 0) return GDT_ERR_INVALID_ARGUMENT;
    if (!have_device()) return GDT_ERR_NO_DEVICE;
    if (n == 0) return GDT_OK;
    long long blocks = ceil_div_ll(n, 256);
    const long long cap = (long long)sm_count_current_device() * 8;
    if (blocks > cap) blocks = cap;
    topk_unpack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(keys, n, scores, idx);
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}
