// N3 -- diverse-anchor mining on the device, for sm_100a (second half).
// Replaces the greedy loop of `DiverseAnchorsDataset._select_positive_pairs_db`
// (mdir/components/data/dataset/cirtorch_datasets.py:68-100): qsize-1 sequential rounds of
//     dist = qvecs.T @ qvecs[:, idx]; dists = cat(dists, dist); most_similar = dists.max(1)
//     idx  = most_similar.argsort()[dissimilar_split:similar_split][choice]
// i.e. per round one matrix-vector product, a running maximum and ONE order statistic of the pool (the element at a
// given ascending rank). The reference grows a [pool, rounds] matrix, re-reduces it and fully sorts the pool every round,
// with a host sync (`.item()`) per round. Here a round is two launches and no host involvement:
//   diverse_step_kernel  one warp per pool row: exact score (fp64-accumulated, select.cuh) against the row picked last
//                        (its index is read from device memory), most_similar[i] = max(most_similar[i], score)
//   select_rank_kernel   one CTA: radix select (6 digits of the 64-bit key value-bits | index) of the element at the
//                        requested ascending rank under the total order (value asc, index asc); appends it to the picks
// The per-round target ranks depend only on the pool size and the exclude / include fractions (and on the caller's
// random choices), never on the data, so the host computes them up front (gandtr_b200/mining.py).
#include "common.cuh"
#include "select.cuh"

namespace gdt {

__global__ void __launch_bounds__(256)
diverse_step_kernel(const float* __restrict__ pool, int n, int d, int dpad, const int32_t* __restrict__ picked, int step,
                    float* __restrict__ most_similar) {
    extern __shared__ __align__(16) float cur[];          // [dpad] the row picked in the previous round
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = picked[step];
    const float* crow = pool + (size_t)c * d;
    for (int i = tid; i < dpad; i += 256) cur[i] = i < d ? crow[i] : 0.f;
    __syncthreads();
    const int row = blockIdx.x * 8 + warp;
    if (row >= n) return;
    const float s = warp_exact_dot(cur, pool + (size_t)row * d, d, lane);
    if (lane == 0) {
        const float m = most_similar[row];
        most_similar[row] = (step == 0 || s > m) ? s : m;
    }
}

// ascending total order on (value, index): 64-bit key, smaller key == earlier
__device__ __forceinline__ uint64_t asc_key(float v, uint32_t idx) { return ((uint64_t)ordered_bits(v) << 32) | idx; }

__global__ void __launch_bounds__(1024)
select_rank_kernel(const float* __restrict__ vals, int n, const int32_t* __restrict__ ranks, int step,
                   int32_t* __restrict__ picked, float* __restrict__ picked_score) {
    __shared__ uint32_t hist[2048];
    __shared__ uint32_t wsum[32];
    __shared__ int s_bin;
    __shared__ uint32_t s_below;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t need = (uint32_t)ranks[step];                 // 0-based rank inside the current candidate subset
    if (need >= (uint32_t)n) need = (uint32_t)n - 1;
    uint64_t prefix = 0ull, mask = 0ull;
    const int shifts[6] = {53, 42, 32, 21, 10, 0};
    const int bits[6] = {11, 11, 10, 11, 11, 10};
    for (int pass = 0; pass < 6; ++pass) {
        const int shift = shifts[pass], nb = 1 << bits[pass];
        for (int i = tid; i < 2048; i += 1024) hist[i] = 0;
        if (tid == 0) { s_bin = 0; s_below = 0; }
        __syncthreads();
        for (int i = tid; i < n; i += 1024) {
            const uint64_t key = asc_key(vals[i], (uint32_t)i);
            if ((key & mask) == prefix) atomicAdd(&hist[(uint32_t)(key >> shift) & (uint32_t)(nb - 1)], 1u);
        }
        __syncthreads();
        // thread t owns bins [2t, 2t + 2): exclusive scan from the bottom, then the owner of the target walks its bins
        const uint32_t h0 = 2 * tid < nb ? hist[2 * tid] : 0u, h1 = 2 * tid + 1 < nb ? hist[2 * tid + 1] : 0u;
        const uint32_t local = h0 + h1;
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
        if (local != 0 && excl <= need && need < excl + local) {
            if (need < excl + h0) { s_bin = 2 * tid; s_below = excl; }
            else { s_bin = 2 * tid + 1; s_below = excl + h0; }
        }
        __syncthreads();
        prefix |= (uint64_t)(uint32_t)s_bin << shift;
        mask |= (uint64_t)(uint32_t)(nb - 1) << shift;
        need -= s_below;
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t idx = (uint32_t)prefix;
        picked[step + 1] = (int32_t)idx;
        picked_score[step] = vals[idx];
    }
}

}  // namespace gdt

using namespace gdt;

extern "C" size_t gdt_diverse_anchors_workspace_bytes(int n) { return n > 0 ? align_up((size_t)n * sizeof(float), 256) + 256 : 0; }

extern "C" int gdt_diverse_anchors(const float* pool, int n, int d, const int32_t* ranks_dev, int steps, int first,
                                   int32_t* picked_dev, float* picked_score_dev, void* ws, size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!pool || !picked_dev || !ws || n <= 0 || d <= 0 || steps < 0 || first < 0 || first >= n) return GDT_ERR_INVALID_ARGUMENT;
    if (steps > 0 && (!ranks_dev || !picked_score_dev)) return GDT_ERR_INVALID_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return GDT_ERR_NO_DEVICE; }
    if (ws_bytes < gdt_diverse_anchors_workspace_bytes(n) || (((uintptr_t)ws) & 255)) return GDT_ERR_WORKSPACE_TOO_SMALL;
    const int dpad = (d + 3) & ~3;
    if ((size_t)dpad * 4 > 48 * 1024) return GDT_ERR_UNSUPPORTED;
    float* most_similar = (float*)ws;
    const int32_t first32 = first;
    GDT_CUDA(cudaMemcpyAsync(picked_dev, &first32, sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    for (int t = 0; t < steps; ++t) {
        diverse_step_kernel<<<ceil_div(n, 8), 256, (size_t)dpad * 4, stream>>>(pool, n, d, dpad, picked_dev, t, most_similar);
        select_rank_kernel<<<1, 1024, 0, stream>>>(most_similar, n, ranks_dev, t, picked_dev, picked_score_dev);
    }
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}
