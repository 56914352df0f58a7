// N2 -- the Gram products of whitening learning on the device, for sm_100a.
// `whitenlearn` / `pcawhitenlearn` (mdir/external/cirtorch/utils/whiten.py:14-50) spend their time in three symmetric
// rank-n updates of D x D matrices over n = 10^4..10^5 descriptors, in float64:
//     S = df df^T / n      D = (P (X - m)) (P (X - m))^T      Xcov = Xc Xc^T
// gdt_syrk_f64 computes C = alpha * A A^T for a row-major A [d][n]: only the lower-triangular 64 x 64 tiles are computed
// (half the flops of a general product) and mirrored, n is split over the grid so that small D still fills 148 SMs, and
// the per-split partial tiles are summed in a FIXED order -- the result is bit-reproducible run to run (no fp64 atomics),
// which keeps the Cholesky "not positive definite" retry logic of the stage function deterministic.
// The factorisations themselves (Cholesky, inverse, symmetric eigendecomposition) stay with torch.linalg / cuSOLVER, as
// the survey prescribes. fp64 FMA throughput bounds this kernel; it runs once per learned whitening.
#include "common.cuh"

namespace gdt {

constexpr int kSyrkTile = 64;   // C tile edge
constexpr int kSyrkK = 16;      // k chunk

__global__ void __launch_bounds__(256)
syrk_f64_kernel(const double* __restrict__ A, int d, long long n, long long lda, double* __restrict__ part, int nsplit,
                int ntiles_edge) {
    __shared__ double As[kSyrkK][kSyrkTile + 2];
    __shared__ double Bs[kSyrkK][kSyrkTile + 2];
    // lower-triangular tile (bi >= bj) from the linear index
    int t = blockIdx.x, bi = 0;
    while (t >= bi + 1) { t -= bi + 1; ++bi; }
    const int bj = t;
    (void)ntiles_edge;
    const int split = blockIdx.y;
    const long long chunk = (n + nsplit - 1) / nsplit;
    const long long k0 = (long long)split * chunk, k1 = min(k0 + chunk, n);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int lrow = tid >> 2, lk = (tid & 3) * 4;        // loader: row of the tile, first of 4 consecutive k
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    const int ri = bi * kSyrkTile + lrow, rj = bj * kSyrkTile + lrow;
    for (long long k = k0; k < k1; k += kSyrkK) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const long long kk = k + lk + c;
            As[lk + c][lrow] = (ri < d && kk < k1) ? A[(size_t)ri * lda + kk] : 0.0;
            Bs[lk + c][lrow] = (rj < d && kk < k1) ? A[(size_t)rj * lda + kk] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kSyrkK; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) { a[c] = As[kk][ty * 4 + c]; b[c] = Bs[kk][tx * 4 + c]; }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
        }
        __syncthreads();
    }
    double* out = part + ((size_t)split * gridDim.x + blockIdx.x) * (kSyrkTile * kSyrkTile);
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) out[(ty * 4 + x) * kSyrkTile + tx * 4 + y] = acc[x][y];
}

// sums the splits in order 0, 1, 2, ... and writes both halves of C
__global__ void __launch_bounds__(256)
syrk_f64_finish_kernel(const double* __restrict__ part, int nsplit, int ntile_total, int d, double alpha, double* __restrict__ C) {
    int t = blockIdx.x, bi = 0;
    while (t >= bi + 1) { t -= bi + 1; ++bi; }
    const int bj = t;
    for (int e = threadIdx.x; e < kSyrkTile * kSyrkTile; e += 256) {
        double s = 0.0;
        for (int sp = 0; sp < nsplit; ++sp) s += part[((size_t)sp * ntile_total + blockIdx.x) * (kSyrkTile * kSyrkTile) + e];
        s *= alpha;
        const int i = bi * kSyrkTile + e / kSyrkTile, j = bj * kSyrkTile + e % kSyrkTile;
        if (i < d && j < d) {
            if (bi != bj || j <= i) {
                C[(size_t)i * d + j] = s;
                C[(size_t)j * d + i] = s;
            }
        }
    }
}

static void syrk_plan(int d, long long n, int& edge, int& ntiles, int& nsplit) {
    edge = ceil_div(d, kSyrkTile);
    ntiles = edge * (edge + 1) / 2;
    const int sms = sm_count_current_device();
    nsplit = ceil_div(2 * sms, ntiles);
    const long long max_split = ceil_div_ll(n, 4 * kSyrkK);     // at least 64 columns per split
    if (nsplit > max_split) nsplit = (int)max_split;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > 64) nsplit = 64;
}

}  // namespace gdt

using namespace gdt;

extern "C" size_t gdt_syrk_f64_workspace_bytes(int d, long long n) {
    if (d <= 0 || n <= 0) return 0;
    int edge, ntiles, nsplit;
    syrk_plan(d, n, edge, ntiles, nsplit);
    return align_up((size_t)nsplit * ntiles * kSyrkTile * kSyrkTile * sizeof(double), 256) + 256;
}

extern "C" int gdt_syrk_f64(const double* A, int d, long long n, long long lda, double alpha, double* C, void* ws,
                            size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!A || !C || !ws || d <= 0 || n <= 0 || lda < n) return GDT_ERR_INVALID_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return GDT_ERR_NO_DEVICE; }
    if (ws_bytes < gdt_syrk_f64_workspace_bytes(d, n) || (((uintptr_t)ws) & 255)) return GDT_ERR_WORKSPACE_TOO_SMALL;
    int edge, ntiles, nsplit;
    syrk_plan(d, n, edge, ntiles, nsplit);
    double* part = (double*)ws;
    syrk_f64_kernel<<<dim3(ntiles, nsplit), 256, 0, stream>>>(A, d, n, lda, part, nsplit, edge);
    GDT_LAUNCH_CHECK();
    syrk_f64_finish_kernel<<<ntiles, 256, 0, stream>>>(part, nsplit, ntiles, d, alpha, C);
    GDT_LAUNCH_CHECK();
    return GDT_OK;
}
