// Microbenchmark: what does one scattered lattice-record gather per pixel cost on a B200 SM, by record size, pipe (LSU /
// texture) and table layout?  The access pattern is K1's: a thread owns 4 consecutive pixels, the lattice cell of every
// pixel comes from the synthetic image family of SURVEY 8(d) (smooth sinusoid + N(0, 8) noise), so a warp request touches
// ~31 distinct cells. Prints SM cycles per pixel (148 SMs) for each variant; K1's passes cost 1.0 (A) and 2.2 (B) today.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_rate gather_rate.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

static const int H = 768, W = 1024, NIMG = 16;
static const int NCELL = 33 * 33 * 33;

__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg((const uint4*)p); }
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ uint32_t mix(uint4 v) { return v.x ^ v.y ^ v.z ^ v.w; }

// MODE: see main()
template <int MODE>
__global__ void __launch_bounds__(256, 4)
gather(const uint32_t* __restrict__ cells, const uint8_t* __restrict__ t16, const uint8_t* __restrict__ t32,
       const uint8_t* __restrict__ t64, cudaTextureObject_t x16, cudaTextureObject_t x32, cudaTextureObject_t x64,
       uint32_t* __restrict__ out, long long ngroups) {
    __shared__ uint2 spl[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) spl[i] = make_uint2(i, i * 3);
    __syncthreads();
    uint32_t acc = 0;
    for (long long g = (long long)blockIdx.x * 256 + threadIdx.x; g < ngroups; g += (long long)gridDim.x * 256) {
        const uint4 c4 = __ldg((const uint4*)cells + g);
        const uint32_t c[4] = {c4.x, c4.y, c4.z, c4.w};
        uint4 r[4][3];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            r[i][0] = r[i][1] = r[i][2] = make_uint4(0, 0, 0, 0);
            if (MODE == 0) r[i][0] = ldg128(t16 + (size_t)c[i] * 16);
            if (MODE == 1) {
                r[i][0] = ldg128(t64 + (size_t)c[i] * 64);
                r[i][1] = ldg128(t64 + (size_t)c[i] * 64 + 16);
                r[i][2] = ldg128(t64 + (size_t)c[i] * 64 + 32);
            }
            if (MODE == 2) {
                ldg256(t64 + (size_t)c[i] * 64, r[i][0], r[i][1]);
                r[i][2] = ldg128(t64 + (size_t)c[i] * 64 + 32);
            }
            if (MODE == 3) r[i][0] = tex1Dfetch<uint4>(x16, c[i]);
            if (MODE == 4) {
                r[i][0] = tex1Dfetch<uint4>(x32, c[i] * 2);
                r[i][1] = tex1Dfetch<uint4>(x32, c[i] * 2 + 1);
            }
            if (MODE == 5) {
                r[i][0] = tex1Dfetch<uint4>(x64, c[i] * 4);
                r[i][1] = tex1Dfetch<uint4>(x64, c[i] * 4 + 1);
                r[i][2] = tex1Dfetch<uint4>(x64, c[i] * 4 + 2);
            }
            if (MODE == 6) {
                r[i][0] = ldg128(t64 + (size_t)c[i] * 64);
                r[i][1] = tex1Dfetch<uint4>(x64, c[i] * 4 + 1);
                r[i][2] = tex1Dfetch<uint4>(x64, c[i] * 4 + 2);
            }
            if (MODE == 7) {           // LSU for the 16-byte table, texture pipe for the 32-byte one (separate tables)
                r[i][0] = ldg128(t16 + (size_t)c[i] * 16);
                r[i][1] = tex1Dfetch<uint4>(x32, c[i] * 2);
                r[i][2] = tex1Dfetch<uint4>(x32, c[i] * 2 + 1);
            }
            if (MODE == 8) {           // 8-byte records (what a half-size record would cost)
                const uint2 v = __ldg((const uint2*)(t16 + (size_t)c[i] * 8));
                r[i][0] = make_uint4(v.x, v.y, 0, 0);
            }
            if (MODE == 9) {           // 4-byte records
                r[i][0].x = __ldg((const uint32_t*)(t16 + (size_t)c[i] * 4));
            }
            if (MODE == 10) {          // three random 8-byte x2 shared-memory lookups per pixel (pass B's spline pattern)
                const uint32_t h0 = (c[i] * 2654435761u) >> 22, h1 = (c[i] * 40503u + 77u) & 1023u, h2 = (c[i] * 9176u + 5u) & 1023u;
                const uint2 a0 = spl[h0], a1 = spl[1024 + h0], b0 = spl[h1], b1 = spl[1024 + h1], d0 = spl[h2], d1 = spl[1024 + h2];
                r[i][0] = make_uint4(a0.x, a0.y, a1.x, a1.y);
                r[i][1] = make_uint4(b0.x, b0.y, b1.x, b1.y);
                r[i][2] = make_uint4(d0.x, d0.y, d1.x, d1.y);
            }
            if (MODE == 11) {          // the same three lookups as single 16-byte shared-memory loads
                const uint32_t h0 = (c[i] * 2654435761u) >> 22, h1 = (c[i] * 40503u + 77u) & 1023u, h2 = (c[i] * 9176u + 5u) & 1023u;
                const uint4* s4 = (const uint4*)spl;
                r[i][0] = s4[h0]; r[i][1] = s4[h1]; r[i][2] = s4[h2];
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) acc += mix(r[i][0]) + mix(r[i][1]) * 3u + mix(r[i][2]) * 5u;
    }
    out[blockIdx.x * 256 + threadIdx.x] = acc;
}

static cudaTextureObject_t make_tex(void* p, size_t bytes) {
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = p;
    rd.res.linear.desc = cudaCreateChannelDesc<uint4>();
    rd.res.linear.sizeInBytes = bytes;
    cudaTextureDesc td = {};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t t = 0;
    cudaCreateTextureObject(&t, &rd, &td, nullptr);
    return t;
}

static int cell_of(int r, int g, int b, bool morton) {
    auto t = [](int v) { return (int)((((unsigned)v * 514u + 4u) >> 8) >> 4); };
    const int tr = t(r), tg = t(g), tb = t(b);
    if (!morton) return (tr * 33 + tg) * 33 + tb;
    // 2x2x2 blocks of cells contiguous (one 128-byte line of 16-byte records); 17^3 blocks
    return (((tr >> 1) * 17 + (tg >> 1)) * 17 + (tb >> 1)) * 8 + ((tr & 1) << 2 | (tg & 1) << 1 | (tb & 1));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs; %d images of %dx%d, smooth + N(0,8) family\n", p.name, sms, NIMG, W, H);
    const size_t npx = (size_t)NIMG * H * W;
    std::vector<uint32_t> cells(npx), cells_m(npx), cells_u(npx);
    srand(1);
    auto gauss = []() { double s = 0; for (int i = 0; i < 12; ++i) s += rand() / (double)RAND_MAX; return s - 6.0; };
    double distinct = 0; long nreq = 0;
    for (int n = 0; n < NIMG; ++n) {
        const double gain = 0.6 + 0.8 * (n / (double)NIMG);
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                int v[3];
                for (int c = 0; c < 3; ++c) {
                    double b = (128 + 70 * (sin(x / (37.0 + 5 * c)) + cos(y / (23.0 + 3 * c)))) * gain + 8.0 * gauss();
                    v[c] = b < 0 ? 0 : (b > 255 ? 255 : (int)b);
                }
                const size_t i = ((size_t)n * H + y) * W + x;
                cells[i] = cell_of(v[0], v[1], v[2], false);
                cells_m[i] = cell_of(v[0], v[1], v[2], true);
                cells_u[i] = cell_of(rand() & 255, rand() & 255, rand() & 255, false);
            }
    }
    // distinct cells / 32-byte sectors / 128-byte lines of 16-byte records per warp request (lane l, pixel 4l + i)
    double dl = 0, ds = 0, dlm = 0;
    for (size_t base = 0; base + 128 <= npx && nreq < 200000; base += 128 * 37)
        for (int i = 0; i < 4; ++i) {
            int uc = 0, us = 0, ul = 0, um = 0;
            uint32_t seen[32], seen_s[32], seen_l[32], seen_m[32];
            for (int l = 0; l < 32; ++l) {
                const uint32_t c = cells[base + 4 * l + i], m = cells_m[base + 4 * l + i];
                bool f = false; for (int k = 0; k < uc; ++k) f |= seen[k] == c; if (!f) seen[uc++] = c;
                f = false; for (int k = 0; k < us; ++k) f |= seen_s[k] == c / 2; if (!f) seen_s[us++] = c / 2;
                f = false; for (int k = 0; k < ul; ++k) f |= seen_l[k] == c / 8; if (!f) seen_l[ul++] = c / 8;
                f = false; for (int k = 0; k < um; ++k) f |= seen_m[k] == m / 8; if (!f) seen_m[um++] = m / 8;
            }
            distinct += uc; ds += us; dl += ul; dlm += um; ++nreq;
        }
    printf("per warp request: %.1f distinct cells, %.1f sectors and %.1f lines (16-byte records, b fastest), %.1f lines (2x2x2 blocks)\n",
           distinct / nreq, ds / nreq, dl / nreq, dlm / nreq);

    uint32_t *dc, *dcm, *dcu, *dout;
    uint8_t *t16, *t32, *t64;
    cudaMalloc(&dc, npx * 4); cudaMalloc(&dcm, npx * 4); cudaMalloc(&dcu, npx * 4);
    cudaMemcpy(dc, cells.data(), npx * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dcm, cells_m.data(), npx * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dcu, cells_u.data(), npx * 4, cudaMemcpyHostToDevice);
    const size_t ncell_pad = 17 * 17 * 17 * 8;      // covers both layouts
    cudaMalloc(&t16, ncell_pad * 16); cudaMalloc(&t32, ncell_pad * 32); cudaMalloc(&t64, ncell_pad * 64);
    cudaMemset(t16, 1, ncell_pad * 16); cudaMemset(t32, 2, ncell_pad * 32); cudaMemset(t64, 3, ncell_pad * 64);
    cudaMalloc(&dout, sizeof(uint32_t) * 256 * sms * 4);
    cudaTextureObject_t x16 = make_tex(t16, ncell_pad * 16), x32 = make_tex(t32, ncell_pad * 32), x64 = make_tex(t64, ncell_pad * 64);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);

    auto run = [&](const char* name, auto kern, const uint32_t* cellptr) {
        const long long ngroups = (long long)npx / 4;
        float best = 1e9f;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            kern<<<sms * 4, 256>>>(cellptr, t16, t32, t64, x16, x32, x64, dout, ngroups);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep >= 2 && ms < best) best = ms;
        }
        printf("%-72s %7.3f ms  %6.3f ns/px/SM-normalised  ~%5.2f SM-cycles/px at %d MHz\n", name, best,
               best * 1e6 / npx * sms, best * 1e-3 * (clk * 1e3) * sms / npx, clk / 1000);
    };
    run("0  LDG.128, 16-B records", gather<0>, dc);
    run("0u LDG.128, 16-B records, UNIFORM random colours", gather<0>, dcu);
    run("0m LDG.128, 16-B records, 2x2x2-block (Morton) layout", gather<0>, dcm);
    run("1  3 x LDG.128 from one 64-B record", gather<1>, dc);
    run("1m 3 x LDG.128 from one 64-B record, 2x2x2-block layout", gather<1>, dcm);
    run("2  LDG.256 + LDG.128 from one 64-B record", gather<2>, dc);
    run("3  TEX.128, 16-B records", gather<3>, dc);
    run("3m TEX.128, 16-B records, 2x2x2-block layout", gather<3>, dcm);
    run("4  2 x TEX.128 from one 32-B record (pass B today)", gather<4>, dc);
    run("5  3 x TEX.128 from one 64-B record", gather<5>, dc);
    run("6  LDG.128 + 2 x TEX.128 from one 64-B record", gather<6>, dc);
    run("7  LDG.128 (16-B table) + 2 x TEX.128 (32-B table)", gather<7>, dc);
    run("7m same, 2x2x2-block layout", gather<7>, dcm);
    run("8  LDG.64, 8-B records", gather<8>, dc);
    run("9  LDG.32, 4-B records", gather<9>, dc);
    run("10 3 x (2 x LDS.64) random shared-memory lookups", gather<10>, dc);
    run("11 3 x LDS.128 random shared-memory lookups", gather<11>, dc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
