// Microbenchmark: issue rate of packed FFMA2 / FMUL2 against scalar FFMA / FMUL on sm_100a, alone and interleaved with
// integer ALU work (the mix of K1's pass B). Prints warp-instructions per cycle per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_rate f32x2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

template <int MODE>
__global__ void __launch_bounds__(512) bench(float* out, float one, float c, unsigned izero, long long* cycles) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    unsigned u[4] = {threadIdx.x, threadIdx.x * 3u, threadIdx.x * 7u + 1u, threadIdx.x * 11u + 5u};
    const float2 one2 = make_float2(one, one), c2 = make_float2(c, c);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {            // 8 scalar FFMA
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __fmaf_rn(a[i], c, one);
        } else if (MODE == 1) {     // 4 FFMA2 (same flops)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float2 t = __ffma2_rn(make_float2(a[2 * i], a[2 * i + 1]), c2, one2);
                a[2 * i] = t.x; a[2 * i + 1] = t.y;
            }
        } else if (MODE == 2) {     // 8 scalar FFMA + 8 integer ops
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __fmaf_rn(a[i], c, one);
#pragma unroll
            for (int i = 0; i < 4; ++i) { u[i] = (u[i] ^ izero) + 0x9e3779b9u; u[i] = (u[i] << 5) | (u[i] >> 27); }
        } else if (MODE == 3) {     // 4 FFMA2 + 8 integer ops
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float2 t = __ffma2_rn(make_float2(a[2 * i], a[2 * i + 1]), c2, one2);
                a[2 * i] = t.x; a[2 * i + 1] = t.y;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { u[i] = (u[i] ^ izero) + 0x9e3779b9u; u[i] = (u[i] << 5) | (u[i] >> 27); }
        } else if (MODE == 4) {     // 8 scalar FMUL+FADD pairs (separate roundings)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __fadd_rn(__fmul_rn(a[i], c), one);
        } else if (MODE == 5) {     // 4 x (FMUL2 ; FFMA2 by an opaque 1.0 == add), separate roundings
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float2 t = __fmul2_rn(make_float2(a[2 * i], a[2 * i + 1]), c2);
                t = __ffma2_rn(t, one2, one2);
                a[2 * i] = t.x; a[2 * i + 1] = t.y;
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(u[0] ^ u[1] ^ u[2] ^ u[3]);
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, int instr_per_iter, float* out, long long* cyc, int sms) {
    const int blocks = sms * 2, threads = 512;     // 32 warps / SM
    bench<MODE><<<blocks, threads>>>(out, 1.0f, 0.999f, 0u, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<MODE><<<blocks, threads>>>(out, 1.0f, 0.999f, 0u, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[4096];
    cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += (double)h[i];
    avg /= blocks;
    // per SM: 32 warps x ITERS x instr_per_iter warp-instructions in `avg` cycles
    const double wipc = 32.0 * ITERS * instr_per_iter / avg;
    printf("%-44s %8.3f ms  %10.0f cycles  %6.2f warp-instr/cycle/SM (%d counted instr/iter)\n", name, ms, avg, wipc, instr_per_iter);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * 2 * 512);
    cudaMalloc(&cyc, sizeof(long long) * 4096);
    run<0>("8 FFMA", 8, out, cyc, sms);
    run<1>("4 FFMA2 (same flops)", 4, out, cyc, sms);
    run<2>("8 FFMA + 16 int ALU", 24, out, cyc, sms);
    run<3>("4 FFMA2 + 16 int ALU", 20, out, cyc, sms);
    run<4>("8 x (FMUL ; FADD)", 16, out, cyc, sms);
    run<5>("4 x (FMUL2 ; FFMA2 by opaque 1.0)", 8, out, cyc, sms);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
