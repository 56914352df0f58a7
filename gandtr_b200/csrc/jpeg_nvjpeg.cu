// N1 -- JPEG decoding on the GPU for the device-side image loader: batches of JPEG bit streams handed to the
// nvJPEG library, so that the host cores no longer bound the ingest of real photo collections
// (mdir/external/cirtorch/datasets/datahelpers.py:20-27 `pil_loader` decodes on the host, one image per DataLoader worker:
// ~100 ms per 7 MP photo and core). Backends are tried in the order hardware JPEG engines -> GPU-hybrid (Huffman decoding
// on the GPU, for batches of >= 50 baseline streams) -> default. Measured on the B200 boxes of this project with nvJPEG
// 12.4: nvjpegCreateEx(NVJPEG_BACKEND_HARDWARE) answers NVJPEG_STATUS_ARCH_MISMATCH (7), so batches run on the GPU-hybrid
// backend: 146 photos/s of 3072x2304 through decode + K5 + K1 in batches of 64, against 84 with 16 PIL decoding workers
// (profiles/loader_throughput_r2u.json).
// LIBRARY code, like cuBLAS: nothing here is a kernel of ours, and the decoded pixels are NOT bit-identical to libjpeg's
// (different IDCT / upsampling arithmetic: a few grey levels); the default loader therefore keeps PIL decoding, this path
// is opt-in (`DeviceImageLoader(decode="nvjpeg")`). Everything after the decode (K5 thumbnail, K1) is ours and exact.
// libnvjpeg is opened lazily with dlopen: libgandtr_b200.so carries no link-time dependency on it, and a machine without
// it only loses these entry points (GDT_ERR_UNSUPPORTED).
#include <dlfcn.h>
#include <unistd.h>
#include <nvjpeg.h>

#include <mutex>

#include "common.cuh"

namespace gdt {

struct NvjpegApi {
    void* so = nullptr;
    decltype(&nvjpegCreateEx) CreateEx = nullptr;
    decltype(&nvjpegCreateSimple) CreateSimple = nullptr;
    decltype(&nvjpegJpegStateCreate) StateCreate = nullptr;
    decltype(&nvjpegGetImageInfo) GetImageInfo = nullptr;
    decltype(&nvjpegDecodeBatchedInitialize) BatchedInit = nullptr;
    decltype(&nvjpegDecodeBatched) Batched = nullptr;
    bool ok = false;
};

// per device: a hardware-engine decoder, the GPU-hybrid backend (Huffman decoding on the GPU for large batches of baseline
// streams) and the default backend (Huffman decoding on host threads) as fallbacks, tried in this order
struct JpegDev {
    nvjpegHandle_t h[3] = {nullptr, nullptr, nullptr};
    nvjpegJpegState_t s[3] = {nullptr, nullptr, nullptr};
    int batch[3] = {0, 0, 0};
    bool tried = false;
};

static NvjpegApi g_nvj;
static JpegDev g_jpeg[32];
static std::mutex g_jpeg_mutex;
static int g_jpeg_last_backend = 0;     // 1 = hardware engines, 2 = GPU-hybrid backend, 3 = default backend
static int g_jpeg_status[4] = {-1, -1, -1, -1};   // nvjpegStatus_t of: hardware create, hardware batched init, hardware
                                                  // decode, hybrid / default decode (gdt_debug_jpeg_status)

static bool nvjpeg_load() {
    if (g_nvj.so) return g_nvj.ok;
    const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12",
                           "/usr/local/cuda/lib64/libnvjpeg.so"};
    for (const char* n : names) {
        g_nvj.so = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (g_nvj.so) break;
    }
    if (!g_nvj.so) { g_nvj.so = (void*)1; return false; }
#define GDT_NVJ(field, sym) g_nvj.field = (decltype(g_nvj.field))dlsym(g_nvj.so, #sym)
    GDT_NVJ(CreateEx, nvjpegCreateEx);
    GDT_NVJ(CreateSimple, nvjpegCreateSimple);
    GDT_NVJ(StateCreate, nvjpegJpegStateCreate);
    GDT_NVJ(GetImageInfo, nvjpegGetImageInfo);
    GDT_NVJ(BatchedInit, nvjpegDecodeBatchedInitialize);
    GDT_NVJ(Batched, nvjpegDecodeBatched);
#undef GDT_NVJ
    g_nvj.ok = g_nvj.CreateEx && g_nvj.CreateSimple && g_nvj.StateCreate && g_nvj.GetImageInfo && g_nvj.BatchedInit && g_nvj.Batched;
    return g_nvj.ok;
}

static JpegDev* jpeg_dev() {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return nullptr;
    JpegDev& J = g_jpeg[dev];
    if (!J.tried) {
        J.tried = true;
        const nvjpegBackend_t backends[3] = {NVJPEG_BACKEND_HARDWARE, NVJPEG_BACKEND_GPU_HYBRID, NVJPEG_BACKEND_DEFAULT};
        for (int b = 0; b < 3; ++b) {
            const nvjpegStatus_t st = g_nvj.CreateEx(backends[b], nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &J.h[b]);
            if (b == 0) g_jpeg_status[0] = (int)st;
            if (st != NVJPEG_STATUS_SUCCESS || g_nvj.StateCreate(J.h[b], &J.s[b]) != NVJPEG_STATUS_SUCCESS) {
                J.h[b] = nullptr; J.s[b] = nullptr;       // e.g. no hardware JPEG engines exposed to this process
            }
        }
        cudaGetLastError();
    }
    return (J.h[0] || J.h[1] || J.h[2]) ? &J : nullptr;
}

}  // namespace gdt

using namespace gdt;

extern "C" int gdt_jpeg_available(void) {
    std::lock_guard<std::mutex> lock(g_jpeg_mutex);
    return nvjpeg_load() ? 1 : 0;
}

extern "C" int gdt_debug_jpeg_last_backend(void) { return g_jpeg_last_backend; }

extern "C" int gdt_debug_jpeg_status(int* out4) {
    if (!out4) return GDT_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < 4; ++i) out4[i] = g_jpeg_status[i];
    return GDT_OK;
}

extern "C" int gdt_jpeg_dims(const uint8_t* jpeg, size_t nbytes, int* width, int* height) {
    if (!jpeg || !nbytes || !width || !height) return GDT_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(g_jpeg_mutex);
    if (!nvjpeg_load()) return GDT_ERR_UNSUPPORTED;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return GDT_ERR_NO_DEVICE; }
    JpegDev* J = jpeg_dev();
    if (!J) return GDT_ERR_UNSUPPORTED;
    int ncomp = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t ss;
    nvjpegHandle_t any = J->h[2] ? J->h[2] : (J->h[1] ? J->h[1] : J->h[0]);
    if (g_nvj.GetImageInfo(any, jpeg, nbytes, &ncomp, &ss, ws, hs) != NVJPEG_STATUS_SUCCESS)
        return GDT_ERR_INVALID_ARGUMENT;
    *width = ws[0];
    *height = hs[0];
    return GDT_OK;
}

// jpegs[i] / nbytes[i]: host bit streams; dev_rgb[i]: device buffer of heights[i] * widths[i] * 3 bytes (interleaved RGB, what
// gdt_resize_u8 and gdt_clahe_u8 consume). A backend that cannot take the batch (no hardware engines, progressive
// streams, unusual subsampling) hands it on to the next one.
extern "C" int gdt_jpeg_decode_batch(const uint8_t* const* jpegs, const size_t* nbytes, int n, uint8_t* const* dev_rgb,
                                     const int* widths, const int* heights, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!jpegs || !nbytes || !dev_rgb || !widths || !heights || n < 0) return GDT_ERR_INVALID_ARGUMENT;
    if (n == 0) return GDT_OK;
    std::lock_guard<std::mutex> lock(g_jpeg_mutex);
    if (!nvjpeg_load()) return GDT_ERR_UNSUPPORTED;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return GDT_ERR_NO_DEVICE; }
    JpegDev* J = jpeg_dev();
    if (!J) return GDT_ERR_UNSUPPORTED;
    nvjpegImage_t* out = new (std::nothrow) nvjpegImage_t[n];
    if (!out) return GDT_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < n; ++i) {
        if (!jpegs[i] || !dev_rgb[i] || widths[i] <= 0 || heights[i] <= 0) { delete[] out; return GDT_ERR_INVALID_ARGUMENT; }
        for (int c = 0; c < NVJPEG_MAX_COMPONENT; ++c) { out[i].channel[c] = nullptr; out[i].pitch[c] = 0; }
        out[i].channel[0] = dev_rgb[i];
        out[i].pitch[0] = (size_t)widths[i] * 3;
    }
    int rc = GDT_ERR_UNSUPPORTED;
    int host_threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (host_threads < 1) host_threads = 1;
    if (host_threads > 32) host_threads = 32;
    for (int b = 0; b < 3 && rc != GDT_OK; ++b) {
        if (!J->h[b]) continue;
        // the GPU-hybrid backend pays off (and switches to GPU Huffman decoding) only for large batches
        if (b == 1 && n < 50) continue;
        bool ok = true;
        if (J->batch[b] != n) {
            const nvjpegStatus_t st = g_nvj.BatchedInit(J->h[b], J->s[b], n, host_threads, NVJPEG_OUTPUT_RGBI);
            if (b == 0) g_jpeg_status[1] = (int)st;
            ok = st == NVJPEG_STATUS_SUCCESS;
            J->batch[b] = ok ? n : 0;
        }
        if (ok) {
            const nvjpegStatus_t st = g_nvj.Batched(J->h[b], J->s[b], jpegs, nbytes, out, stream);
            g_jpeg_status[b == 0 ? 2 : 3] = (int)st;
            ok = st == NVJPEG_STATUS_SUCCESS;
        }
        if (ok) {
            rc = GDT_OK;
            g_jpeg_last_backend = b + 1;
        } else {
            J->batch[b] = 0;
            cudaGetLastError();
            rc = GDT_ERR_INVALID_ARGUMENT;       // unless a later backend takes the batch: not decodable JPEG streams
        }
    }
    delete[] out;
    return rc;
}
