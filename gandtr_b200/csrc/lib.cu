// Library-level entry points of libgandtr_b200.so: ABI version, status strings, error capture.
#include "common.cuh"

namespace gdt {

char* tls_error_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

int sm_count_current_device() {
    static int cached[32] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace gdt

extern "C" int gdt_abi_version(void) { return GDT_ABI_VERSION; }

extern "C" const char* gdt_status_string(int status) {
    switch (status) {
        case GDT_OK: return "ok";
        case GDT_ERR_INVALID_ARGUMENT: return "invalid argument";
        case GDT_ERR_NOT_INITIALISED: return "gdt_init has not been called on this device";
        case GDT_ERR_NO_DEVICE: return "no CUDA device";
        case GDT_ERR_WORKSPACE_TOO_SMALL: return "workspace too small or misaligned";
        case GDT_ERR_CUDA: return "CUDA error (see gdt_last_cuda_error)";
        case GDT_ERR_UNSUPPORTED: return "unsupported shape or configuration";
        case GDT_ERR_CANDIDATE_OVERFLOW: return "score_topk candidate overflow";
        default: return "unknown status";
    }
}

extern "C" const char* gdt_last_cuda_error(void) { return gdt::tls_error_buf(); }
