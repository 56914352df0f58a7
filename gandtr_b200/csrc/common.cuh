// Shared host-side plumbing of libgandtr_b200.so: status codes, error capture, launch helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gandtr_b200.h"

namespace gdt {

// last CUDA error text of the calling thread (gdt_last_cuda_error)
char* tls_error_buf();

inline int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    snprintf(tls_error_buf(), 512, "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    return GDT_ERR_CUDA;
}

#define GDT_CUDA(call)                                                          \
    do {                                                                        \
        cudaError_t _e = (call);                                                \
        if (_e != cudaSuccess) return ::gdt::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define GDT_LAUNCH_CHECK()                                                      \
    do {                                                                        \
        cudaError_t _e = cudaGetLastError();                                    \
        if (_e != cudaSuccess) return ::gdt::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// bump allocator over the caller's workspace
struct Workspace {
    char* base;
    size_t size;
    size_t off;
    Workspace(void* p, size_t n) : base((char*)p), size(n), off(0) {}
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 256);
        T* r = (T*)(base + off);
        off += count * sizeof(T);
        return r;
    }
    bool ok() const { return off <= size && (((uintptr_t)base) & 255) == 0; }
};

// per-device immutable tables of the CLAHE path (clahe_sm100.cu)
struct ClaheTables;
const ClaheTables* clahe_tables_for_current_device();

int sm_count_current_device();

// index of the calling thread's current device, clamped to [0, 32): function attributes and occupancy results are per
// device, so the "already configured" flags of the launch helpers are arrays indexed by it
inline int current_device_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) dev = 0;
    return dev;
}

}  // namespace gdt
