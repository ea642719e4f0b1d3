// Stand-in for <cuda_runtime.h> when the library's CUDA sources are compiled FOR THE CPU by
// tests/emu/build_emu_lib.py (test infrastructure only).  "Device" memory is host memory, every
// stream is synchronous (a launch returns when its last CTA has run, see cuda_emu.h), events are
// host timestamps.  Only what consel_b200/csrc/{capi,capi_stats,stats_kernels,resample_kernels}.cu
// and cgpu_internal.h use is provided.
#pragma once
#include <math.h>

#include <cmath>
using std::isfinite;

#include <chrono>
#include <cstdlib>
#include <cstring>

#include "../cuda_emu.h"

struct int2 {
    int x, y;
};
static inline int2 make_int2(int x, int y) { return int2{x, y}; }

enum cudaError_t { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4 };
struct EmuStream {
    int id;
};
typedef EmuStream *cudaStream_t;
struct EmuEvent {
    double t;
};
typedef EmuEvent *cudaEvent_t;
typedef void *cudaMemPool_t;
struct cudaDeviceProp {
    int major = 10, minor = 0, multiProcessorCount = 148;
    char name[64] = "cpu emulation";
};

static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { *p = cudaDeviceProp(); return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = new EmuStream{1}; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new EmuStream{2}; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t *p, int) { *p = nullptr; return cudaSuccess; }
static inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void *) { return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = *t = (size_t)64 << 30; return cudaSuccess; }
// every allocation is exactly sized (plus nothing) and 256-byte aligned like cudaMalloc's
static inline cudaError_t cudaMalloc(void **p, size_t bytes) {
    *p = aligned_alloc(256, bytes ? (bytes + 255) / 256 * 256 : 256);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <class T>
static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return cudaMalloc(reinterpret_cast<void **>(p), bytes); }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocAsync(void **p, size_t bytes, cudaStream_t) { return cudaMalloc(p, bytes); }
static inline cudaError_t cudaFreeAsync(void *p, cudaStream_t) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height,
                                            cudaMemcpyKind, cudaStream_t = nullptr) {
    for (size_t r = 0; r < height; r++) memmove(static_cast<char *>(d) + r * dpitch, static_cast<const char *>(s) + r * spitch, width);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline double emu_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new EmuEvent{emu_now_ms()}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { if (e) e->t = emu_now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
template <class K>
static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
