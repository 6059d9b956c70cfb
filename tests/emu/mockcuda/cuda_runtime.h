// Mock of the few CUDA runtime calls rodio_b200/csrc/rb_api.cu makes, on host memory (test infrastructure): with it the
// library's HOST code -- planner, session bookkeeping, FIFO accounting, state blobs -- compiles with g++ and runs on the CPU,
// the kernels behind it replaced by the SIMT emulator (tests/emu/hostemu.cpp).  "Device" pointers are host pointers; every
// allocation is registered so that the emulator can check what the kernels read.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>

typedef int cudaError_t;
typedef struct mock_stream_* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };

extern "C" void mock_cuda_register(void* p, size_t n);
extern "C" void mock_cuda_unregister(void* p);

static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "mock CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 148; return cudaSuccess; }
template <class T>
static inline cudaError_t cudaMalloc(T** p, size_t n) {
    const size_t bytes = (n + 255) / 256 * 256 + 256;           // a little slack, 256-byte aligned like cudaMalloc
    void* q = std::aligned_alloc(256, bytes);
    if (!q) return cudaErrorMemoryAllocation;
    std::memset(q, 0xCD, bytes);                                 // uninitialised device memory is not zero
    mock_cuda_register(q, bytes);
    *p = (T*)q;
    return cudaSuccess;
}
static inline cudaError_t cudaFree(void* p) {
    if (p) mock_cuda_unregister(p), std::free(p);
    return cudaSuccess;
}
template <class T>
static inline cudaError_t cudaMallocHost(T** p, size_t n) {
    *p = (T*)std::malloc(n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < h; r++) std::memmove((char*)d + r * dp, (const char*)s + r * sp, w);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
