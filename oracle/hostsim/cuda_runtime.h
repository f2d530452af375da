// TEST INFRASTRUCTURE ONLY - never part of the product library.
//
// A stand-in for <cuda_runtime.h> that lets graphaligner_b200/csrc/ga_kernels.cu be compiled by g++ with -DGA_HOSTSIM:
// device memory is host memory, copies are memcpy, streams and events do nothing.  The kernels' per-stream code
// (ga_core.cuh, ga_trace.cuh) is written so that it also compiles for the host; ga_kernels.cu replaces each launch by a
// loop over streams under GA_HOSTSIM.  The result, oracle/_ref/libga_hostsim.so, exports the same C ABI as the product
// and is loaded through GA_LIB by the differential fuzzer and by the CPU tests: it checks the device ALGORITHM against
// the reference in a container that has no GPU.  It is not a fallback: the product library never loads it, and
// without a CUDA device the product's ga_create still fails.
#ifndef GA_HOSTSIM_CUDA_RUNTIME_H
#define GA_HOSTSIM_CUDA_RUNTIME_H
#ifndef GA_HOSTSIM
#error "oracle/hostsim/cuda_runtime.h is only for -DGA_HOSTSIM builds"
#endif
#include <cstdint>
#include <cstdlib>
#include <cstring>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaHostAllocPortable = 1 };
struct cudaDeviceProp { int multiProcessorCount; };

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = { x, y, z, w }; return r; }

#define __constant__
#define __host__
#define __device__
#define __restrict__

static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "hostsim error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = 148; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 1.0f; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t bytes) { return cudaMalloc((void**)p, bytes); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t bytes, int) { *p = malloc(bytes ? bytes : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T> static inline cudaError_t cudaHostAlloc(T** p, size_t bytes, int f) { return cudaHostAlloc((void**)p, bytes, f); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t) { if (bytes) memcpy(dst, src, bytes); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind) { if (bytes) memcpy(dst, src, bytes); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* dst, int v, size_t bytes, cudaStream_t) { if (bytes) memset(dst, v, bytes); return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMemcpyToSymbol(T& symbol, const void* src, size_t bytes) { memcpy(&symbol, src, bytes); return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t* freeB, size_t* totalB) { *freeB = (size_t)8 << 30; *totalB = (size_t)16 << 30; return cudaSuccess; }
#endif
