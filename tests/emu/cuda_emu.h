// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A minimal single-threaded stand-in for the CUDA runtime and device builtins so
// that josefine_b200/csrc/engine.cu + raft_device.cuh (the exact device code)
// can be compiled by g++ and exercised against the oracle in the CPU test suite
// (`-m "not gpu"`).  Kernels run one thread after another; that is exact for the
// step kernel because a launch only reads the previous step's mailboxes.
// Nothing on the product path includes this file: libjosefine_b200.so is built
// by nvcc without JR_EMU and has no CPU fallback.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>

struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static

namespace jr_emu {
inline uint3& tid() { static uint3 v{0, 0, 0}; return v; }
inline uint3& bid() { static uint3 v{0, 0, 0}; return v; }
inline uint3& bdim() { static uint3 v{1, 1, 1}; return v; }
inline uint3& gdim() { static uint3 v{1, 1, 1}; return v; }
template <class F>
inline void launch(unsigned grid, unsigned block, F&& body) {
  gdim().x = grid;
  bdim().x = block;
  for (unsigned b = 0; b < grid; ++b)
    for (unsigned t = 0; t < block; ++t) {
      bid().x = b;
      tid().x = t;
      body();
    }
}
}  // namespace jr_emu
#define threadIdx (jr_emu::tid())
#define blockIdx (jr_emu::bid())
#define blockDim (jr_emu::bdim())
#define gridDim (jr_emu::gdim())
#define JR_LAUNCH(kernel, grid, block, stream, ...) \
  jr_emu::launch((unsigned)(grid), (unsigned)(block), [&]() { kernel(__VA_ARGS__); })

template <class T> inline T __ldg(const T* p) { return *p; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline void __syncthreads() {}
template <class T> inline T __shfl_down_sync(unsigned, T v, int) { return v; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = *p; *p = std::max(o, v); return o; }
using std::max;
using std::min;

// ---- runtime ---------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef void* cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
enum { cudaStreamNonBlocking = 1 };
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)aligned_alloc(64, (n + 63) / 64 * 64); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = (void*)1; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
