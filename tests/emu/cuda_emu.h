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

#include <pthread.h>

#include <thread>
#include <vector>

namespace jr_emu {
struct Ctx {
  uint3 tid{0, 0, 0}, bid{0, 0, 0}, bdim{1, 1, 1}, gdim{1, 1, 1};
  void* smem = nullptr;
  pthread_barrier_t* bar = nullptr;
};
inline Ctx& ctx() { static thread_local Ctx c; return c; }
// Barrier-free kernels: one thread after another on the calling thread.
template <class F>
inline void launch(unsigned grid, unsigned block, F&& body) {
  Ctx& c = ctx();
  c.gdim.x = grid; c.bdim.x = block; c.smem = nullptr; c.bar = nullptr;
  for (unsigned b = 0; b < grid; ++b)
    for (unsigned t = 0; t < block; ++t) {
      c.bid.x = b; c.tid.x = t;
      body();
    }
}
// Kernels that use __syncthreads() / dynamic shared memory: one OS thread per CUDA
// thread of a CTA and a pthread barrier.  CTAs run one after another by default;
// JR_EMU_CTAS=k runs them k at a time (each with its own shared memory and barrier),
// which is what the ThreadSanitizer harness (tsan_split.cpp) uses to check the
// split-launch hand-over between CTAs.
inline unsigned concurrent_ctas() {
  const char* v = getenv("JR_EMU_CTAS");
  const int k = v ? atoi(v) : 1;
  return (unsigned)(k < 1 ? 1 : (k > 16 ? 16 : k));
}
template <class F>
inline void launch_coop(unsigned grid, unsigned block, size_t smem_bytes, F&& body) {
  const unsigned K = concurrent_ctas();
  for (unsigned b0 = 0; b0 < grid; b0 += K) {
    const unsigned nb = std::min(K, grid - b0);
    std::vector<void*> sm(nb);
    std::vector<pthread_barrier_t> bar(nb);
    for (unsigned i = 0; i < nb; ++i) {
      sm[i] = aligned_alloc(64, (smem_bytes + 63) / 64 * 64 + 64);
      pthread_barrier_init(&bar[i], nullptr, block);
    }
    std::vector<std::thread> th;
    th.reserve((size_t)nb * block);
    for (unsigned i = 0; i < nb; ++i)
      for (unsigned t = 0; t < block; ++t)
        th.emplace_back([&, i, t]() {
          Ctx& c = ctx();
          c.gdim.x = grid; c.bdim.x = block; c.bid.x = b0 + i; c.tid.x = t; c.smem = sm[i]; c.bar = &bar[i];
          body();
        });
    for (auto& x : th) x.join();
    for (unsigned i = 0; i < nb; ++i) {
      pthread_barrier_destroy(&bar[i]);
      free(sm[i]);
    }
  }
}
inline void sync() { if (ctx().bar) pthread_barrier_wait(ctx().bar); }
}  // namespace jr_emu
#define threadIdx (jr_emu::ctx().tid)
#define blockIdx (jr_emu::ctx().bid)
#define blockDim (jr_emu::ctx().bdim)
#define gridDim (jr_emu::ctx().gdim)
#define JR_LAUNCH(kernel, grid, block, stream, ...) \
  jr_emu::launch((unsigned)(grid), (unsigned)(block), [&]() { kernel(__VA_ARGS__); })
#define JR_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) \
  jr_emu::launch_coop((unsigned)(grid), (unsigned)(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })
#define JR_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(jr_emu::ctx().smem)

template <class T> inline T __ldg(const T* p) { return *p; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline void __syncthreads() { jr_emu::sync(); }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline T __shfl_down_sync(unsigned, T v, int) { return v; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
#ifdef __SANITIZE_THREAD__
inline void __threadfence() {}  // TSAN does not model fences; the hand-over's release store carries the ordering here
#else
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
#endif
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = *p; *p = std::max(o, v); return o; }  // barrier-free kernels only
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
enum { cudaHostAllocMapped = 2 };
inline cudaError_t cudaHostAlloc(void** p, size_t n, int) { *p = malloc(n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = (void*)1; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
typedef void* cudaEvent_t;
enum { cudaEventDisableTiming = 2 };
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, int) { *e = (void*)1; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
