// jr_cuda.h -- the one place the CUDA runtime enters the engine sources.
// JR_EMU is defined only by the CPU test harness (tests/emu/), never by the product build.
#pragma once
#ifdef JR_EMU
#include "cuda_emu.h"
#define JR_DEVICE_CODE 1
#else
#include <cuda_runtime.h>
#define JR_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define JR_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define JR_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#ifdef __CUDACC__
#define JR_DEVICE_CODE 1
#endif
#endif
