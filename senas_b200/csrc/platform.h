// platform.h -- the few macros that let the kernel sources also be compiled by the TEST-ONLY
// kernel-logic emulator (tests/emu/cpu_emu.h, -DSENAS_EMU).  The product library is always the
// nvcc/sm_100a build; the emulator build is never loaded by the senas_b200 package.
#pragma once
#include <stdint.h>

#ifdef SENAS_EMU
#include "cpu_emu.h"
#else
#include <cuda_runtime.h>
#define SENAS_LAUNCH(kern, grid, block, smem, stream, ...)                      \
  do {                                                                          \
    kern<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);     \
    ++g_launch_count;                                                           \
  } while (0)
#define SENAS_DYN_SMEM(T, name)                                       \
  extern __shared__ __align__(16) unsigned char name##_raw_[];        \
  T *name = reinterpret_cast<T *>(name##_raw_)
#define SENAS_DEVFN __device__ __forceinline__
#endif

static int64_t g_launch_count = 0;
