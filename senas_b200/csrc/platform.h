// platform.h -- the few macros that let the kernel sources also be compiled by the TEST-ONLY
// kernel-logic emulator (tests/emu/cpu_emu.h, -DSENAS_EMU).  The product library is always the
// nvcc/sm_100a build; the emulator build is never loaded by the senas_b200 package.
#pragma once
#include <stdint.h>

static int64_t g_launch_count = 0;
// tag of the next launch: kernel family name + its ALGORITHMIC flops / bytes
static const char *g_tag = "other";
static double g_tag_flops = 0, g_tag_bytes = 0;
#define SENAS_TAG(name, fl, by) (g_tag = (name), g_tag_flops = (double)(fl), g_tag_bytes = (double)(by))

#ifdef SENAS_EMU
#include "cpu_emu.h"
#else
#include <cuda_runtime.h>

#include <vector>
// optional per-launch timing (senas_profile): CUDA events around every tagged launch, on the launch stream
struct ProfRec {
  const char *name;
  double flops, bytes;
  cudaEvent_t e0, e1;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static inline int prof_begin(const char *name, double flops, double bytes, void *stream) {
  ProfRec r;
  r.name = name, r.flops = flops, r.bytes = bytes;
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, (cudaStream_t)stream);
  g_prof.push_back(r);
  return (int)g_prof.size() - 1;
}
// first launch that the runtime refused (bad configuration, too much shared memory ...): reported by check_cuda() with
// the family tag of the kernel, so that a launch that never ran cannot pass silently
static cudaError_t g_launch_err = cudaSuccess;
static const char *g_launch_err_tag = "";
#define SENAS_LAUNCH(kern, grid, block, smem, stream, ...)                                  \
  do {                                                                                      \
    const int pr_ = g_prof_on ? prof_begin(g_tag, g_tag_flops, g_tag_bytes, (stream)) : -1; \
    kern<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);                 \
    {                                                                                       \
      const cudaError_t le_ = cudaPeekAtLastError();                                        \
      if (le_ != cudaSuccess && g_launch_err == cudaSuccess) g_launch_err = le_, g_launch_err_tag = g_tag; \
    }                                                                                       \
    if (pr_ >= 0) cudaEventRecord(g_prof[pr_].e1, (cudaStream_t)(stream));                  \
    ++g_launch_count;                                                                       \
    g_tag = "other", g_tag_flops = g_tag_bytes = 0;                                         \
  } while (0)
#define SENAS_DYN_SMEM(T, name)                                       \
  extern __shared__ __align__(16) unsigned char name##_raw_[];        \
  T *name = reinterpret_cast<T *>(name##_raw_)
#define SENAS_DEVFN __device__ __forceinline__
#endif


