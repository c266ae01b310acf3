// cpu_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A minimal single-OS-thread emulator of the CUDA execution model (grid of blocks, threads of a
// block as cooperative fibers, __syncthreads / warp shuffles as rendezvous points, static and
// dynamic shared memory) so that the *logic* of the kernels in senas_b200/csrc can be checked
// against the oracle in a container without a GPU.  It is compiled only into
// tests/emu/libsenas_emu.so by tests/emu/build_emu.py and is never loaded by the senas_b200
// package: the product path is the nvcc build for sm_100a and fails loudly without it.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) float4 {
  float x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

typedef void *cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
#define SENAS_DEVFN static inline
#define __ldg(p) (*(p))

static dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace emu {
constexpr int kMaxThreads = 1024;
constexpr size_t kStack = 256 * 1024;

struct Fiber {
  ucontext_t ctx;
  bool done;
};
struct BlockState {
  Fiber fib[kMaxThreads];
  char *stacks = nullptr;
  ucontext_t sched;
  int cur = 0, nthreads = 0, live = 0;
  int bar_arrived = 0;
  unsigned bar_gen = 0;
  int warp_arrived[kMaxThreads / 32 + 1];
  unsigned warp_gen[kMaxThreads / 32 + 1];
  uint32_t slot[2][kMaxThreads];
  std::function<void()> body;
};
static BlockState S;
alignas(16) static unsigned char dyn_smem[232 * 1024];

static inline void set_tid(int t) {
  threadIdx.x = t % blockDim.x;
  threadIdx.y = (t / blockDim.x) % blockDim.y;
  threadIdx.z = t / (blockDim.x * blockDim.y);
}
static inline void yield() {
  int me = S.cur;
  swapcontext(&S.fib[me].ctx, &S.sched);
  set_tid(me);
}
static void trampoline() {
  S.body();
  S.fib[S.cur].done = true;
  swapcontext(&S.fib[S.cur].ctx, &S.sched);
}
static inline void run_block(const std::function<void()> &body) {
  S.nthreads = blockDim.x * blockDim.y * blockDim.z;
  if (S.nthreads > kMaxThreads || S.nthreads <= 0) {
    fprintf(stderr, "emu: bad block size %d\n", S.nthreads);
    abort();
  }
  if (!S.stacks) S.stacks = (char *)malloc(kStack * kMaxThreads);
  S.body = body;
  S.live = S.nthreads;
  S.bar_arrived = 0;
  for (int w = 0; w < (S.nthreads + 31) / 32; ++w) S.warp_arrived[w] = 0;
  for (int t = 0; t < S.nthreads; ++t) {
    getcontext(&S.fib[t].ctx);
    S.fib[t].ctx.uc_stack.ss_sp = S.stacks + kStack * t;
    S.fib[t].ctx.uc_stack.ss_size = kStack;
    S.fib[t].ctx.uc_link = nullptr;
    S.fib[t].done = false;
    makecontext(&S.fib[t].ctx, trampoline, 0);
  }
  while (S.live > 0) {
    int before = S.live;
    bool progressed = false;
    unsigned gen0 = S.bar_gen;
    for (int t = 0; t < S.nthreads; ++t) {
      if (S.fib[t].done) continue;
      S.cur = t;
      set_tid(t);
      swapcontext(&S.sched, &S.fib[t].ctx);
      if (S.fib[t].done) {
        --S.live;
        progressed = true;
        if (S.live > 0 && S.bar_arrived == S.live) {  // exited threads no longer hold the barrier
          S.bar_arrived = 0;
          ++S.bar_gen;
        }
      }
    }
    (void)before;
    (void)progressed;
    (void)gen0;
  }
}
template <class F>
static inline void launch(dim3 grid, dim3 block, F &&body) {
  gridDim = grid;
  blockDim = block;
  std::function<void()> fn = body;
  for (unsigned z = 0; z < grid.z; ++z)
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x = 0; x < grid.x; ++x) {
        blockIdx = dim3(x, y, z);
        run_block(fn);
      }
}
static inline int flat_tid() { return S.cur; }
template <class T>
static inline T warp_exchange(T v, int src_lane_of_me(int lane, int arg), int arg) {
  static_assert(sizeof(T) == 4, "4-byte shuffles only");
  int t = flat_tid(), w = t / 32, lane = t % 32;
  unsigned g = S.warp_gen[w];
  uint32_t bits;
  memcpy(&bits, &v, 4);
  S.slot[g & 1][t] = bits;
  const int lanes = S.nthreads - w * 32 < 32 ? S.nthreads - w * 32 : 32;
  if (++S.warp_arrived[w] == lanes) {
    S.warp_arrived[w] = 0;
    ++S.warp_gen[w];
  } else {
    while (S.warp_gen[w] == g) yield();
  }
  int src = src_lane_of_me(lane, arg);
  uint32_t r = S.slot[g & 1][w * 32 + src];
  T out;
  memcpy(&out, &r, 4);
  return out;
}
static inline int lane_xor(int lane, int m) { return lane ^ m; }
static inline int lane_down(int lane, int d) { return lane + d < 32 ? lane + d : lane; }
static inline int lane_idx(int, int s) { return s & 31; }
}  // namespace emu

static inline void __syncthreads() {
  using namespace emu;
  unsigned g = S.bar_gen;
  if (++S.bar_arrived == S.live) {
    S.bar_arrived = 0;
    ++S.bar_gen;
  } else {
    while (S.bar_gen == g) yield();
  }
}
static inline void __syncwarp(unsigned = 0xffffffffu) { (void)emu::warp_exchange<int>(0, emu::lane_xor, 0); }
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int m) {
  return emu::warp_exchange<T>(v, emu::lane_xor, m);
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, int d) {
  return emu::warp_exchange<T>(v, emu::lane_down, d);
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int s) {
  return emu::warp_exchange<T>(v, emu::lane_idx, s);
}
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }


#define SENAS_LAUNCH(kern, grid, block, smem, stream, ...)                  \
  do {                                                                      \
    (void)(stream);                                                         \
    if ((size_t)(smem) > sizeof(emu::dyn_smem)) {                           \
      fprintf(stderr, "emu: dynamic smem %zu too large\n", (size_t)(smem)); \
      abort();                                                              \
    }                                                                       \
    emu::launch((grid), (block), [=]() { kern(__VA_ARGS__); });             \
    ++g_launch_count;                                                       \
    g_tag = "other";                                                        \
  } while (0)
#define SENAS_DYN_SMEM(T, name) T *name = reinterpret_cast<T *>(emu::dyn_smem)

static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) {
  memset(p, v, n);
  return cudaSuccess;
}
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
