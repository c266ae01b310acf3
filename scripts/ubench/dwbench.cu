// Micro-benchmark of depthwise-convolution skeletons on B200 (16 x 32ch x 256 x 256, NHWC fp32).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I senas_b200/csrc scripts/ubench/dwbench.cu -o gpurun_out/dwbench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "kernels.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

// ---- lane = channel skeleton, one K per kernel, WT columns per thread, next input row prefetched ----
// MODE 0: statistics of z only (sum, sum of squares per channel) ; MODE 1: also store z
template <int K, int WT, int MODE, bool PREFETCH>
__global__ void __launch_bounds__(128) lane_dw_kernel(const float *x, const float *w, float *z, float *partials, int H, int W,
                                                         int tiles_x, int tile_rows) {
  constexpr int C = 32, P = K / 2, NX = WT + K - 1, T = K * K;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, n = blockIdx.y;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int c0 = (tx * 4 + warp) * WT, r0 = ty * tile_rows, r1 = min(r0 + tile_rows, H);
  float wr[T];
#pragma unroll
  for (int t = 0; t < T; ++t) wr[t] = __ldg(w + lane * T + t);
  bool cok[NX];
#pragma unroll
  for (int j = 0; j < NX; ++j) cok[j] = c0 - P + j >= 0 && c0 - P + j < W;
  float acc[K][WT];
#pragma unroll
  for (int s = 0; s < K; ++s)
#pragma unroll
    for (int j = 0; j < WT; ++j) acc[s][j] = 0.f;
  const float *xb = x + (int64_t)n * H * W * C + lane;
  float *zb = z + (int64_t)n * H * W * C + lane;
  const int R = r1 - r0, niter = R + K - 1, r_first = r0 - P;
  float s0 = 0.f, s1 = 0.f;
  float xn[NX];
  auto load_row = [&](int rr, float *dst) {
    if (rr >= 0 && rr < H) {
      const float *rowp = xb + ((int64_t)rr * W + (c0 - P)) * C;
#pragma unroll
      for (int j = 0; j < NX; ++j) dst[j] = cok[j] ? rowp[(int64_t)j * C] : 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < NX; ++j) dst[j] = 0.f;
    }
  };
  if (PREFETCH) load_row(r_first, xn);
  for (int i0 = 0; i0 < niter; i0 += K) {
#pragma unroll
    for (int u = 0; u < K; ++u) {
      const int i = i0 + u, rr = r_first + i;
      if (i < niter) {
        float xv[NX];
        if (PREFETCH) {
#pragma unroll
          for (int j = 0; j < NX; ++j) xv[j] = xn[j];
          if (i + 1 < niter) load_row(rr + 1, xn);
        } else {
          load_row(rr, xv);
        }
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          if ((unsigned)(i - ky) < (unsigned)R) {
            const int s = (u - ky + K) % K;
#pragma unroll
            for (int kx = 0; kx < K; ++kx)
#pragma unroll
              for (int j = 0; j < WT; ++j) acc[s][j] = fmaf(xv[kx + j], wr[ky * K + kx], acc[s][j]);
          }
        }
        const int sc_ = (u + 1) % K;
        if (i >= K - 1) {
          const int o = r0 + i - (K - 1);
#pragma unroll
          for (int j = 0; j < WT; ++j) {
            if (c0 + j < W) {
              const float v = acc[sc_][j];
              s0 += v, s1 += v * v;
              if (MODE == 1) zb[((int64_t)o * W + c0 + j) * C] = v;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < WT; ++j) acc[sc_][j] = 0.f;
      }
    }
  }
  __shared__ float s_red[4][2][32];
  s_red[warp][0][lane] = s0, s_red[warp][1][lane] = s1;
  __syncthreads();
  if (tid < 64) {
    const int j = tid >> 5, c = tid & 31;
    partials[((int64_t)n * gridDim.x + blockIdx.x) * 64 + tid] = (s_red[0][j][c] + s_red[1][j][c]) + (s_red[2][j][c] + s_red[3][j][c]);
  }
}

// ---- lane = channel weight gradient: dW[c][ky][kx] = sum x[o+ky-P][col+kx-P][c] * dz[o][col][c] ----
template <int K, int WT>
__global__ void __launch_bounds__(128) lane_dw_wgrad_kernel(const float *x, const float *dz, float *partials, int H, int W,
                                                               int tiles_x, int tile_rows) {
  constexpr int C = 32, P = K / 2, NX = WT + K - 1, T = K * K;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, n = blockIdx.y;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int c0 = (tx * 4 + warp) * WT, r0 = ty * tile_rows, r1 = min(r0 + tile_rows, H);
  bool cok[NX];
#pragma unroll
  for (int j = 0; j < NX; ++j) cok[j] = c0 - P + j >= 0 && c0 - P + j < W;
  float acc[T];
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t] = 0.f;
  float dzr[K][WT];  // dz rows in flight (slot = row mod K)
#pragma unroll
  for (int s = 0; s < K; ++s)
#pragma unroll
    for (int j = 0; j < WT; ++j) dzr[s][j] = 0.f;
  const float *xb = x + (int64_t)n * H * W * C + lane;
  const float *dzb = dz + (int64_t)n * H * W * C + lane;
  const int R = r1 - r0, niter = R + K - 1, r_first = r0 - P;
  for (int i0 = 0; i0 < niter; i0 += K) {
#pragma unroll
    for (int u = 0; u < K; ++u) {
      const int i = i0 + u, rr = r_first + i;
      if (i < niter) {
        // dz row entering the window: output row r0 + i (slot u)
#pragma unroll
        for (int j = 0; j < WT; ++j)
          dzr[u][j] = (i < R && c0 + j < W) ? dzb[((int64_t)(r0 + i) * W + c0 + j) * C] : 0.f;
        if (rr >= 0 && rr < H) {
          float xv[NX];
          const float *rowp = xb + ((int64_t)rr * W + (c0 - P)) * C;
#pragma unroll
          for (int j = 0; j < NX; ++j) xv[j] = cok[j] ? rowp[(int64_t)j * C] : 0.f;
#pragma unroll
          for (int ky = 0; ky < K; ++ky) {
            if ((unsigned)(i - ky) < (unsigned)R) {
              const int s = (u - ky + K) % K;
#pragma unroll
              for (int kx = 0; kx < K; ++kx)
#pragma unroll
                for (int j = 0; j < WT; ++j) acc[ky * K + kx] = fmaf(xv[kx + j], dzr[s][j], acc[ky * K + kx]);
            }
          }
        }
      }
    }
  }
  __shared__ float s_red[4][T][32];
#pragma unroll
  for (int t = 0; t < T; ++t) s_red[warp][t][lane] = acc[t];
  __syncthreads();
  for (int o = tid; o < T * 32; o += 128) {
    const int t = o >> 5, c = o & 31;
    partials[((int64_t)n * gridDim.x + blockIdx.x) * T * 32 + c * T + t] =
        (s_red[0][t][c] + s_red[1][t][c]) + (s_red[2][t][c] + s_red[3][t][c]);
  }
}

template <class F>
static float time_ms(F f, int iters = 5) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  f();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / iters;
}

int main() {
  const int B = 16, H = 256, W = 256, C = 32;
  const size_t npx = (size_t)B * H * W;
  float *x, *z3, *z5, *w3, *w5, *part;
  CK(cudaMalloc(&x, npx * C * 4)); CK(cudaMalloc(&z3, npx * C * 4)); CK(cudaMalloc(&z5, npx * C * 4));
  CK(cudaMalloc(&w3, C * 9 * 4)); CK(cudaMalloc(&w5, C * 25 * 4)); CK(cudaMalloc(&part, (size_t)64 << 20));
  std::vector<float> hx(npx * C), hw(C * 25);
  for (auto &v : hx) v = (float)rand() / RAND_MAX - 0.5f;
  for (auto &v : hw) v = (float)rand() / RAND_MAX - 0.5f;
  CK(cudaMemcpy(x, hx.data(), npx * C * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(w3, hw.data(), C * 9 * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(w5, hw.data(), C * 25 * 4, cudaMemcpyHostToDevice));
  const double gf5 = 2.0 * npx * C * 25 / 1e9, gf3 = 2.0 * npx * C * 9 / 1e9;
  // baseline: dw_multi_kernel<32, true>, k3 + k5 of one edge in one launch (writes z, statistics)
  {
    DwMultiArgs a;
    memset(&a, 0, sizeof(a));
    a.n = 2, a.H = H, a.W = W;
    a.tiles_x = (W + 2 * 16 - 1) / (2 * 16), a.tile_rows = 32;
    const int nblk = a.tiles_x * ((H + a.tile_rows - 1) / a.tile_rows);
    for (int m = 0; m < 2; ++m) {
      a.it[m].in = x, a.it[m].in_ld = C, a.it[m].out = m ? z5 : z3, a.it[m].out_ld = C, a.it[m].w = m ? w5 : w3;
      a.it[m].partials = part + (size_t)m * (8 << 20), a.it[m].k = m ? 5 : 3;
    }
    auto kern = dw_multi_kernel<32, true>;
    float ms = time_ms([&] { kern<<<dim3(nblk, B), 128>>>(a); });
    printf("dw_multi<32,stats> k3+k5 (writes z)        %7.3f ms  %6.2f TFLOP/s  %7.1f GB/s (x + 2z)\n", ms, (gf3 + gf5) / ms, 3.0 * npx * C * 4 / ms / 1e6);
  }
#define RUN(K_, WT_, MODE_, PF_, ROWS_)                                                                                   \
  {                                                                                                                        \
    const int tiles_x = (W + 4 * WT_ - 1) / (4 * WT_), nblk = tiles_x * ((H + ROWS_ - 1) / ROWS_);                        \
    auto kern = lane_dw_kernel<K_, WT_, MODE_, PF_>;                                                                       \
    float ms = time_ms([&] { kern<<<dim3(nblk, B), 128>>>(x, K_ == 5 ? w5 : w3, z5, part, H, W, tiles_x, ROWS_); });       \
    cudaFuncAttributes fa;                                                                                                 \
    cudaFuncGetAttributes(&fa, kern);                                                                                      \
    printf("lane_dw K=%d WT=%d mode=%d prefetch=%d rows=%d regs=%d  %7.3f ms  %6.2f TFLOP/s\n", K_, WT_, MODE_, (int)PF_, ROWS_, \
           fa.numRegs, ms, (K_ == 5 ? gf5 : gf3) / ms);                                                                    \
  }
  RUN(5, 4, 0, false, 32)
  RUN(5, 4, 0, true, 32)
  RUN(5, 8, 0, false, 32)
  RUN(5, 8, 0, true, 32)
  RUN(5, 8, 0, true, 64)
  RUN(5, 4, 1, true, 32)
  RUN(5, 8, 1, true, 32)
  RUN(3, 4, 0, false, 32)
  RUN(3, 4, 0, true, 32)
  RUN(3, 8, 0, true, 32)
  RUN(3, 8, 1, true, 32)
  // weight gradient baseline: dw_wgrad_multi_kernel<32> (k3 + k5)
  {
    DwMultiArgs a;
    memset(&a, 0, sizeof(a));
    a.n = 2, a.H = H, a.W = W;
    a.tiles_x = (W + 16 - 1) / 16, a.tile_rows = 32;
    const int nblk = a.tiles_x * ((H + a.tile_rows - 1) / a.tile_rows);
    for (int m = 0; m < 2; ++m) {
      a.it[m].in = x, a.it[m].in_ld = C, a.it[m].in2 = m ? z5 : z3, a.it[m].w = m ? w5 : w3;
      a.it[m].partials = part + (size_t)m * (8 << 20), a.it[m].k = m ? 5 : 3;
    }
    auto kern = dw_wgrad_multi_kernel<32>;
    float ms = time_ms([&] { kern<<<dim3(nblk, B), 128>>>(a); });
    printf("dw_wgrad_multi<32> k3+k5                   %7.3f ms  %6.2f TFLOP/s  %7.1f GB/s (x + 2dz)\n", ms, (gf3 + gf5) / ms, 3.0 * npx * C * 4 / ms / 1e6);
  }
#define RUNW(K_, WT_, ROWS_)                                                                                     \
  {                                                                                                               \
    const int tiles_x = (W + 4 * WT_ - 1) / (4 * WT_), nblk = tiles_x * ((H + ROWS_ - 1) / ROWS_);               \
    auto kern = lane_dw_wgrad_kernel<K_, WT_>;                                                                    \
    float ms = time_ms([&] { kern<<<dim3(nblk, B), 128>>>(x, K_ == 5 ? z5 : z3, part, H, W, tiles_x, ROWS_); });  \
    cudaFuncAttributes fa;                                                                                        \
    cudaFuncGetAttributes(&fa, kern);                                                                             \
    printf("lane_dw_wgrad K=%d WT=%d rows=%d regs=%d  %7.3f ms  %6.2f TFLOP/s\n", K_, WT_, ROWS_, fa.numRegs, ms, \
           (K_ == 5 ? gf5 : gf3) / ms);                                                                           \
  }
  RUNW(5, 4, 32)
  RUNW(5, 8, 32)
  RUNW(5, 4, 64)
  RUNW(3, 4, 32)
  RUNW(3, 8, 32)
  return 0;
}
