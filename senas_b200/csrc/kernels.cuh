// kernels.cuh -- fp32 (exact-mode) device code of the SENAS MixedOp / Cell hot path.
//
// Layout: activations NHWC fp32; `ld` = floats between consecutive pixels.  Every candidate of a
// MixedOp (utils/operations.py:8-21) produces its *pre-BatchNorm* 8-channel output y_k once, with
// the per-channel batch statistics reduced in the producer's epilogue (two-stage, fixed order, no
// float atomics => bit-reproducible).  The BatchNorm affine, the SE gate, the softmax(alpha)
// weight, the edge beta, the node sum, the ReLU and the concat are then ONE streaming kernel per
// node (node_combine_kernel) that writes the node's channel slice of the concat buffer; no
// post-BN / per-candidate weighted tensor ever reaches HBM.  Backward mirrors it: one statistics
// sweep per node (S1 = sum g, S2_k = sum g*yhat_k, which also yield d alpha and d beta), then the
// data/weight gradient kernels read dy_k = A*g + B*y_k + C on the fly.
#pragma once
#include "platform.h"

#define SENAS_EPS 1e-5f
#define SENAS_MOMENTUM 0.1f

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
SENAS_DEVFN float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
SENAS_DEVFN void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
SENAS_DEVFN float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// bf16 storage of the dep-sep intermediate (z, then dz in place) in bf16 mode: 64 instead of 128 bytes per pixel for the
// tensor that dominates the traffic of the depthwise-separable candidates.  Round to nearest even, written with integer
// arithmetic so that the emulator build (tests/emu) and the device agree bit for bit.  `bf` is uniform per launch.
#ifdef SENAS_EMU
static inline uint32_t senas_f_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float senas_bits_f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#else
SENAS_DEVFN uint32_t senas_f_bits(float f) { return __float_as_uint(f); }
SENAS_DEVFN float senas_bits_f(uint32_t u) { return __uint_as_float(u); }
#endif
struct alignas(8) senas_bf16x4 {
  uint32_t lo, hi;
};
SENAS_DEVFN uint32_t senas_f2bf(float f) {
  uint32_t u = senas_f_bits(f);
  u += 0x7fffu + ((u >> 16) & 1u);
  return u >> 16;
}
SENAS_DEVFN uint32_t senas_pack_bf2(float a, float b) { return senas_f2bf(a) | (senas_f2bf(b) << 16); }
// element e of a tensor stored as fp32 (bf == 0) or bf16 (bf != 0); e is a multiple of 4 elements
SENAS_DEVFN float4 ldx4(const float *base, int64_t e, int bf) {
  if (bf) {
    const senas_bf16x4 r = *reinterpret_cast<const senas_bf16x4 *>(reinterpret_cast<const uint16_t *>(base) + e);
    return make_float4(senas_bits_f(r.lo << 16), senas_bits_f(r.lo & 0xffff0000u), senas_bits_f(r.hi << 16),
                       senas_bits_f(r.hi & 0xffff0000u));
  }
  return ld4(base + e);
}
SENAS_DEVFN void stx4(float *base, int64_t e, float4 v, int bf) {
  if (bf) {
    senas_bf16x4 r;
    r.lo = senas_pack_bf2(v.x, v.y), r.hi = senas_pack_bf2(v.z, v.w);
    *reinterpret_cast<senas_bf16x4 *>(reinterpret_cast<uint16_t *>(base) + e) = r;
  } else {
    st4(base + e, v);
  }
}
// the value a bf16 store keeps (statistics must describe what the consumers will read)
SENAS_DEVFN float4 roundx4(float4 v, int bf) {
  if (!bf) return v;
  return make_float4(senas_bits_f(senas_f2bf(v.x) << 16), senas_bits_f(senas_f2bf(v.y) << 16),
                     senas_bits_f(senas_f2bf(v.z) << 16), senas_bits_f(senas_f2bf(v.w) << 16));
}

SENAS_DEVFN float warp_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

// Sum NV per-thread values over the whole block (fixed order) and store them to dst[0..NV).
// Must be called by every thread of the block.  blockDim.x <= 1024, multiple of 32.
template <int NV>
SENAS_DEVFN void block_sum_store(const float *v, float *dst) {
  __shared__ float s_part[32][NV];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
  __syncthreads();  // protect s_part against a previous call
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float r = warp_sum(v[i]);
    if (lane == 0) s_part[warp][i] = r;
  }
  __syncthreads();
  if (tid < NV) {
    float r = 0.f;
    for (int w = 0; w < nwarp; ++w) r += s_part[w][tid];
    dst[tid] = r;
  }
}

// Tap table of one convolution variant expressed on a "base grid" (see graph.cu make_taps()):
//   gathered pixel = base * si + (dy, dx)     produced pixel = base * so + phase
struct TapTable {
  int32_t n, nphase;
  int32_t pstart[5];
  int32_t min_dy, max_dy, min_dx, max_dx;
  int8_t dy[25], dx[25], widx[25], phase[25];
};

// ------------------------------------------------------------------------------------------------
// gather-MAC: dense / dilated / strided / transposed convolution forward AND data gradient.
//   dst[b*so+ph][n] (+)= sum_{t in ph} sum_k src[b*si + d_t][k] * W[t][k][n]
// forward: src = x (KC = c_in), dst = y (NC = 8), statistics epilogue.
// dgrad  : src = dy = A*gm + B*y + C (KC = 8), dst = dx (NC = c_in), optional accumulate.
// One thread per base pixel of an 8x16 tile; gathered tile + weights staged in shared memory in
// 8-channel chunks.
// ------------------------------------------------------------------------------------------------
struct GatherArgs {
  const float *src;
  int64_t src_ld;
  int32_t src_h, src_w;
  const float *src2;  // non-null => affine dy mode (src = gm with ld 8, src2 = y)
  int64_t src2_ld;
  const float *coefA, *coefB, *coefC;  // [B][8]
  float *dst;
  int64_t dst_ld;
  int32_t dst_h, dst_w;
  int32_t accumulate;
  int32_t base_h, base_w, si, so, tiles_x;
  const float *w;
  int32_t ws_t, ws_k, ws_n;
  float *partials;  // [B][tiles][16] or null
  TapTable taps;
};

constexpr int kTileH = 8, kTileW = 16, kTileThreads = 128;

// PIX base pixels per thread (columns tx, tx + 16, ...): every weight float4 read from shared memory feeds PIX x 4 FMAs
// instead of 4 (the PIX = 1 version issued 18 LDS per 64 FMA and ran at ~15 % of the FMA peak); lanes keep consecutive
// pixels, so the activation reads stay conflict-free.
template <int KC, int NC, int NPH, int PIX>
__global__ void __launch_bounds__(kTileThreads) gather_mac_kernel(GatherArgs a) {
  constexpr int NCH = KC / 8;
  constexpr bool kPhaseOuter = (NCH == 1);  // single chunk: one live accumulator set
  constexpr int NLIVE = kPhaseOuter ? 1 : NPH;
  constexpr int TW = kTileW * PIX;
  SENAS_DYN_SMEM(float4, smem);
  __shared__ float s_coef[24];
  const int tid = threadIdx.x, n = blockIdx.y, tile = blockIdx.x;
  const int by0 = (tile / a.tiles_x) * kTileH, bx0 = (tile % a.tiles_x) * TW;
  const int ty = tid / kTileW, tx = tid % kTileW;
  const int R = (kTileH - 1) * a.si + (a.taps.max_dy - a.taps.min_dy) + 1;
  const int Cc = (TW - 1) * a.si + (a.taps.max_dx - a.taps.min_dx) + 1;
  const int npx = R * Cc;
  float4 *s_lo = smem, *s_hi = smem + npx;
  float *s_w = reinterpret_cast<float *>(smem + 2 * npx);
  const bool affine = a.src2 != nullptr;
  if (affine) {
    if (tid < 24) {
      const float *t = tid < 8 ? a.coefA : (tid < 16 ? a.coefB : a.coefC);
      s_coef[tid] = t[n * 8 + (tid & 7)];
    }
    __syncthreads();
  }
  float acc[PIX][NLIVE][NC];
#pragma unroll
  for (int q = 0; q < PIX; ++q)
#pragma unroll
    for (int p = 0; p < NLIVE; ++p)
#pragma unroll
      for (int i = 0; i < NC; ++i) acc[q][p][i] = 0.f;
  float st_s[8], st_q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) st_s[i] = st_q[i] = 0.f;

  const int gy0 = by0 * a.si + a.taps.min_dy, gx0 = bx0 * a.si + a.taps.min_dx;
  const float *srcn = a.src + (int64_t)n * a.src_h * a.src_w * a.src_ld;
  const float *src2n = affine ? a.src2 + (int64_t)n * a.src_h * a.src_w * a.src2_ld : nullptr;
  const int by = by0 + ty;
  float *dstn = a.dst + (int64_t)n * a.dst_h * a.dst_w * a.dst_ld;

  auto epilogue = [&](int q, int ph, float *r) {
    const int bx = bx0 + tx + q * kTileW;
    const int oy = by * a.so + (ph >> 1), ox = bx * a.so + (ph & 1);
    if (by < a.base_h && bx < a.base_w && oy < a.dst_h && ox < a.dst_w) {
      float *o = dstn + ((int64_t)oy * a.dst_w + ox) * a.dst_ld;
#pragma unroll
      for (int j = 0; j < NC; j += 4) {
        float4 v = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        if (a.accumulate) {
          float4 u = ld4(o + j);
          v.x += u.x, v.y += u.y, v.z += u.z, v.w += u.w;
        }
        st4(o + j, v);
      }
      if (NC == 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) st_s[j] += r[j], st_q[j] += r[j] * r[j];
      }
    }
  };

  for (int ch = 0; ch < NCH; ++ch) {
    if (ch > 0) __syncthreads();
    for (int i = tid; i < npx; i += kTileThreads) {
      const int r = i / Cc, c = i - r * Cc;
      const int gy = gy0 + r, gx = gx0 + c;
      float4 lo = f4zero(), hi = f4zero();
      if (gy >= 0 && gy < a.src_h && gx >= 0 && gx < a.src_w) {
        const int64_t pix = (int64_t)gy * a.src_w + gx;
        const float *p = srcn + pix * a.src_ld + ch * 8;
        lo = ld4(p), hi = ld4(p + 4);
        if (affine) {
          const float *q = src2n + pix * a.src2_ld;
          const float4 ylo = ld4(q), yhi = ld4(q + 4);
          lo.x = s_coef[0] * lo.x + s_coef[8] * ylo.x + s_coef[16];
          lo.y = s_coef[1] * lo.y + s_coef[9] * ylo.y + s_coef[17];
          lo.z = s_coef[2] * lo.z + s_coef[10] * ylo.z + s_coef[18];
          lo.w = s_coef[3] * lo.w + s_coef[11] * ylo.w + s_coef[19];
          hi.x = s_coef[4] * hi.x + s_coef[12] * yhi.x + s_coef[20];
          hi.y = s_coef[5] * hi.y + s_coef[13] * yhi.y + s_coef[21];
          hi.z = s_coef[6] * hi.z + s_coef[14] * yhi.z + s_coef[22];
          hi.w = s_coef[7] * hi.w + s_coef[15] * yhi.w + s_coef[23];
        }
      }
      s_lo[i] = lo, s_hi[i] = hi;
    }
    for (int i = tid; i < a.taps.n * 8 * NC; i += kTileThreads) {
      const int nn = i % NC, kk = (i / NC) & 7, t = i / (NC * 8);
      s_w[i] = __ldg(a.w + (int64_t)a.taps.widx[t] * a.ws_t + (int64_t)(ch * 8 + kk) * a.ws_k + (int64_t)nn * a.ws_n);
    }
    __syncthreads();
#pragma unroll
    for (int ph = 0; ph < NPH; ++ph) {
      const int pl = kPhaseOuter ? 0 : ph;
      if (kPhaseOuter && NPH > 1) {
#pragma unroll
        for (int q = 0; q < PIX; ++q)
#pragma unroll
          for (int i = 0; i < NC; ++i) acc[q][0][i] = 0.f;
      }
      for (int t = a.taps.pstart[ph]; t < a.taps.pstart[ph + 1]; ++t) {
        const int idx = (ty * a.si + a.taps.dy[t] - a.taps.min_dy) * Cc + (tx * a.si + a.taps.dx[t] - a.taps.min_dx);
        float xv[PIX][8];
#pragma unroll
        for (int q = 0; q < PIX; ++q) {
          const float4 lo = s_lo[idx + q * kTileW * a.si], hi = s_hi[idx + q * kTileW * a.si];
          xv[q][0] = lo.x, xv[q][1] = lo.y, xv[q][2] = lo.z, xv[q][3] = lo.w;
          xv[q][4] = hi.x, xv[q][5] = hi.y, xv[q][6] = hi.z, xv[q][7] = hi.w;
        }
        const float *wt = s_w + t * 8 * NC;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
          for (int j = 0; j < NC; j += 4) {
            const float4 w4 = ld4(wt + kk * NC + j);
#pragma unroll
            for (int q = 0; q < PIX; ++q) {
              float *r = acc[q][pl];
              r[j] = fmaf(xv[q][kk], w4.x, r[j]);
              r[j + 1] = fmaf(xv[q][kk], w4.y, r[j + 1]);
              r[j + 2] = fmaf(xv[q][kk], w4.z, r[j + 2]);
              r[j + 3] = fmaf(xv[q][kk], w4.w, r[j + 3]);
            }
          }
        }
      }
      if (kPhaseOuter) {
#pragma unroll
        for (int q = 0; q < PIX; ++q) epilogue(q, ph, acc[q][0]);
      }
    }
  }
  if (!kPhaseOuter) {
#pragma unroll
    for (int q = 0; q < PIX; ++q)
#pragma unroll
      for (int ph = 0; ph < NPH; ++ph) epilogue(q, ph, acc[q][ph]);
  }
  if (a.partials != nullptr) {  // uniform
    float v[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = st_s[j], v[8 + j] = st_q[j];
    block_sum_store<16>(v, a.partials + ((int64_t)n * gridDim.x + blockIdx.x) * 16);
  }
}


// ------------------------------------------------------------------------------------------------
// gather-MMA: the same tap-table convolution (forward and data gradient) as gather_mac_kernel, with the inner product on
// the tensor cores: warp-level mma.sync m16n8k8, TF32 operands (10-bit mantissa, rounded to nearest when the tile is
// staged), fp32 accumulation.  bf16 mode only (gate 2e-2); it takes every dense / dilated / strided / transposed
// convolution that the tcgen05 path does not: the 8 -> 8 node edges (K = 8 per tap is exactly one k8 step, no channel
// padding) and all maps that are not a multiple of 64 pixels wide (M = 16-pixel row segments, no strip constraint).
// Tile, staging (8-channel chunks as float4 lo / hi per gathered pixel) and tap tables are gather_mac's; a warp owns two
// tile rows = 2 * PIX m16 tiles.  Fragments (PTX ISA, m16n8k8 .tf32): g = lane >> 2, t = lane & 3
//   A: a0 = (row g, k t)  a1 = (row g + 8, k t)  a2 = (row g, k t + 4)  a3 = (row g + 8, k t + 4)     row = pixel
//   B: b0 = (k t, n g)    b1 = (k t + 4, n g)                                                          n = out channel
//   D: d0 = (row g, n 2t) d1 = (row g, n 2t + 1) d2 = (row g + 8, n 2t) d3 = (row g + 8, n 2t + 1)
// A reads: lane (g, t) takes component t of pixel g's float4 -> 32 lanes cover 128 contiguous bytes (stride 1).
// ------------------------------------------------------------------------------------------------
#ifdef SENAS_EMU
static inline float senas_tf32(float v) { return v; }  // the emulator checks the indexing in exact arithmetic
static inline void senas_mma_tf32(float (&d)[4], const float (&a)[4], const float (&b)[2]) {
  const int lane = emu::flat_tid() % 32, g = lane >> 2, t = lane & 3;
  float add[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < 8; ++k) {
    const float a_lo = __shfl_sync(0xffffffffu, k < 4 ? a[0] : a[2], g * 4 + (k & 3));   // A[g][k]
    const float a_hi = __shfl_sync(0xffffffffu, k < 4 ? a[1] : a[3], g * 4 + (k & 3));   // A[g + 8][k]
    const float b_0 = __shfl_sync(0xffffffffu, k < 4 ? b[0] : b[1], (2 * t) * 4 + (k & 3));      // B[k][2t]
    const float b_1 = __shfl_sync(0xffffffffu, k < 4 ? b[0] : b[1], (2 * t + 1) * 4 + (k & 3));  // B[k][2t + 1]
    add[0] += a_lo * b_0, add[1] += a_lo * b_1, add[2] += a_hi * b_0, add[3] += a_hi * b_1;
  }
  for (int i = 0; i < 4; ++i) d[i] += add[i];
}
#else
SENAS_DEVFN float senas_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
SENAS_DEVFN void senas_mma_tf32(float (&d)[4], const float (&a)[4], const float (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
        "r"(__float_as_uint(b[0])), "r"(__float_as_uint(b[1])));
}
#endif

template <int NC>
struct GatherMma {
  static constexpr int NCP = NC == 8 ? 8 : NC + 8;  // weight row stride in shared memory (conflict-free B fragments)
};

template <int KC, int NC, int NPH, int PIX>
__global__ void __launch_bounds__(kTileThreads) gather_mma_kernel(GatherArgs a) {
  constexpr int NCH = KC / 8, NT = NC / 8, MT = 2 * PIX, NCP = GatherMma<NC>::NCP;
  constexpr bool kPhaseOuter = (NCH == 1);
  constexpr int NLIVE = kPhaseOuter ? 1 : NPH;
  constexpr int TW = kTileW * PIX;
  SENAS_DYN_SMEM(float4, smem);
  __shared__ float s_coef[24];
  __shared__ float s_st[4][16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  const int n = blockIdx.y, tile = blockIdx.x;
  const int by0 = (tile / a.tiles_x) * kTileH, bx0 = (tile % a.tiles_x) * TW;
  const int R = (kTileH - 1) * a.si + (a.taps.max_dy - a.taps.min_dy) + 1;
  const int Cc = (TW - 1) * a.si + (a.taps.max_dx - a.taps.min_dx) + 1;
  const int npx = R * Cc;
  float4 *s_lo = smem, *s_hi = smem + npx;
  const float *f_lo = reinterpret_cast<const float *>(s_lo), *f_hi = reinterpret_cast<const float *>(s_hi);
  float *s_w = reinterpret_cast<float *>(smem + 2 * npx);
  const bool affine = a.src2 != nullptr;
  if (affine) {
    if (tid < 24) {
      const float *t = tid < 8 ? a.coefA : (tid < 16 ? a.coefB : a.coefC);
      s_coef[tid] = t[n * 8 + (tid & 7)];
    }
    __syncthreads();
  }
  float acc[NLIVE][MT][NT][4];
#pragma unroll
  for (int p = 0; p < NLIVE; ++p)
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[p][m][j][i] = 0.f;
  float st_s[2] = {0.f, 0.f}, st_q[2] = {0.f, 0.f};

  const int gy0 = by0 * a.si + a.taps.min_dy, gx0 = bx0 * a.si + a.taps.min_dx;
  const float *srcn = a.src + (int64_t)n * a.src_h * a.src_w * a.src_ld;
  const float *src2n = affine ? a.src2 + (int64_t)n * a.src_h * a.src_w * a.src2_ld : nullptr;
  float *dstn = a.dst + (int64_t)n * a.dst_h * a.dst_w * a.dst_ld;

  // m-tile m of this warp: tile row 2 * warp + m / PIX, columns (m % PIX) * 16 .. + 15
  auto epilogue = [&](int m, int ph, float (&r)[NT][4]) {
    const int by = by0 + 2 * warp + m / PIX;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int bx = bx0 + (m % PIX) * 16 + g + half * 8;
      const int oy = by * a.so + (ph >> 1), ox = bx * a.so + (ph & 1);
      if (by < a.base_h && bx < a.base_w && oy < a.dst_h && ox < a.dst_w) {
        float *o = dstn + ((int64_t)oy * a.dst_w + ox) * a.dst_ld + 2 * tq;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          float v0 = r[j][2 * half], v1 = r[j][2 * half + 1];
          if (a.accumulate) v0 += o[j * 8], v1 += o[j * 8 + 1];
          o[j * 8] = v0, o[j * 8 + 1] = v1;
          if (NC == 8) st_s[0] += v0, st_s[1] += v1, st_q[0] += v0 * v0, st_q[1] += v1 * v1;
        }
      }
    }
  };

  for (int ch = 0; ch < NCH; ++ch) {
    if (ch > 0) __syncthreads();
    for (int i = tid; i < npx; i += kTileThreads) {
      const int r = i / Cc, c = i - r * Cc;
      const int gy = gy0 + r, gx = gx0 + c;
      float4 lo = f4zero(), hi = f4zero();
      if (gy >= 0 && gy < a.src_h && gx >= 0 && gx < a.src_w) {
        const int64_t pix = (int64_t)gy * a.src_w + gx;
        const float *p = srcn + pix * a.src_ld + ch * 8;
        lo = ld4(p), hi = ld4(p + 4);
        if (affine) {
          const float *q = src2n + pix * a.src2_ld;
          const float4 ylo = ld4(q), yhi = ld4(q + 4);
          lo.x = s_coef[0] * lo.x + s_coef[8] * ylo.x + s_coef[16];
          lo.y = s_coef[1] * lo.y + s_coef[9] * ylo.y + s_coef[17];
          lo.z = s_coef[2] * lo.z + s_coef[10] * ylo.z + s_coef[18];
          lo.w = s_coef[3] * lo.w + s_coef[11] * ylo.w + s_coef[19];
          hi.x = s_coef[4] * hi.x + s_coef[12] * yhi.x + s_coef[20];
          hi.y = s_coef[5] * hi.y + s_coef[13] * yhi.y + s_coef[21];
          hi.z = s_coef[6] * hi.z + s_coef[14] * yhi.z + s_coef[22];
          hi.w = s_coef[7] * hi.w + s_coef[15] * yhi.w + s_coef[23];
        }
      }
      s_lo[i] = make_float4(senas_tf32(lo.x), senas_tf32(lo.y), senas_tf32(lo.z), senas_tf32(lo.w));
      s_hi[i] = make_float4(senas_tf32(hi.x), senas_tf32(hi.y), senas_tf32(hi.z), senas_tf32(hi.w));
    }
    for (int i = tid; i < a.taps.n * 8 * NC; i += kTileThreads) {
      const int nn = i % NC, kk = (i / NC) & 7, t = i / (NC * 8);
      s_w[(t * 8 + kk) * NCP + nn] =
          senas_tf32(__ldg(a.w + (int64_t)a.taps.widx[t] * a.ws_t + (int64_t)(ch * 8 + kk) * a.ws_k + (int64_t)nn * a.ws_n));
    }
    __syncthreads();
#pragma unroll
    for (int ph = 0; ph < NPH; ++ph) {
      const int pl = kPhaseOuter ? 0 : ph;
      if (kPhaseOuter && NPH > 1) {
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
          for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[0][m][j][i] = 0.f;
      }
      for (int t = a.taps.pstart[ph]; t < a.taps.pstart[ph + 1]; ++t) {
        float bf[NT][2];
        const float *wt = s_w + t * 8 * NCP;
#pragma unroll
        for (int j = 0; j < NT; ++j) bf[j][0] = wt[tq * NCP + j * 8 + g], bf[j][1] = wt[(tq + 4) * NCP + j * 8 + g];
        const int tap_off = (a.taps.dy[t] - a.taps.min_dy) * Cc + (a.taps.dx[t] - a.taps.min_dx);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const int row = 2 * warp + m / PIX, col = (m % PIX) * 16;
          const int i0 = tap_off + row * a.si * Cc + (col + g) * a.si, i1 = i0 + 8 * a.si;
          const float af[4] = {f_lo[i0 * 4 + tq], f_lo[i1 * 4 + tq], f_hi[i0 * 4 + tq], f_hi[i1 * 4 + tq]};
#pragma unroll
          for (int j = 0; j < NT; ++j) senas_mma_tf32(acc[pl][m][j], af, bf[j]);
        }
      }
      if (kPhaseOuter) {
#pragma unroll
        for (int m = 0; m < MT; ++m) epilogue(m, ph, acc[0][m]);
      }
    }
  }
  if (!kPhaseOuter) {
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int ph = 0; ph < NPH; ++ph) epilogue(m, ph, acc[ph][m]);
  }
  if (a.partials != nullptr) {  // uniform.  lane (g, tq) holds channels 2tq, 2tq + 1: sum over g, then over the warps
#pragma unroll
    for (int m = 4; m < 32; m <<= 1) {
      st_s[0] += __shfl_xor_sync(0xffffffffu, st_s[0], m), st_s[1] += __shfl_xor_sync(0xffffffffu, st_s[1], m);
      st_q[0] += __shfl_xor_sync(0xffffffffu, st_q[0], m), st_q[1] += __shfl_xor_sync(0xffffffffu, st_q[1], m);
    }
    if (lane < 4) {
      s_st[warp][2 * lane] = st_s[0], s_st[warp][2 * lane + 1] = st_s[1];
      s_st[warp][8 + 2 * lane] = st_q[0], s_st[warp][8 + 2 * lane + 1] = st_q[1];
    }
    __syncthreads();
    if (tid < 16)
      a.partials[((int64_t)n * gridDim.x + blockIdx.x) * 16 + tid] =
          (s_st[0][tid] + s_st[1][tid]) + (s_st[2][tid] + s_st[3][tid]);
  }
}

// ------------------------------------------------------------------------------------------------
// convolution weight gradient: dW[t][ci][co] = sum_b x[b*si + d_t][ci] * dy[b*so + phase_t][co]
// thread = (ci, tap group); x read straight from global/L1 (32 consecutive ci = one 128 B line),
// dy = A*gm + B*y + C staged per tile in shared memory.  Persistent blocks loop over (sample, tile)
// and write one partial per block; wgrad_reduce_kernel sums them in fixed order.
// ------------------------------------------------------------------------------------------------
struct WgradArgs {
  const float *x;
  int64_t x_ld;
  int32_t x_h, x_w;
  const float *gm, *y;
  int64_t y_ld;
  int32_t o_h, o_w;
  const float *coefA, *coefB, *coefC;
  int32_t base_h, base_w, si, so, tiles_x, tiles_y, batch;
  float *partials;  // [gridDim.x][T][KC][8]
  TapTable taps;
};

template <int KC, int TPT>
__global__ void conv_wgrad_kernel(WgradArgs a) {
  SENAS_DYN_SMEM(float4, smem);  // dy tile: 2 planes of (kTileH*so)*(kTileW*so) float4
  const int tid = threadIdx.x, ci = tid % KC, tg = tid / KC;
  const int oth = kTileH * a.so, otw = kTileW * a.so, onpx = oth * otw;
  float4 *s_lo = smem, *s_hi = smem + onpx;
  float acc[TPT][8];
#pragma unroll
  for (int j = 0; j < TPT; ++j)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;
  int tdy[TPT], tdx[TPT], tph[TPT];
#pragma unroll
  for (int j = 0; j < TPT; ++j) tdy[j] = a.taps.dy[tg * TPT + j], tdx[j] = a.taps.dx[tg * TPT + j], tph[j] = a.taps.phase[tg * TPT + j];
  const int tiles = a.tiles_x * a.tiles_y, total = tiles * a.batch;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int n = item / tiles, tile = item - n * tiles;
    const int by0 = (tile / a.tiles_x) * kTileH, bx0 = (tile % a.tiles_x) * kTileW;
    const float *gmn = a.gm + (int64_t)n * a.o_h * a.o_w * 8;
    const float *yn = a.y + (int64_t)n * a.o_h * a.o_w * a.y_ld;
    const float *xn = a.x + (int64_t)n * a.x_h * a.x_w * a.x_ld;
    __syncthreads();
    for (int i = tid; i < onpx; i += blockDim.x) {
      const int r = i / otw, c = i - r * otw;
      const int oy = by0 * a.so + r, ox = bx0 * a.so + c;
      float4 lo = f4zero(), hi = f4zero();
      if (oy < a.o_h && ox < a.o_w) {
        const int64_t pix = (int64_t)oy * a.o_w + ox;
        const float4 glo = ld4(gmn + pix * 8), ghi = ld4(gmn + pix * 8 + 4);
        const float4 ylo = ld4(yn + pix * a.y_ld), yhi = ld4(yn + pix * a.y_ld + 4);
        const float *A = a.coefA + n * 8, *B = a.coefB + n * 8, *C = a.coefC + n * 8;
        lo.x = A[0] * glo.x + B[0] * ylo.x + C[0];
        lo.y = A[1] * glo.y + B[1] * ylo.y + C[1];
        lo.z = A[2] * glo.z + B[2] * ylo.z + C[2];
        lo.w = A[3] * glo.w + B[3] * ylo.w + C[3];
        hi.x = A[4] * ghi.x + B[4] * yhi.x + C[4];
        hi.y = A[5] * ghi.y + B[5] * yhi.y + C[5];
        hi.z = A[6] * ghi.z + B[6] * yhi.z + C[6];
        hi.w = A[7] * ghi.w + B[7] * yhi.w + C[7];
      }
      s_lo[i] = lo, s_hi[i] = hi;
    }
    __syncthreads();
#pragma unroll 2
    for (int b = 0; b < kTileH * kTileW; ++b) {
      const int ty = b / kTileW, tx = b - ty * kTileW;
      const int by = by0 + ty, bx = bx0 + tx;
      const bool bok = by < a.base_h && bx < a.base_w;
      float xv_[TPT];
#pragma unroll
      for (int j = 0; j < TPT; ++j) {  // all loads of this pixel first
        const int iy = by * a.si + tdy[j], ix = bx * a.si + tdx[j];
        const bool ok = bok && iy >= 0 && iy < a.x_h && ix >= 0 && ix < a.x_w;
        xv_[j] = ok ? __ldg(xn + ((int64_t)iy * a.x_w + ix) * a.x_ld + ci) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < TPT; ++j) {
        const float xv = xv_[j];
        const int idx = (ty * a.so + (tph[j] >> 1)) * otw + tx * a.so + (tph[j] & 1);
        const float4 lo = s_lo[idx], hi = s_hi[idx];
        acc[j][0] = fmaf(xv, lo.x, acc[j][0]);
        acc[j][1] = fmaf(xv, lo.y, acc[j][1]);
        acc[j][2] = fmaf(xv, lo.z, acc[j][2]);
        acc[j][3] = fmaf(xv, lo.w, acc[j][3]);
        acc[j][4] = fmaf(xv, hi.x, acc[j][4]);
        acc[j][5] = fmaf(xv, hi.y, acc[j][5]);
        acc[j][6] = fmaf(xv, hi.z, acc[j][6]);
        acc[j][7] = fmaf(xv, hi.w, acc[j][7]);
      }
    }
  }
  float *out = a.partials + (int64_t)blockIdx.x * a.taps.n * KC * 8;
#pragma unroll
  for (int j = 0; j < TPT; ++j) {
    const int t = tg * TPT + j;
    float *o = out + ((int64_t)t * KC + ci) * 8;
    st4(o, make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]));
    st4(o + 4, make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]));
  }
}

// Sum `rows` rows of length V (fixed order): block = 256 threads = 8 warps x 32 consecutive columns; warp w takes
// rows w, w+8, ...; the 8 partial sums are combined in warp order.  Returns the column sum in warp 0's lanes.
SENAS_DEVFN float block_rows_sum(const float *src, int rows, int V, int col) {
  __shared__ float s_rs[8][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float r = 0.f;
  if (col < V) {
    float r1 = 0.f, r2 = 0.f, r3 = 0.f;
    int b = warp;
    for (; b + 24 < rows; b += 32) {  // four loads in flight (these reductions are pure load latency)
      r += src[(int64_t)b * V + col], r1 += src[(int64_t)(b + 8) * V + col];
      r2 += src[(int64_t)(b + 16) * V + col], r3 += src[(int64_t)(b + 24) * V + col];
    }
    for (; b < rows; b += 8) r += src[(int64_t)b * V + col];
    r = (r + r1) + (r2 + r3);
  }
  __syncthreads();
  s_rs[warp][lane] = r;
  __syncthreads();
  float t = 0.f;
  if (warp == 0)
    for (int w = 0; w < 8; ++w) t += s_rs[w][lane];
  return t;
}

// dst[group][V] = sum over rows of src[group][rows][V];  grid = (ceil(V/8), groups), block = 1024 = 128 row lanes x 8
// columns (a 32-byte sector per row and block): V is a few hundred at most, so 32 columns per block left the reduction of
// a few thousand partial rows to ~10-25 blocks.  These launches sit on every candidate's backward chain (statistics ->
// reduce -> coefficients -> dz), i.e. their latency is step time: 128 row lanes with 4 loads in flight each instead of 32
// (ncu, round 2: 15 us at 2048 rows x 320 columns, long-scoreboard stalls 50 per issue -- pure load latency).
// Fixed order: lane r sums rows r, r+128, ... in four interleaved chains; lanes combined 16 at a time, then 8.
constexpr int kRowsReduceCols = 8, kRowsReduceLanes = 128, kRowsReduceThreads = kRowsReduceCols * kRowsReduceLanes;
// column sums of one block: valid in threads < kRowsReduceCols (column blockIdx.x * 8 + threadIdx.x)
SENAS_DEVFN float rows_reduce_block(const float *src, int rows, int V) {
  __shared__ float s_rr[kRowsReduceLanes][kRowsReduceCols + 1];
  __shared__ float s_r2[8][kRowsReduceCols + 1];
  constexpr int L = kRowsReduceLanes;
  const int cl = threadIdx.x & 7, rl = threadIdx.x >> 3, col = blockIdx.x * kRowsReduceCols + cl;
  const float *p = src + (int64_t)blockIdx.y * rows * V + col;
  float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
  if (col < V) {
    int b = rl;
    for (; b + 3 * L < rows; b += 4 * L) {
      r0 += p[(int64_t)b * V], r1 += p[(int64_t)(b + L) * V], r2 += p[(int64_t)(b + 2 * L) * V], r3 += p[(int64_t)(b + 3 * L) * V];
    }
    for (; b < rows; b += L) r0 += p[(int64_t)b * V];
  }
  s_rr[rl][cl] = (r0 + r1) + (r2 + r3);
  __syncthreads();
  if (threadIdx.x < 64) {  // (column, group of 16 lanes)
    const int c = threadIdx.x & 7, g = threadIdx.x >> 3;
    float t = 0.f;
    for (int r = 0; r < 16; ++r) t += s_rr[g * 16 + r][c];
    s_r2[g][c] = t;
  }
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < kRowsReduceCols)
    for (int g = 0; g < 8; ++g) t += s_r2[g][threadIdx.x];
  return t;
}
__global__ void __launch_bounds__(kRowsReduceThreads) rows_reduce_kernel(const float *src, float *dst, int rows, int V) {
  const float t = rows_reduce_block(src, rows, V);
  const int col = blockIdx.x * kRowsReduceCols + threadIdx.x;
  if (threadIdx.x < kRowsReduceCols && col < V) dst[(int64_t)blockIdx.y * V + col] = t;
}
// The pass-1 fold of the dep-sep pointwise backward WITH its finalize (every output of pw_bfin_kernel depends on its own
// column sum only): sums [10C] = (sum du | sum du zhat | dW_pw) -> d beta1, d gamma1, dW_pw and the BN1-backward coefficients
// [3][C] in one launch; 540 launches per search step fewer on the candidates' backward chains (round 2).  grid = (ceil(10C / 8), 1)
__global__ void __launch_bounds__(kRowsReduceThreads) pw_reduce_fin_kernel(const float *src, int rows, int C, float M, const float *g1,
                                                                           const float *istd1, float *coef /*[3][C]*/, float *g_gamma1,
                                                                           float *g_beta1, float *g_wpw) {
  const float s = rows_reduce_block(src, rows, 10 * C);
  const int i = blockIdx.x * kRowsReduceCols + threadIdx.x;
  if (threadIdx.x >= kRowsReduceCols || i >= 10 * C) return;
  if (i < C) {
    g_beta1[i] = s;
    coef[C + i] = s / M;
    coef[i] = g1[i] * istd1[i];
  } else if (i < 2 * C) {
    g_gamma1[i - C] = s;
    coef[2 * C + i - C] = s / M;
  } else {
    g_wpw[i - 2 * C] = s;
  }
}

// dst[widx_t*ws_t + ci*ws_k + co*ws_n] = sum_blk partials[blk][t][ci][co];  grid = ceil(T*KC*8/32), block = 256
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float *partials, int nblk, int T, int KC, float *dst,
                                                           int ws_t, int ws_k, int ws_n, TapTable taps) {
  const int total = T * KC * 8, i = blockIdx.x * 32 + (threadIdx.x & 31);
  const float s = block_rows_sum(partials, nblk, total, i);
  if (threadIdx.x < 32 && i < total) {
    const int co = i & 7, ci = (i >> 3) % KC, t = i / (8 * KC);
    dst[(int64_t)taps.widx[t] * ws_t + (int64_t)ci * ws_k + (int64_t)co * ws_n] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// depthwise convolution (dep_sep_conv_*, first half): z = dw(x), statistics of z.
// thread = (base pixel, channel quad); x through L1 (a pixel's C floats are one or a quarter line).
// ------------------------------------------------------------------------------------------------
struct DwArgs {
  const float *x;
  int64_t x_ld;
  int32_t x_h, x_w;
  float *z;  // [B][o_h][o_w][C]
  int32_t o_h, o_w, base_h, base_w, si, so;
  const float *w;   // [C][T] (PyTorch [C,1,k,k])
  float *partials;  // [B][gridDim.x][2C]
  TapTable taps;
};

template <int C>
__global__ void __launch_bounds__(128) dw_fwd_kernel(DwArgs a) {
  constexpr int Q = C / 4, PPB = 128 / Q;
  __shared__ float s_w[25 * C];
  __shared__ float s_red[128][8];
  const int tid = threadIdx.x, n = blockIdx.y, pl = tid / Q, q = tid % Q;
  const int T = a.taps.n;
  for (int i = tid; i < T * C; i += 128) {
    const int c = i % C, t = i / C;
    s_w[i] = __ldg(a.w + c * T + a.taps.widx[t]);
  }
  __syncthreads();
  const int b = blockIdx.x * PPB + pl;
  const bool ok = b < a.base_h * a.base_w;
  const int by = ok ? b / a.base_w : 0, bx = ok ? b - by * a.base_w : 0;
  const float *xn = a.x + (int64_t)n * a.x_h * a.x_w * a.x_ld + q * 4;
  float *zn = a.z + (int64_t)n * a.o_h * a.o_w * C + q * 4;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
  if (ok) {
    for (int ph = 0; ph < a.taps.nphase; ++ph) {
      float4 r = f4zero();
      for (int t = a.taps.pstart[ph]; t < a.taps.pstart[ph + 1]; ++t) {
        const int iy = by * a.si + a.taps.dy[t], ix = bx * a.si + a.taps.dx[t];
        if (iy < 0 || iy >= a.x_h || ix < 0 || ix >= a.x_w) continue;
        const float4 xv = ld4(xn + ((int64_t)iy * a.x_w + ix) * a.x_ld);
        const float4 wv = ld4(s_w + t * C + q * 4);
        r.x = fmaf(xv.x, wv.x, r.x), r.y = fmaf(xv.y, wv.y, r.y), r.z = fmaf(xv.z, wv.z, r.z), r.w = fmaf(xv.w, wv.w, r.w);
      }
      const int oy = by * a.so + (ph >> 1), ox = bx * a.so + (ph & 1);
      if (oy < a.o_h && ox < a.o_w) {
        st4(zn + ((int64_t)oy * a.o_w + ox) * C, r);
        s[0] += r.x, s[1] += r.y, s[2] += r.z, s[3] += r.w;
        sq[0] += r.x * r.x, sq[1] += r.y * r.y, sq[2] += r.z * r.z, sq[3] += r.w * r.w;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) s_red[tid][j] = s[j], s_red[tid][4 + j] = sq[j];
  __syncthreads();
  if (tid < 2 * C) {
    const int c = tid % C, which = tid / C;
    float r = 0.f;
    for (int p = 0; p < PPB; ++p) r += s_red[p * Q + c / 4][which * 4 + (c & 3)];
    a.partials[((int64_t)n * gridDim.x + blockIdx.x) * 2 * C + tid] = r;
  }
}

// pointwise half: y = W_pw . relu(BN1(z)), statistics of y.  thread = pixel.
struct PwArgs {
  const float *z;
  int64_t z_ld;  // floats between pixels of z (C for the dep-sep intermediate; the state's pixel stride for adapters)
  int32_t hw;
  float *y;  // [B][hw][8]
  const float *mean1, *istd1, *g1, *b1;
  const float *wpw;  // [8][C]
  float *partials;   // [B][gridDim.x][16]
  int32_t z_bf, pad_;  // z is the bf16-stored dep-sep intermediate
};

// LP = C/4 lanes share a pixel: lane l loads channels 4l..4l+3 as one float4 (a warp reads 512 contiguous bytes), applies
// BN1 + ReLU, multiplies its 4 x 8 slice of W_pw, and a transpose-reduce over the LP lanes (7 shuffles for C = 32)
// leaves output channel l (C = 32) / channels 4l..4l+3 (C = 8) in lane l, so the y store is coalesced too.  (The
// thread-per-pixel version read its own 128-byte line with 8 loads: 8 x the L1 wavefronts for the same bytes.)
constexpr int kPwPx = 512;  // pixels per block
// PLAIN: no BN1 / ReLU in front (the 1x1 of an identity / up_sample AdapterBlock); partials may then be null (no statistics).
// (256, 3): ncu showed the BN1 + ReLU variant at 88 registers = 2 blocks per SM, 63 us and long-scoreboard bound, against
// 44 us for the PLAIN variant at 80 registers = 3 blocks per SM on the same traffic
template <int C, bool PLAIN>
__global__ void __launch_bounds__(256, 3) pw_fwd_kernel(PwArgs a) {
  constexpr int LP = C / 4, PPW = 32 / LP, NOUT = 8 / LP, UN = 4;
  __shared__ float s_red[8][16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, n = blockIdx.y;
  const int l = lane % LP, sub = lane / LP;
  float sc[4], sh[4], w[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = l * 4 + j;
    sc[j] = 1.f, sh[j] = 0.f;
    if (!PLAIN) sc[j] = a.g1[c] * a.istd1[c], sh[j] = a.b1[c] - a.mean1[c] * sc[j];
#pragma unroll
    for (int co = 0; co < 8; ++co) w[j][co] = __ldg(a.wpw + co * C + c);
  }
  float ssum[NOUT], ssq[NOUT];
#pragma unroll
  for (int j = 0; j < NOUT; ++j) ssum[j] = ssq[j] = 0.f;
  const int p_begin = blockIdx.x * kPwPx, p_end = min(p_begin + kPwPx, a.hw);
  const int per_warp = kPwPx / 8;
  const int w_begin = p_begin + warp * per_warp, w_end = min(w_begin + per_warp, p_end);
  const int64_t z0 = (int64_t)n * a.hw * a.z_ld + l * 4;
  const int z_bf = PLAIN ? 0 : a.z_bf;
  float *yn = a.y + (int64_t)n * a.hw * 8;
  for (int p0 = w_begin; p0 < w_end; p0 += PPW * UN) {
    float4 zv[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int p = p0 + u * PPW + sub;
      zv[u] = p < w_end ? ldx4(a.z, z0 + (int64_t)p * a.z_ld, z_bf) : f4zero();
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int p = p0 + u * PPW + sub;
      float r[4] = {zv[u].x, zv[u].y, zv[u].z, zv[u].w};
      if (!PLAIN) {
#pragma unroll
        for (int j = 0; j < 4; ++j) r[j] = fmaxf(fmaf(r[j], sc[j], sh[j]), 0.f);
      }
      float v[8];
#pragma unroll
      for (int co = 0; co < 8; ++co) v[co] = r[0] * w[0][co];
#pragma unroll
      for (int j = 1; j < 4; ++j)
#pragma unroll
        for (int co = 0; co < 8; ++co) v[co] = fmaf(r[j], w[j][co], v[co]);
      // transpose-reduce over the LP lanes of the pixel (fixed order)
#pragma unroll
      for (int m = LP / 2, nk = 4; m >= 1; m >>= 1, nk >>= 1) {
        const bool hi = (lane & m) != 0;
#pragma unroll
        for (int j = 0; j < nk; ++j) {
          const float send = hi ? v[j] : v[j + nk], keep = hi ? v[j + nk] : v[j];
          v[j] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
      }
      if (p < w_end) {
        float *yp = yn + (int64_t)p * 8 + l * NOUT;
        if (NOUT == 4) st4(yp, make_float4(v[0], v[1], v[2], v[3]));
        else yp[0] = v[0];
#pragma unroll
        for (int j = 0; j < NOUT; ++j) ssum[j] += v[j], ssq[j] += v[j] * v[j];
      }
    }
  }
  // statistics: combine the pixel groups of a warp (lanes with equal l), then the 8 warps, in fixed order
#pragma unroll
  for (int m = LP; m < 32; m <<= 1)
#pragma unroll
    for (int j = 0; j < NOUT; ++j)
      ssum[j] += __shfl_xor_sync(0xffffffffu, ssum[j], m), ssq[j] += __shfl_xor_sync(0xffffffffu, ssq[j], m);
  if (lane < LP) {
#pragma unroll
    for (int j = 0; j < NOUT; ++j) s_red[warp][l * NOUT + j] = ssum[j], s_red[warp][8 + l * NOUT + j] = ssq[j];
  }
  __syncthreads();
  if (tid < 16 && a.partials != nullptr) {
    float r = 0.f;
    for (int wv = 0; wv < 8; ++wv) r += s_red[wv][tid];
    a.partials[((int64_t)n * gridDim.x + blockIdx.x) * 16 + tid] = r;
  }
}

// ------------------------------------------------------------------------------------------------
// adapters (AdapterBlock, operations.py:167-183): identity / avg_pool / up_sample -> 1x1 -> y
// ------------------------------------------------------------------------------------------------
// bilinear x2, align_corners=False: src = max((o + 0.5)/2 - 0.5, 0)
SENAS_DEVFN void up_src(int o, int n_in, int &i0, int &i1, float &lam) {
  float s = (o + 0.5f) * 0.5f - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  lam = s - (float)i0;
}
// weight with which input index i contributes to bilinear output o (0 if it does not)
SENAS_DEVFN float up_weight(int o, int i, int n_in) {
  if (o < 0 || o >= 2 * n_in) return 0.f;
  int i0, i1;
  float lam;
  up_src(o, n_in, i0, i1, lam);
  float w = 0.f;
  if (i0 == i) w += 1.f - lam;
  if (i1 == i) w += lam;
  return w;
}
// AvgPool2d(3, stride 2, pad 1, count_include_pad=False): number of valid taps along one axis
SENAS_DEVFN int pool_cnt(int o, int n_in) {
  const int lo = 2 * o - 1 < 0 ? 0 : 2 * o - 1, hi = 2 * o + 1 > n_in - 1 ? n_in - 1 : 2 * o + 1;
  return hi - lo + 1;
}

enum { AD_IDENTITY = 1, AD_POOL = 2, AD_UP = 3 };
constexpr int kPxTilesPerBlock = 8;  // 128-pixel tiles per block in the pixel-wise weight-gradient kernels

struct AdapterArgs {
  const float *x;
  int64_t x_ld;
  int32_t x_h, x_w;
  float *y;  // [B][o_h][o_w][8]; null for stats-only (identity 8->8)
  int32_t o_h, o_w;
  const float *w;  // [8][C] 1x1 weight, null when C == 8 identity
  float *partials;
};

// 32-channel (or 8-channel) "adapter input" a[c] at output pixel (oy, ox)
template <int C, int KIND>
SENAS_DEVFN void adapter_input(const float *xn, int64_t x_ld, int x_h, int x_w, int oy, int ox, float *av) {
  if (KIND == AD_IDENTITY) {
    const float *p = xn + ((int64_t)oy * x_w + ox) * x_ld;
#pragma unroll
    for (int c = 0; c < C; c += 4) {
      const float4 v = ld4(p + c);
      av[c] = v.x, av[c + 1] = v.y, av[c + 2] = v.z, av[c + 3] = v.w;
    }
  } else if (KIND == AD_POOL) {
#pragma unroll
    for (int c = 0; c < C; ++c) av[c] = 0.f;
    int cnt = 0;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int iy = 2 * oy + dy, ix = 2 * ox + dx;
        if (iy < 0 || iy >= x_h || ix < 0 || ix >= x_w) continue;
        ++cnt;
        const float *p = xn + ((int64_t)iy * x_w + ix) * x_ld;
#pragma unroll
        for (int c = 0; c < C; c += 4) {
          const float4 v = ld4(p + c);
          av[c] += v.x, av[c + 1] += v.y, av[c + 2] += v.z, av[c + 3] += v.w;
        }
      }
    const float inv = 1.f / (float)cnt;
#pragma unroll
    for (int c = 0; c < C; ++c) av[c] *= inv;
  } else {  // AD_UP
    int y0, y1, x0, x1;
    float ly, lx;
    up_src(oy, x_h, y0, y1, ly);
    up_src(ox, x_w, x0, x1, lx);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const float *p00 = xn + ((int64_t)y0 * x_w + x0) * x_ld, *p01 = xn + ((int64_t)y0 * x_w + x1) * x_ld;
    const float *p10 = xn + ((int64_t)y1 * x_w + x0) * x_ld, *p11 = xn + ((int64_t)y1 * x_w + x1) * x_ld;
#pragma unroll
    for (int c = 0; c < C; c += 4) {
      const float4 a = ld4(p00 + c), b = ld4(p01 + c), d = ld4(p10 + c), e = ld4(p11 + c);
      av[c] = w00 * a.x + w01 * b.x + w10 * d.x + w11 * e.x;
      av[c + 1] = w00 * a.y + w01 * b.y + w10 * d.y + w11 * e.y;
      av[c + 2] = w00 * a.z + w01 * b.z + w10 * d.z + w11 * e.z;
      av[c + 3] = w00 * a.w + w01 * b.w + w10 * d.w + w11 * e.w;
    }
  }
}

template <int C, int KIND>
__global__ void __launch_bounds__(128) adapter_fwd_kernel(AdapterArgs a) {
  __shared__ float s_w[C * 8];  // [ci][co]
  const int tid = threadIdx.x, n = blockIdx.y;
  if (a.w != nullptr)
    for (int i = tid; i < C * 8; i += 128) s_w[i] = __ldg(a.w + (i & 7) * C + (i >> 3));
  __syncthreads();
  const int p = blockIdx.x * 128 + tid, hw = a.o_h * a.o_w;
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = 0.f;
  if (p < hw) {
    const int oy = p / a.o_w, ox = p - oy * a.o_w;
    float av[C];
    adapter_input<C, KIND>(a.x + (int64_t)n * a.x_h * a.x_w * a.x_ld, a.x_ld, a.x_h, a.x_w, oy, ox, av);
    float acc[8];
    if (a.w != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float4 w0 = ld4(s_w + c * 8), w1 = ld4(s_w + c * 8 + 4);
        acc[0] = fmaf(av[c], w0.x, acc[0]), acc[1] = fmaf(av[c], w0.y, acc[1]);
        acc[2] = fmaf(av[c], w0.z, acc[2]), acc[3] = fmaf(av[c], w0.w, acc[3]);
        acc[4] = fmaf(av[c], w1.x, acc[4]), acc[5] = fmaf(av[c], w1.y, acc[5]);
        acc[6] = fmaf(av[c], w1.z, acc[6]), acc[7] = fmaf(av[c], w1.w, acc[7]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = av[j];  // C == 8 identity: y aliases x
    }
    if (a.y != nullptr) {
      float *yp = a.y + ((int64_t)n * hw + p) * 8;
      st4(yp, make_float4(acc[0], acc[1], acc[2], acc[3]));
      st4(yp + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = acc[j], v[8 + j] = acc[j] * acc[j];
  }
  block_sum_store<16>(v, a.partials + ((int64_t)n * gridDim.x + blockIdx.x) * 16);
}

// adapter backward, data gradient.  thread = input pixel: s[8] = gather of dy (kind specific), then
// dx[ci] (+)= sum_co s[co] * W[co][ci]  (or dx (+)= s when there is no 1x1).
struct AdapterBwdArgs {
  const float *x;  // forward input (dW only)
  int64_t x_ld;
  int32_t x_h, x_w;
  const float *gm, *y;  // dy = A*gm + B*y + C on the output grid
  int64_t y_ld;
  int32_t o_h, o_w;
  const float *coefA, *coefB, *coefC;
  const float *w;  // [8][C] or null
  float *dx;
  int64_t dx_ld;
  int32_t accumulate;
  float *partials;  // dW partials [B*gridDim.x][8*C]
};

SENAS_DEVFN void load_dy8(const AdapterBwdArgs &a, int n, int oy, int ox, float *d) {
  const int64_t pix = ((int64_t)n * a.o_h + oy) * a.o_w + ox;
  const float4 glo = ld4(a.gm + pix * 8), ghi = ld4(a.gm + pix * 8 + 4);
  const float4 ylo = ld4(a.y + pix * a.y_ld), yhi = ld4(a.y + pix * a.y_ld + 4);
  const float *A = a.coefA + n * 8, *B = a.coefB + n * 8, *C = a.coefC + n * 8;
  d[0] = A[0] * glo.x + B[0] * ylo.x + C[0], d[1] = A[1] * glo.y + B[1] * ylo.y + C[1];
  d[2] = A[2] * glo.z + B[2] * ylo.z + C[2], d[3] = A[3] * glo.w + B[3] * ylo.w + C[3];
  d[4] = A[4] * ghi.x + B[4] * yhi.x + C[4], d[5] = A[5] * ghi.y + B[5] * yhi.y + C[5];
  d[6] = A[6] * ghi.z + B[6] * yhi.z + C[6], d[7] = A[7] * ghi.w + B[7] * yhi.w + C[7];
}

// s[8] = d(1x1 output) gathered onto input pixel (iy, ix)
template <int KIND>
SENAS_DEVFN void adapter_gather_dy(const AdapterBwdArgs &a, int n, int iy, int ix, float *s) {
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  if (KIND == AD_IDENTITY) {
    load_dy8(a, n, iy, ix, s);
  } else if (KIND == AD_POOL) {
    // output windows containing i: o in [ceil((i-1)/2), floor((i+1)/2)]
    for (int oy = iy >> 1; oy <= (iy + 1) >> 1; ++oy) {
      if (oy >= a.o_h) continue;
      for (int ox = ix >> 1; ox <= (ix + 1) >> 1; ++ox) {
        if (ox >= a.o_w) continue;
        float d[8];
        load_dy8(a, n, oy, ox, d);
        const float inv = 1.f / (float)(pool_cnt(oy, a.x_h) * pool_cnt(ox, a.x_w));
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += d[j] * inv;
      }
    }
  } else {  // AD_UP: outputs 2i-1 .. 2i+2 per axis
    for (int oy = 2 * iy - 1; oy <= 2 * iy + 2; ++oy) {
      const float wy = up_weight(oy, iy, a.x_h);
      if (wy == 0.f) continue;
      for (int ox = 2 * ix - 1; ox <= 2 * ix + 2; ++ox) {
        const float wx = up_weight(ox, ix, a.x_w);
        if (wx == 0.f) continue;
        float d[8];
        load_dy8(a, n, oy, ox, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += d[j] * (wy * wx);
      }
    }
  }
}

template <int C, int KIND>
__global__ void __launch_bounds__(128) adapter_dx_kernel(AdapterBwdArgs a) {
  __shared__ float s_w[C * 8];  // [co][ci]
  const int tid = threadIdx.x, n = blockIdx.y;
  if (a.w != nullptr)
    for (int i = tid; i < C * 8; i += 128) s_w[i] = __ldg(a.w + i);
  __syncthreads();
  const int p = blockIdx.x * 128 + tid;
  if (p >= a.x_h * a.x_w) return;
  const int iy = p / a.x_w, ix = p - iy * a.x_w;
  float s[8];
  adapter_gather_dy<KIND>(a, n, iy, ix, s);
  float *o = a.dx + ((int64_t)n * a.x_h * a.x_w + p) * a.dx_ld;
#pragma unroll
  for (int c = 0; c < C; c += 4) {
    float4 r;
    if (a.w != nullptr) {
      r = f4zero();
#pragma unroll
      for (int co = 0; co < 8; ++co) {
        const float4 w4 = ld4(s_w + co * C + c);
        r.x = fmaf(s[co], w4.x, r.x), r.y = fmaf(s[co], w4.y, r.y), r.z = fmaf(s[co], w4.z, r.z), r.w = fmaf(s[co], w4.w, r.w);
      }
    } else {
      r = make_float4(s[c], s[c + 1], s[c + 2], s[c + 3]);
    }
    if (a.accumulate) {
      const float4 u = ld4(o + c);
      r.x += u.x, r.y += u.y, r.z += u.z, r.w += u.w;
    }
    st4(o + c, r);
  }
}

// adapter / pointwise weight gradient: dW[co][ci] = sum_p s[p][co] * a[p][ci] over a 128-pixel tile,
// both staged in shared memory; one partial per block.
//   AD_IDENTITY: p over x grid, a = x,        s = dy
//   AD_POOL    : p over output grid, a = pooled(x), s = dy
//   AD_UP      : p over x (low-res) grid, a = x, s = bilinear^T(dy)
template <int C, int KIND>
__global__ void __launch_bounds__(128) adapter_dw_kernel(AdapterBwdArgs a) {
  __shared__ float s_a[128][C + 1];
  __shared__ float s_s[128][9];
  const int tid = threadIdx.x, n = blockIdx.y;
  const int gh = (KIND == AD_POOL) ? a.o_h : a.x_h, gw = (KIND == AD_POOL) ? a.o_w : a.x_w;
  float racc[(8 * C + 127) / 128];
#pragma unroll
  for (int i = 0; i < (8 * C + 127) / 128; ++i) racc[i] = 0.f;
  for (int tile = 0; tile < kPxTilesPerBlock; ++tile) {
  const int p = (blockIdx.x * kPxTilesPerBlock + tile) * 128 + tid;
  float av[C], s[8];
#pragma unroll
  for (int c = 0; c < C; ++c) av[c] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  __syncthreads();
  if (p < gh * gw) {
    const int py = p / gw, px = p - py * gw;
    const float *xn = a.x + (int64_t)n * a.x_h * a.x_w * a.x_ld;
    if (KIND == AD_POOL) {
      adapter_input<C, AD_POOL>(xn, a.x_ld, a.x_h, a.x_w, py, px, av);
      load_dy8(a, n, py, px, s);
    } else {
      adapter_input<C, AD_IDENTITY>(xn, a.x_ld, a.x_h, a.x_w, py, px, av);
      adapter_gather_dy<KIND>(a, n, py, px, s);
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) s_a[tid][c] = av[c];
#pragma unroll
  for (int j = 0; j < 8; ++j) s_s[tid][j] = s[j];
  __syncthreads();
  for (int o = tid, i = 0; o < 8 * C; o += 128, ++i) {
    const int co = o / C, ci = o - co * C;
    float r = racc[i];
    for (int q = 0; q < 128; ++q) r = fmaf(s_s[q][co], s_a[q][ci], r);
    racc[i] = r;
  }
  }
  float *out = a.partials + ((int64_t)n * gridDim.x + blockIdx.x) * 8 * C;
  for (int o = tid, i = 0; o < 8 * C; o += 128, ++i) out[o] = racc[i];
}

// ------------------------------------------------------------------------------------------------
// BatchNorm statistics finalize (one block per BN instance)
// ------------------------------------------------------------------------------------------------
struct Bases {
  float *p[8];
  int64_t ld[8];
};
enum { SP_SAVED = 0, SP_SCRATCH = 1, SP_IN0 = 2, SP_IN1 = 3, SP_OUT = 4, SP_GOUT = 5, SP_NULL = 7 };
struct Ref {
  int32_t space, pad;
  int64_t off, ld;  // ld == 0: take the space's ld
};
SENAS_DEVFN float *ref_ptr(const Ref &r, const Bases &b) { return r.space == SP_NULL ? nullptr : b.p[r.space] + r.off; }
SENAS_DEVFN int64_t ref_ld(const Ref &r, const Bases &b) { return r.ld ? r.ld : b.ld[r.space]; }

struct BnDesc {
  int32_t C, nblk, zero_input, pad;  // zero_input: the 'none' candidate (no partials, y == 0)
  float count_per_sample;
  float pad2;
  int64_t part_off;                   // scratch: [B][nblk][2C]
  int64_t psum_off;                   // scratch: [B][2C] per-sample sums (stage 1 -> stage 2)
  int64_t mean_off, istd_off;         // saved
  int64_t ysum_off;                   // saved [B][C], -1 when not needed
  float *gamma, *beta, *rmean, *rvar;
  int64_t *nbt;
};

// stage 1: per-sample sums.  grid = (instances, batch), block = 256: thread = (row group g, value j)
__global__ void __launch_bounds__(256) bn_reduce_kernel(const BnDesc *descs, Bases bases) {
  const BnDesc d = descs[blockIdx.x];
  if (d.zero_input) return;
  __shared__ float s_g[256];
  float *scratch = bases.p[SP_SCRATCH];
  const int tid = threadIdx.x, n = blockIdx.y, V = 2 * d.C, G = 256 / V, j = tid % V, g = tid / V;
  const float *part = scratch + d.part_off + (int64_t)n * d.nblk * V;
  float r = 0.f;
  for (int b = g; b < d.nblk; b += G) r += part[(int64_t)b * V + j];
  s_g[tid] = r;
  __syncthreads();
  if (tid < V) {
    float t = 0.f;
    for (int gg = 0; gg < G; ++gg) t += s_g[gg * V + tid];
    scratch[d.psum_off + (int64_t)n * V + tid] = t;
  }
}

// stage 2: one block per BN instance
__global__ void __launch_bounds__(64) bn_finalize_kernel(const BnDesc *descs, Bases bases, int batch, int training) {
  const BnDesc d = descs[blockIdx.x];
  float *saved = bases.p[SP_SAVED], *scratch = bases.p[SP_SCRATCH];
  __shared__ double s_tot[64];
  const int tid = threadIdx.x, V = 2 * d.C;
  if (tid < V) {
    double t = 0.0;
    if (!d.zero_input) {
      for (int n = 0; n < batch; ++n) {
        const float v = scratch[d.psum_off + (int64_t)n * V + tid];
        t += (double)v;
        if (d.ysum_off >= 0 && tid < d.C) saved[d.ysum_off + (int64_t)n * d.C + tid] = v;
      }
    }
    s_tot[tid] = t;
  }
  __syncthreads();
  if (tid < d.C) {
    const double cnt = (double)d.count_per_sample * batch;
    const double mean = s_tot[tid] / cnt;
    double var = s_tot[d.C + tid] / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    float m_use, v_use;
    if (training) {
      m_use = (float)mean, v_use = (float)var;
      const double unb = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
      d.rmean[tid] = (1.f - SENAS_MOMENTUM) * d.rmean[tid] + SENAS_MOMENTUM * (float)mean;
      d.rvar[tid] = (1.f - SENAS_MOMENTUM) * d.rvar[tid] + SENAS_MOMENTUM * (float)unb;
      if (tid == 0) *d.nbt += 1;
    } else {
      m_use = d.rmean[tid], v_use = d.rvar[tid];
    }
    saved[d.mean_off + tid] = m_use;
    saved[d.istd_off + tid] = 1.0f / sqrtf(v_use + SENAS_EPS);
  }
}

// ------------------------------------------------------------------------------------------------
// node coefficients + SE gate, node combine
// ------------------------------------------------------------------------------------------------
constexpr int kMaxTerms = 24;
struct TermDesc {
  int32_t kind, edge, cand, has_y;
  Ref y;                                 // pre-BN candidate output (8 channels)
  int64_t mean_off, istd_off, ysum_off;  // saved
  int64_t se_off;                        // saved: s[B][8], q[B][8], h[B]
  int64_t scale_off;                     // scratch [B][8]   (forward)
  int64_t coef_off;                      // scratch [3][B][8] (backward A, B, C)
  float *gamma, *beta, *w1, *w2;
  int64_t g_gamma, g_beta, g_w1, g_w2;  // offsets in the flat gradient buffer
  float hw, pad;
};
struct NodeDesc {
  int32_t nterms, node, nedges, pad;
  int32_t edges[4];
  int64_t bias_off;     // scratch [B][8]
  int64_t gm_off;       // scratch [B][HW][8]
  int64_t dnode_off;    // scratch [B][HW][8], -1 when the node feeds no edge
  int64_t bpart_off;    // scratch [B][nblk][(1+nterms)*8]
  int64_t bsum_off;     // scratch [B][(1+nterms)*8]: rows_reduce of bpart
  int32_t nblk, hw;
  TermDesc t[kMaxTerms];
};

// one block.  Phase A: thread = (term, channel) gathers the per-term constants (all loads independent; the first version
// walked the terms serially with ~6 dependent global loads each, ~20 us on the critical path of every node);
// phase B: thread = (sample, channel) sums the bias over the terms and evaluates the SE gates.
__global__ void node_coef_kernel(const NodeDesc *nodes, int node, Bases bases, const float *alpha, const float *beta,
                                 int batch) {
  const NodeDesc &nd = nodes[node];
  float *saved = bases.p[SP_SAVED], *scratch = bases.p[SP_SCRATCH];
  __shared__ float s_scale[kMaxTerms * 8], s_bias[kMaxTerms * 8];
  for (int i = threadIdx.x; i < nd.nterms * 8; i += blockDim.x) {
    const TermDesc &t = nd.t[i >> 3];
    const int c = i & 7;
    const float kappa = alpha[t.edge * 6 + t.cand] * (beta ? beta[t.edge] : 1.f);
    const float g = t.gamma[c], b = t.beta[c], mean = saved[t.mean_off + c], istd = saved[t.istd_off + c];
    s_scale[i] = kappa * g * istd, s_bias[i] = kappa * (b - g * mean * istd);
  }
  __syncthreads();
  const int total = batch * 8, rounded = (total + 31) & ~31;
  for (int i = threadIdx.x; i < rounded; i += blockDim.x) {
    const bool ok = i < total;
    const int n = ok ? i >> 3 : 0, c = i & 7;
    float bias = 0.f;
    for (int ti = 0; ti < nd.nterms; ++ti) {
      const TermDesc &t = nd.t[ti];
      float s = 1.f;
      if (t.kind == 5) {  // SE_CONV: gate from the per-sample mean of BN(y)
        const float g = t.gamma[c], b = t.beta[c], mean = saved[t.mean_off + c], istd = saved[t.istd_off + c];
        const float q = g * (saved[t.ysum_off + n * 8 + c] / t.hw - mean) * istd + b;
        float h = q * t.w1[c];
        h += __shfl_xor_sync(0xffffffffu, h, 1);
        h += __shfl_xor_sync(0xffffffffu, h, 2);
        h += __shfl_xor_sync(0xffffffffu, h, 4);
        const float a = fmaxf(h, 0.f);
        s = 1.f / (1.f + expf(-t.w2[c] * a));
        if (ok) {
          saved[t.se_off + n * 8 + c] = s;
          saved[t.se_off + batch * 8 + n * 8 + c] = q;
          if (c == 0) saved[t.se_off + batch * 16 + n] = h;
        }
      }
      if (ok && t.has_y) scratch[t.scale_off + n * 8 + c] = s_scale[ti * 8 + c] * s;
      bias += s_bias[ti * 8 + c] * s;
    }
    if (ok) scratch[nd.bias_off + n * 8 + c] = bias;
  }
}

__global__ void __launch_bounds__(128) node_combine_kernel(const NodeDesc *nodes, int node, Bases bases, int relu) {
  const NodeDesc &nd = nodes[node];
  const float *scratch = bases.p[SP_SCRATCH];
  // per-(sample, term, channel) scales staged once per block: inside the term loop they were 8 dependent global loads in
  // front of every pair of y loads (20 us per launch at 16 x 16 pixels: pure latency)
  __shared__ float s_sc[kMaxTerms * 8 + 8];
  const int n = blockIdx.y, p = blockIdx.x * 128 + threadIdx.x;
  for (int i = threadIdx.x; i < nd.nterms * 8; i += 128) {
    const TermDesc &t = nd.t[i >> 3];
    s_sc[8 + i] = t.has_y ? scratch[t.scale_off + n * 8 + (i & 7)] : 0.f;
  }
  if (threadIdx.x < 8) s_sc[threadIdx.x] = scratch[nd.bias_off + n * 8 + threadIdx.x];
  __syncthreads();
  if (p >= nd.hw) return;
  const int64_t pix = (int64_t)n * nd.hw + p;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = s_sc[j];
  for (int ti = 0; ti < nd.nterms; ++ti) {
    const TermDesc &t = nd.t[ti];
    if (!t.has_y) continue;
    const float *y = ref_ptr(t.y, bases) + pix * ref_ld(t.y, bases);
    const float *sc = s_sc + 8 + ti * 8;
    const float4 lo = ld4(y), hi = ld4(y + 4);
    acc[0] = fmaf(sc[0], lo.x, acc[0]), acc[1] = fmaf(sc[1], lo.y, acc[1]);
    acc[2] = fmaf(sc[2], lo.z, acc[2]), acc[3] = fmaf(sc[3], lo.w, acc[3]);
    acc[4] = fmaf(sc[4], hi.x, acc[4]), acc[5] = fmaf(sc[5], hi.y, acc[5]);
    acc[6] = fmaf(sc[6], hi.z, acc[6]), acc[7] = fmaf(sc[7], hi.w, acc[7]);
  }
  if (relu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.f);
  }
  float *o = bases.p[SP_OUT] + pix * bases.ld[SP_OUT] + nd.node * 8;
  st4(o, make_float4(acc[0], acc[1], acc[2], acc[3]));
  st4(o + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
}

// ------------------------------------------------------------------------------------------------
// backward: node statistics sweep
//   gm = (g_out[node] + d_node) * relu'  ->  scratch;  per block: S1 = sum gm, S2_t = sum gm * yhat_t
// ------------------------------------------------------------------------------------------------
// thread = pixel, one block-wide reduction per term.  (A channel-lane rewrite -- lane = (pixel of a 4-pixel step, channel),
// gm of 64-128 pixels in registers, terms streamed one after the other, one combine per block -- measured 6 ms per step
// SLOWER (157 -> 148 img/s): 4x the load instructions for the same bytes and a dependent chain per term.)
__global__ void __launch_bounds__(128) node_bstats_kernel(const NodeDesc *nodes, int node, Bases bases, int relu) {
  const NodeDesc &nd = nodes[node];
  float *scratch = bases.p[SP_SCRATCH];
  const float *saved = bases.p[SP_SAVED];
  const int n = blockIdx.y, p = blockIdx.x * 128 + threadIdx.x;
  const bool ok = p < nd.hw;
  const int64_t pix = (int64_t)n * nd.hw + (ok ? p : 0);
  float g[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = 0.f;
  if (ok) {
    const float *gp = bases.p[SP_GOUT] + pix * bases.ld[SP_GOUT] + nd.node * 8;
    float4 lo = ld4(gp), hi = ld4(gp + 4);
    if (nd.dnode_off >= 0) {
      const float4 a = ld4(scratch + nd.dnode_off + pix * 8), b = ld4(scratch + nd.dnode_off + pix * 8 + 4);
      lo.x += a.x, lo.y += a.y, lo.z += a.z, lo.w += a.w, hi.x += b.x, hi.y += b.y, hi.z += b.z, hi.w += b.w;
    }
    if (relu) {
      const float *op = bases.p[SP_OUT] + pix * bases.ld[SP_OUT] + nd.node * 8;
      const float4 a = ld4(op), b = ld4(op + 4);
      lo.x = a.x > 0.f ? lo.x : 0.f, lo.y = a.y > 0.f ? lo.y : 0.f, lo.z = a.z > 0.f ? lo.z : 0.f;
      lo.w = a.w > 0.f ? lo.w : 0.f, hi.x = b.x > 0.f ? hi.x : 0.f, hi.y = b.y > 0.f ? hi.y : 0.f;
      hi.z = b.z > 0.f ? hi.z : 0.f, hi.w = b.w > 0.f ? hi.w : 0.f;
    }
    st4(scratch + nd.gm_off + pix * 8, lo);
    st4(scratch + nd.gm_off + pix * 8 + 4, hi);
    g[0] = lo.x, g[1] = lo.y, g[2] = lo.z, g[3] = lo.w, g[4] = hi.x, g[5] = hi.y, g[6] = hi.z, g[7] = hi.w;
  }
  float *part = scratch + nd.bpart_off + ((int64_t)n * nd.nblk + blockIdx.x) * (1 + nd.nterms) * 8;
  block_sum_store<8>(g, part);
  for (int ti = 0; ti < nd.nterms; ++ti) {
    const TermDesc &t = nd.t[ti];
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (t.has_y && ok) {
      const float *y = ref_ptr(t.y, bases) + pix * ref_ld(t.y, bases);
      const float4 lo = ld4(y), hi = ld4(y + 4);
      const float yv[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = g[j] * (yv[j] - saved[t.mean_off + j]) * saved[t.istd_off + j];
    }
    if (t.has_y) block_sum_store<8>(v, part + (1 + ti) * 8);  // uniform branch
  }
}

// backward finalize for one node: reductions, parameter / alpha / beta gradients, dy coefficient tables.
// grid = (incoming edges of the node, 6 candidates): the terms are independent except for d beta_edge; this kernel sits on
// the critical path of every node's backward, 3 per cell (~50 us as a single block walking up to 24 terms, 13 us with one
// block per edge)
__global__ void __launch_bounds__(128) node_bfin_kernel(const NodeDesc *nodes, int node, Bases bases, const float *alpha,
                                                        const float *beta, float *g_alpha, float *g_beta, float *g_params,
                                                        int batch, int training) {
  const NodeDesc &nd = nodes[node];
  float *scratch = bases.p[SP_SCRATCH];
  const float *saved = bases.p[SP_SAVED];
  __shared__ float s_S1[128 * 8], s_S2[128 * 8];  // [n][c], batch <= 128
  __shared__ float s_u[128 * 8], s_dh[128];
  __shared__ float s_T[8], s_dg[8], s_db[8];
  __shared__ float s_gbeta;
  const int my_edge = nd.edges[blockIdx.x];
  const int tid = threadIdx.x, total = batch * 8, V = (1 + nd.nterms) * 8;
  const float *bsum = scratch + nd.bsum_off;
  const float M = (float)batch * (float)nd.hw;
  if (tid == 0) s_gbeta = 0.f;
  for (int i = tid; i < total; i += 128) s_S1[i] = bsum[(int64_t)(i >> 3) * V + (i & 7)];
  __syncthreads();
  for (int ti = 0; ti < nd.nterms; ++ti) {
    const TermDesc &t = nd.t[ti];
    if (t.edge != my_edge) continue;  // block-uniform
    // grid.y = candidate: block (edge, k) finalizes term k of the edge; block (edge, 0) also evaluates the T of the
    // other terms of its edge (a short loop) because d beta_edge sums them in candidate order
    const bool full = t.cand == (int)blockIdx.y;
    if (!full && blockIdx.y != 0) continue;
    const float w = alpha[t.edge * 6 + t.cand], be = beta ? beta[t.edge] : 1.f, kappa = w * be;
    const bool se = t.kind == 5;
    for (int i = tid; i < total; i += 128)
      s_S2[i] = t.has_y ? bsum[(int64_t)(i >> 3) * V + (1 + ti) * 8 + (i & 7)] : 0.f;
    __syncthreads();
    if (se && full) {  // through the gate: u = ds * s(1-s), dh = relu'(h) * sum_c u*W2
      for (int n = tid; n < batch; n += 128) {
        float da = 0.f;
        for (int c = 0; c < 8; ++c) {
          const float s = saved[t.se_off + n * 8 + c];
          const float ds = kappa * (t.gamma[c] * s_S2[n * 8 + c] + t.beta[c] * s_S1[n * 8 + c]);
          const float u = ds * s * (1.f - s);
          s_u[n * 8 + c] = u;
          da += u * t.w2[c];
        }
        s_dh[n] = saved[t.se_off + batch * 16 + n] > 0.f ? da : 0.f;
      }
      __syncthreads();
    }
    if (tid < 8) {
      const int c = tid;
      const float g = t.gamma[c], b = t.beta[c], mean = saved[t.mean_off + c], istd = saved[t.istd_off + c];
      float dg = 0.f, db = 0.f, T = 0.f, dw1 = 0.f, dw2 = 0.f;
      for (int n = 0; n < batch; ++n) {
        const float S1 = s_S1[n * 8 + c], S2 = s_S2[n * 8 + c];
        float s = 1.f, dq = 0.f, Yh = 0.f;
        if (se) s = saved[t.se_off + n * 8 + c];
        if (se && full) {
          dq = s_dh[n] * t.w1[c];
          Yh = (saved[t.ysum_off + n * 8 + c] - t.hw * mean) * istd;
          const float h = saved[t.se_off + batch * 16 + n];
          dw2 += s_u[n * 8 + c] * fmaxf(h, 0.f);
          dw1 += s_dh[n] * saved[t.se_off + batch * 8 + n * 8 + c];
        }
        dg += kappa * s * S2 + dq / t.hw * Yh;
        db += kappa * s * S1 + dq;
        T += s * (g * S2 + b * S1);
      }
      s_T[c] = T, s_dg[c] = dg, s_db[c] = db;
      if (full && t.g_gamma >= 0) g_params[t.g_gamma + c] = dg;
      if (full && t.g_beta >= 0) g_params[t.g_beta + c] = db;
      if (se && full) {
        if (t.g_w1 >= 0) g_params[t.g_w1 + c] = dw1;
        if (t.g_w2 >= 0) g_params[t.g_w2 + c] = dw2;
      }
    }
    __syncthreads();
    if (tid == 0) {
      float Ts = 0.f;
      for (int c = 0; c < 8; ++c) Ts += s_T[c];
      if (full) g_alpha[t.edge * 6 + t.cand] = be * Ts;
      s_gbeta += w * Ts;
    }
    if (t.has_y && full) {
      float *cf = scratch + t.coef_off;
      for (int i = tid; i < total; i += 128) {
        const int n = i >> 3, c = i & 7;
        const float g = t.gamma[c], mean = saved[t.mean_off + c], istd = saved[t.istd_off + c];
        float s = 1.f, dq = 0.f;
        if (se) s = saved[t.se_off + n * 8 + c], dq = s_dh[n] * t.w1[c];
        const float gi = g * istd;
        const float m1 = training ? s_db[c] / M : 0.f, m2 = training ? s_dg[c] / M : 0.f;
        cf[i] = gi * kappa * s;
        cf[total + i] = -gi * istd * m2;
        cf[2 * total + i] = gi * (dq / t.hw - m1 + mean * istd * m2);
      }
    }
    __syncthreads();
  }
  if (g_beta != nullptr && tid == 0 && blockIdx.y == 0) g_beta[my_edge] = s_gbeta;
}

// ------------------------------------------------------------------------------------------------
// dep-sep backward
// ------------------------------------------------------------------------------------------------
struct PwBwdArgs {
  const float *gm, *y;  // y ld = 8
  float *z;             // [B][hw][C]; dz kernel overwrites it with dz
  int32_t hw, batch;
  const float *coefA, *coefB, *coefC;
  const float *mean1, *istd1, *g1, *b1;
  const float *wpw;       // [8][C]
  float *partials;        // pass 1: [B][gridDim.x][10C] = du sums (C), du*zhat sums (C), dWpw (8C)
  const float *bn1_coef;  // pass 2: [3][C] = a1, dbeta1/M, dgamma1/M
  int32_t z_bf, pad_;     // z / dz stored as bf16 (pw_bwd_q_kernel only; the older per-pixel kernels are fp32)
};

// du (gradient at the BN1 output, after the ReLU mask) and side values for one pixel
template <int C>
SENAS_DEVFN void pw_pixel_backward(const PwBwdArgs &a, const float *s_w /*[co][ci]*/, const float *s_sc, const float *s_sh,
                                   int n, int p, float *dy, float *zhat_or_r, float *du, bool want_r) {
  const int64_t pix = (int64_t)n * a.hw + p;
  const float4 glo = ld4(a.gm + pix * 8), ghi = ld4(a.gm + pix * 8 + 4);
  const float4 ylo = ld4(a.y + pix * 8), yhi = ld4(a.y + pix * 8 + 4);
  const float *A = a.coefA + n * 8, *B = a.coefB + n * 8, *Cc = a.coefC + n * 8;
  dy[0] = A[0] * glo.x + B[0] * ylo.x + Cc[0], dy[1] = A[1] * glo.y + B[1] * ylo.y + Cc[1];
  dy[2] = A[2] * glo.z + B[2] * ylo.z + Cc[2], dy[3] = A[3] * glo.w + B[3] * ylo.w + Cc[3];
  dy[4] = A[4] * ghi.x + B[4] * yhi.x + Cc[4], dy[5] = A[5] * ghi.y + B[5] * yhi.y + Cc[5];
  dy[6] = A[6] * ghi.z + B[6] * yhi.z + Cc[6], dy[7] = A[7] * ghi.w + B[7] * yhi.w + Cc[7];
  const float *zp = a.z + pix * C;
#pragma unroll
  for (int c4 = 0; c4 < C; c4 += 4) {
    const float4 zv = ld4(zp + c4);
    const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c4 + k;
      const float u = fmaf(zz[k], s_sc[c], s_sh[c]);
      float dr = 0.f;
#pragma unroll
      for (int co = 0; co < 8; ++co) dr = fmaf(dy[co], s_w[co * C + c], dr);
      du[c] = u > 0.f ? dr : 0.f;
      zhat_or_r[c] = want_r ? fmaxf(u, 0.f) : (zz[k] - a.mean1[c]) * a.istd1[c];
    }
  }
}

// pass 1: statistics of du and the pointwise weight gradient
template <int C>
__global__ void __launch_bounds__(128) pw_bwd_stats_kernel(PwBwdArgs a) {
  __shared__ float s_w[8 * C], s_sc[C], s_sh[C];
  __shared__ float s_r[128][C + 1];
  __shared__ float s_dy[128][9];
  const int tid = threadIdx.x, n = blockIdx.y;
  for (int i = tid; i < 8 * C; i += 128) s_w[i] = __ldg(a.wpw + i);
  if (tid < C) {
    const float sc = a.g1[tid] * a.istd1[tid];
    s_sc[tid] = sc, s_sh[tid] = a.b1[tid] - a.mean1[tid] * sc;
  }
  __syncthreads();
  float racc[(8 * C + 127) / 128], sacc0 = 0.f, sacc1 = 0.f;
#pragma unroll
  for (int i = 0; i < (8 * C + 127) / 128; ++i) racc[i] = 0.f;
  float *out = a.partials + ((int64_t)n * gridDim.x + blockIdx.x) * 10 * C;
  for (int tile = 0; tile < kPxTilesPerBlock; ++tile) {
  const int p = (blockIdx.x * kPxTilesPerBlock + tile) * 128 + tid;
  float dy[8], r[C], du[C];
#pragma unroll
  for (int j = 0; j < 8; ++j) dy[j] = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) r[c] = du[c] = 0.f;
  __syncthreads();
  if (p < a.hw) pw_pixel_backward<C>(a, s_w, s_sc, s_sh, n, p, dy, r, du, true);
#pragma unroll
  for (int c = 0; c < C; ++c) s_r[tid][c] = r[c];
#pragma unroll
  for (int j = 0; j < 8; ++j) s_dy[tid][j] = dy[j];
  __syncthreads();
  for (int o = tid, i = 0; o < 8 * C; o += 128, ++i) {
    const int co = o / C, ci = o - co * C;
    float acc = racc[i];
    for (int q = 0; q < 128; ++q) acc = fmaf(s_dy[q][co], s_r[q][ci], acc);
    racc[i] = acc;
  }
  __syncthreads();
  // du and du*zhat: zhat = (r>0 ? (u - b1)/g1 ...) is not recoverable from r; recompute from z
#pragma unroll
  for (int c = 0; c < C; ++c) s_r[tid][c] = du[c];
  __syncthreads();
  if (tid < C) {
    float acc = sacc0;
    for (int q = 0; q < 128; ++q) acc += s_r[q][tid];
    sacc0 = acc;
  }
  __syncthreads();
  if (p < a.hw) {
    const float *zp = a.z + ((int64_t)n * a.hw + p) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) s_r[tid][c] = du[c] * (zp[c] - a.mean1[c]) * a.istd1[c];
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) s_r[tid][c] = 0.f;
  }
  __syncthreads();
  if (tid < C) {
    float acc = sacc1;
    for (int q = 0; q < 128; ++q) acc += s_r[q][tid];
    sacc1 = acc;
  }
  }
  for (int o = tid, i = 0; o < 8 * C; o += 128, ++i) out[2 * C + o] = racc[i];
  if (tid < C) out[tid] = sacc0, out[C + tid] = sacc1;
}

// (BN1 backward finalize: fused into the fold, pw_reduce_fin_kernel)

// pass 2: dz = a1 * (du - dbeta1/M - zhat * dgamma1/M), in place over z
template <int C>
__global__ void __launch_bounds__(128) pw_bwd_dz_kernel(PwBwdArgs a, int training) {
  __shared__ float s_w[8 * C], s_sc[C], s_sh[C], s_k[3 * C];
  const int tid = threadIdx.x, n = blockIdx.y;
  for (int i = tid; i < 8 * C; i += 128) s_w[i] = __ldg(a.wpw + i);
  for (int i = tid; i < 3 * C; i += 128) s_k[i] = a.bn1_coef[i];
  if (tid < C) {
    const float sc = a.g1[tid] * a.istd1[tid];
    s_sc[tid] = sc, s_sh[tid] = a.b1[tid] - a.mean1[tid] * sc;
  }
  __syncthreads();
  const int p = blockIdx.x * 128 + tid;
  if (p >= a.hw) return;
  float dy[8], zh[C], du[C];
  pw_pixel_backward<C>(a, s_w, s_sc, s_sh, n, p, dy, zh, du, false);
  float *zp = a.z + ((int64_t)n * a.hw + p) * C;
#pragma unroll
  for (int c = 0; c < C; c += 4) {
    float4 r;
    float *rr = &r.x;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      rr[k] = training ? s_k[c + k] * (du[c + k] - s_k[C + c + k] - zh[c + k] * s_k[2 * C + c + k]) : s_k[c + k] * du[c + k];
    st4(zp + c, r);
  }
}

// depthwise data gradient: dx[i][c] (+)= sum_t dz[b(i,t)*so' ...][c] * w[c][t], with the mirrored
// tap table (same base-grid convention as gather_mac dgrad).  thread = (base pixel, channel quad).
struct DwBwdArgs {
  const float *dz;
  int32_t z_h, z_w;
  float *dx;
  int64_t dx_ld;
  int32_t x_h, x_w;
  int32_t accumulate, base_h, base_w, si, so;
  const float *w;  // [C][T]
  TapTable taps;   // mirrored (dgrad) table
  // wgrad
  const float *x;
  int64_t x_ld;
  float *partials;
  int32_t batch, chunk;
};

template <int C>
__global__ void __launch_bounds__(128) dw_dx_kernel(DwBwdArgs a) {
  constexpr int Q = C / 4, PPB = 128 / Q;
  __shared__ float s_w[25 * C];
  const int tid = threadIdx.x, n = blockIdx.y, pl = tid / Q, q = tid % Q;
  const int T = a.taps.n;
  for (int i = tid; i < T * C; i += 128) {
    const int c = i % C, t = i / C;
    s_w[i] = __ldg(a.w + c * T + a.taps.widx[t]);
  }
  __syncthreads();
  const int b = blockIdx.x * PPB + pl;
  if (b >= a.base_h * a.base_w) return;
  const int by = b / a.base_w, bx = b - by * a.base_w;
  const float *zn = a.dz + (int64_t)n * a.z_h * a.z_w * C + q * 4;
  float *xn = a.dx + (int64_t)n * a.x_h * a.x_w * a.dx_ld + q * 4;
  for (int ph = 0; ph < a.taps.nphase; ++ph) {
    const int oy = by * a.so + (ph >> 1), ox = bx * a.so + (ph & 1);
    if (oy >= a.x_h || ox >= a.x_w) continue;
    float4 r = f4zero();
    for (int t = a.taps.pstart[ph]; t < a.taps.pstart[ph + 1]; ++t) {
      const int iy = by * a.si + a.taps.dy[t], ix = bx * a.si + a.taps.dx[t];
      if (iy < 0 || iy >= a.z_h || ix < 0 || ix >= a.z_w) continue;
      const float4 zv = ld4(zn + ((int64_t)iy * a.z_w + ix) * C);
      const float4 wv = ld4(s_w + t * C + q * 4);
      r.x = fmaf(zv.x, wv.x, r.x), r.y = fmaf(zv.y, wv.y, r.y), r.z = fmaf(zv.z, wv.z, r.z), r.w = fmaf(zv.w, wv.w, r.w);
    }
    float *o = xn + ((int64_t)oy * a.x_w + ox) * a.dx_ld;
    if (a.accumulate) {
      const float4 u = ld4(o);
      r.x += u.x, r.y += u.y, r.z += u.z, r.w += u.w;
    }
    st4(o, r);
  }
}

// depthwise weight gradient: dW[c][t] = sum_b x[b*si + d_t][c] * dz[b*so + phase_t][c]; forward tap table.
// thread = (pixel strip, channel) with all taps in registers; strips are combined through shared memory in
// fixed order; one partial per block.  block = 256 threads, `chunk` base pixels of one sample per block.
template <int C, int T>
__global__ void __launch_bounds__(256) dw_wgrad_kernel(DwBwdArgs a) {
  constexpr int NS = 256 / C;  // strips
  __shared__ float s_red[NS][C * T + 1];
  const int tid = threadIdx.x, c = tid % C, strip = tid / C, n = blockIdx.y;
  const int npix = a.base_h * a.base_w;
  const int b0 = blockIdx.x * a.chunk, b1 = b0 + a.chunk < npix ? b0 + a.chunk : npix;
  const float *xn = a.x + (int64_t)n * a.x_h * a.x_w * a.x_ld + c;
  const float *zn = a.dz + (int64_t)n * a.z_h * a.z_w * C + c;
  float acc[T];
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t] = 0.f;
  for (int b = b0 + strip; b < b1; b += NS) {
    const int by = b / a.base_w, bx = b - by * a.base_w;
    if (a.so == 1) {
      const float zv = zn[((int64_t)by * a.z_w + bx) * C];
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int iy = by * a.si + a.taps.dy[t], ix = bx * a.si + a.taps.dx[t];
        const bool ok = iy >= 0 && iy < a.x_h && ix >= 0 && ix < a.x_w;
        const float xv = ok ? __ldg(xn + ((int64_t)iy * a.x_w + ix) * a.x_ld) : 0.f;
        acc[t] = fmaf(xv, zv, acc[t]);
      }
    } else {
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int iy = by * a.si + a.taps.dy[t], ix = bx * a.si + a.taps.dx[t], ph = a.taps.phase[t];
        const int oy = by * a.so + (ph >> 1), ox = bx * a.so + (ph & 1);
        const bool ok = iy >= 0 && iy < a.x_h && ix >= 0 && ix < a.x_w && oy < a.z_h && ox < a.z_w;
        const float xv = ok ? __ldg(xn + ((int64_t)iy * a.x_w + ix) * a.x_ld) : 0.f;
        const float zv = ok ? zn[((int64_t)oy * a.z_w + ox) * C] : 0.f;
        acc[t] = fmaf(xv, zv, acc[t]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < T; ++t) s_red[strip][c * T + a.taps.widx[t]] = acc[t];
  __syncthreads();
  float *out = a.partials + ((int64_t)n * gridDim.x + blockIdx.x) * C * T;
  for (int o = tid; o < C * T; o += 256) {
    float r = 0.f;
    for (int q = 0; q < NS; ++q) r += s_red[q][o];
    out[o] = r;
  }
}

// depthwise weight gradient, stride-1 output grids (NORM, DOWN): sliding register window along x.
// thread = (channel, strip); a strip walks 64-pixel row segments keeping the K x K input window of its channel in
// registers, so each step loads K*SI new inputs instead of K*K.  Same partial layout as dw_wgrad_kernel.
template <int C, int K, int SI>
__global__ void __launch_bounds__(256) dw_wgrad_sw_kernel(DwBwdArgs a) {
  constexpr int NS = 256 / C, T = K * K, PAD = K / 2, L = 64, WC = K + SI - 1;
  __shared__ float s_red[NS][C * T + 1];
  const int tid = threadIdx.x, c = tid % C, strip = tid / C, n = blockIdx.y;
  const int segs = (a.base_w + L - 1) / L;
  const int row0 = blockIdx.x * a.chunk, rows = min(a.chunk, a.base_h - row0);
  const float *xn = a.x + (int64_t)n * a.x_h * a.x_w * a.x_ld + c;
  const float *zn = a.dz + (int64_t)n * a.z_h * a.z_w * C + c;
  float acc[T];
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t] = 0.f;
  for (int item = strip; item < rows * segs; item += NS) {
    const int by = row0 + item / segs, bx0 = (item % segs) * L, bx1 = min(bx0 + L, a.base_w);
    float win[K][WC];  // win[j][i] = x[by*SI + j - PAD][bx*SI + i - PAD]
    const float *xr[K];
    bool rok[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int iy = by * SI + j - PAD;
      rok[j] = iy >= 0 && iy < a.x_h;
      xr[j] = xn + (int64_t)(rok[j] ? iy : 0) * a.x_w * a.x_ld;
#pragma unroll
      for (int i = SI; i < WC; ++i) {  // columns that survive the first shift
        const int ix = bx0 * SI + (i - SI) - PAD;
        win[j][i] = (rok[j] && ix >= 0 && ix < a.x_w) ? __ldg(xr[j] + (int64_t)ix * a.x_ld) : 0.f;
      }
    }
    for (int bx = bx0; bx < bx1; ++bx) {
#pragma unroll
      for (int j = 0; j < K; ++j) {
#pragma unroll
        for (int i = 0; i + SI < WC; ++i) win[j][i] = win[j][i + SI];
#pragma unroll
        for (int i = WC - SI; i < WC; ++i) {
          const int ix = bx * SI + i - PAD;
          win[j][i] = (rok[j] && ix >= 0 && ix < a.x_w) ? __ldg(xr[j] + (int64_t)ix * a.x_ld) : 0.f;
        }
      }
      const float zv = zn[((int64_t)by * a.z_w + bx) * C];
#pragma unroll
      for (int j = 0; j < K; ++j)
#pragma unroll
        for (int i = 0; i < K; ++i) acc[j * K + i] = fmaf(win[j][i], zv, acc[j * K + i]);
    }
  }
#pragma unroll
  for (int t = 0; t < T; ++t) s_red[strip][c * T + t] = acc[t];
  __syncthreads();
  float *out = a.partials + ((int64_t)n * gridDim.x + blockIdx.x) * C * T;
  for (int o = tid; o < C * T; o += 256) {
    float r = 0.f;
    for (int q = 0; q < NS; ++q) r += s_red[q][o];
    out[o] = r;
  }
}

// dep-sep pointwise backward, channel-centric: lane = channel (C = 32: one pixel per warp step; C = 8: four pixels),
// the 8 dy values of a pixel are computed once per warp and broadcast with shuffles, the column W_pw[:, c] lives in
// registers, and every reduction (sum du, sum du*zhat, dW_pw[:, c]) stays inside the lane -- no shared-memory transposes.
//   PASS 1: statistics + dW_pw partials ([10C] per block, same layout as before);   PASS 2: dz in place over z.
template <int C, int PASS>
__global__ void __launch_bounds__(256) pw_bwd_cc_kernel(PwBwdArgs a, int px_per_block, int training) {
  constexpr int PPW = 32 / C;  // pixels per warp step
  __shared__ float s_part[8][10 * C];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, n = blockIdx.y;
  const int c = lane % C, sub = lane / C;
  const float g1 = a.g1[c], b1 = a.b1[c], mean1 = a.mean1[c], istd1 = a.istd1[c];
  const float sc = g1 * istd1, sh = b1 - mean1 * sc;
  float wc[8];
#pragma unroll
  for (int co = 0; co < 8; ++co) wc[co] = __ldg(a.wpw + co * C + c);
  const int co_l = lane & 7;
  const float cA = a.coefA[n * 8 + co_l], cB = a.coefB[n * 8 + co_l], cC = a.coefC[n * 8 + co_l];
  float k0 = 0.f, k1 = 0.f, k2 = 0.f;
  if (PASS == 2) k0 = a.bn1_coef[c], k1 = a.bn1_coef[C + c], k2 = a.bn1_coef[2 * C + c];
  float s_du = 0.f, s_duz = 0.f, dw[8];
#pragma unroll
  for (int co = 0; co < 8; ++co) dw[co] = 0.f;
  const int p_begin = blockIdx.x * px_per_block, p_end = min(p_begin + px_per_block, a.hw);
  const int per_warp = (px_per_block + 7) / 8;
  const int w_begin = p_begin + warp * per_warp, w_end = min(w_begin + per_warp, p_end);
  constexpr int UN = 4;  // pixel groups in flight per warp step (memory-level parallelism)
  for (int p0 = w_begin; p0 < w_end; p0 += PPW * UN) {
    float zv[UN], dyl[UN];
    bool okv[UN];
    int64_t pixv[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int p = p0 + u * PPW + sub;
      okv[u] = p < w_end;
      pixv[u] = (int64_t)n * a.hw + (okv[u] ? p : w_begin);
      const int pd = p0 + u * PPW + (C == 32 ? 0 : (lane >> 3));
      const int64_t pix_d = (int64_t)n * a.hw + (pd < w_end ? pd : w_begin);
      zv[u] = a.z[pixv[u] * C + c];
      dyl[u] = cA * a.gm[pix_d * 8 + co_l] + cB * a.y[pix_d * 8 + co_l] + cC;
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      float dy[8];
#pragma unroll
      for (int co = 0; co < 8; ++co) dy[co] = __shfl_sync(0xffffffffu, dyl[u], (C == 32 ? 0 : sub * 8) + co);
      const float z = zv[u];
      const bool ok = okv[u];
      const float uu = fmaf(z, sc, sh);
      float dr = 0.f;
#pragma unroll
      for (int co = 0; co < 8; ++co) dr = fmaf(dy[co], wc[co], dr);
      const float du = (uu > 0.f && ok) ? dr : 0.f;
      const float zhat = (z - mean1) * istd1;
      if (PASS == 1) {
        const float r = ok ? fmaxf(uu, 0.f) : 0.f;
        s_du += du, s_duz += du * zhat;
#pragma unroll
        for (int co = 0; co < 8; ++co) dw[co] = fmaf(dy[co], r, dw[co]);
      } else if (ok) {
        a.z[pixv[u] * C + c] = training ? k0 * (du - k1 - zhat * k2) : k0 * du;
      }
    }
  }
  if (PASS == 1) {
    // combine the PPW pixel groups of a warp (C == 8) and the 8 warps in fixed order
    if (C == 8) {
#pragma unroll
      for (int m = 8; m < 32; m <<= 1) {
        s_du += __shfl_xor_sync(0xffffffffu, s_du, m), s_duz += __shfl_xor_sync(0xffffffffu, s_duz, m);
#pragma unroll
        for (int co = 0; co < 8; ++co) dw[co] += __shfl_xor_sync(0xffffffffu, dw[co], m);
      }
    }
    if (lane < C) {
      s_part[warp][c] = s_du, s_part[warp][C + c] = s_duz;
#pragma unroll
      for (int co = 0; co < 8; ++co) s_part[warp][2 * C + co * C + c] = dw[co];
    }
    __syncthreads();
    float *out = a.partials + ((int64_t)n * gridDim.x + blockIdx.x) * 10 * C;
    for (int o = tid; o < 10 * C; o += 256) {
      float r = 0.f;
      for (int w = 0; w < 8; ++w) r += s_part[w][o];
      out[o] = r;
    }
  }
}

// depthwise weight gradient of the transposed (UP) depthwise conv: dW[c][ky][kx] = sum_b x[b + d(ky), b + d(kx)] *
// dz[2b + p(ky), 2b + p(kx)] with p(kk) = (K/2 + kk) & 1, d(kk) = (p + K/2 - kk) / 2.  Per input
// pixel b only a 3x3 window of x and the 2x2 block of dz are needed; the x window slides along the row in registers.
template <int C, int K>
__global__ void __launch_bounds__(256) dw_wgrad_up_kernel(DwBwdArgs a) {
  constexpr int NS = 256 / C, T = K * K, PAD = K / 2, L = 64;
  __shared__ float s_red[NS][C * T + 1];
  const int tid = threadIdx.x, c = tid % C, strip = tid / C, n = blockIdx.y;
  const int segs = (a.base_w + L - 1) / L;
  const int row0 = blockIdx.x * a.chunk, rows = min(a.chunk, a.base_h - row0);
  const float *xn = a.x + (int64_t)n * a.x_h * a.x_w * a.x_ld + c;
  const float *zn = a.dz + (int64_t)n * a.z_h * a.z_w * C + c;
  float acc[T];
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t] = 0.f;
  for (int item = strip; item < rows * segs; item += NS) {
    const int by = row0 + item / segs, bx0 = (item % segs) * L, bx1 = min(bx0 + L, a.base_w);
    float win[3][3];  // win[j][i] = x[by + j - 1][bx + i - 1]
    const float *xr[3];
    bool rok[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int iy = by + j - 1;
      rok[j] = iy >= 0 && iy < a.x_h;
      xr[j] = xn + (int64_t)(rok[j] ? iy : 0) * a.x_w * a.x_ld;
#pragma unroll
      for (int i = 1; i < 3; ++i) {
        const int ix = bx0 + (i - 1) - 1;
        win[j][i] = (rok[j] && ix >= 0 && ix < a.x_w) ? __ldg(xr[j] + (int64_t)ix * a.x_ld) : 0.f;
      }
    }
    for (int bx = bx0; bx < bx1; ++bx) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        win[j][0] = win[j][1], win[j][1] = win[j][2];
        const int ix = bx + 1;
        win[j][2] = (rok[j] && ix < a.x_w) ? __ldg(xr[j] + (int64_t)ix * a.x_ld) : 0.f;
      }
      float zv[2][2];
#pragma unroll
      for (int py = 0; py < 2; ++py)
#pragma unroll
        for (int px = 0; px < 2; ++px) zv[py][px] = zn[((int64_t)(2 * by + py) * a.z_w + 2 * bx + px) * C];
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const int py = (PAD + ky) & 1, dy = (py + PAD - ky) / 2;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const int px = (PAD + kx) & 1, dx = (px + PAD - kx) / 2;
          acc[ky * K + kx] = fmaf(win[dy + 1][dx + 1], zv[py][px], acc[ky * K + kx]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < T; ++t) s_red[strip][c * T + t] = acc[t];
  __syncthreads();
  float *out = a.partials + ((int64_t)n * gridDim.x + blockIdx.x) * C * T;
  for (int o = tid; o < C * T; o += 256) {
    float r = 0.f;
    for (int q = 0; q < NS; ++q) r += s_red[q][o];
    out[o] = r;
  }
}

// ------------------------------------------------------------------------------------------------
// convolution weight gradient, shared-memory tiled (replaces conv_wgrad_kernel on the hot path).
// The ncu capture of the first version (profiles/) showed DRAM traffic ~ algorithmic but only ~21 % of the
// executed instructions being FMAs (per-tap index math + global loads).  Here the x tile (with halo) and the dy tile
// are staged once per tile, every tap address is `row base register + immediate`, and the pixel loop is fully
// unrolled along x:  per pixel and thread  K LDS (x) + 2 LDS.128 (dy) + 8K FMA.
//   thread = (ci, ky, pixel split)   KC = 32: 32 x K threads;   KC = 8: 8 x K x 4 (row-interleaved pixel splits)
// ------------------------------------------------------------------------------------------------
template <int KC, int K, int SI, int SO>
struct WgradTile {
  static constexpr int TH = (SI == 2) ? 4 : 8;
  static constexpr int TW = (SI == 2) ? 8 : 16;
  static constexpr int PS = 32 / KC;  // pixel splits per (ci, ky)
  static constexpr int THREADS = 32 * K;
};

template <int KC, int K, int SI, int SO>
__global__ void __launch_bounds__(32 * K) conv_wgrad2_kernel(WgradArgs a) {
  using TL = WgradTile<KC, K, SI, SO>;
  constexpr int TH = TL::TH, TW = TL::TW, PS = TL::PS, NT = TL::THREADS;
  SENAS_DYN_SMEM(float4, smem);
  const int tid = threadIdx.x, ci = tid % KC, ps = (tid / KC) % PS, ky = tid / 32;
  const int span_y = a.taps.max_dy - a.taps.min_dy, span_x = a.taps.max_dx - a.taps.min_dx;
  const int XR = (TH - 1) * SI + span_y + 1, XC = (TW - 1) * SI + span_x + 1;  // staged x tile (pixels)
  constexpr int OTH = TH * SO, OTW = TW * SO;
  float *s_x = reinterpret_cast<float *>(smem);                // [XR][XC][KC]
  float4 *s_lo = smem + (XR * XC * KC + 3) / 4, *s_hi = s_lo + OTH * OTW;
  // per-thread tap constants: taps of kernel row ky are t = ky*K + j in table order?  The table is phase-sorted, so
  // look the K taps of this kernel row up by their weight index.
  int xoff[K], doff[K], trow[K], tslot[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    int t = 0;
    for (int q = 0; q < a.taps.n; ++q)
      if (a.taps.widx[q] == ky * K + j) t = q;
    tslot[j] = t;
    trow[j] = a.taps.dy[t] - a.taps.min_dy;
    xoff[j] = (a.taps.dx[t] - a.taps.min_dx) * KC + ci;
    doff[j] = (a.taps.phase[t] >> 1) * OTW + (a.taps.phase[t] & 1);
  }
  float acc[K][8];
#pragma unroll
  for (int j = 0; j < K; ++j)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;
  const int tiles = a.tiles_x * a.tiles_y, total = tiles * a.batch;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int n = item / tiles, tile = item - n * tiles;
    const int by0 = (tile / a.tiles_x) * TH, bx0 = (tile % a.tiles_x) * TW;
    const float *gmn = a.gm + (int64_t)n * a.o_h * a.o_w * 8;
    const float *yn = a.y + (int64_t)n * a.o_h * a.o_w * a.y_ld;
    const float *xn = a.x + (int64_t)n * a.x_h * a.x_w * a.x_ld;
    __syncthreads();
    // x tile (zero outside the image)
    const int gy0 = by0 * SI + a.taps.min_dy, gx0 = bx0 * SI + a.taps.min_dx;
    for (int i = tid; i < XR * XC * (KC / 4); i += NT) {
      const int q = i % (KC / 4), pxl = i / (KC / 4);
      const int r = pxl / XC, c = pxl - r * XC;
      const int gy = gy0 + r, gx = gx0 + c;
      float4 v = f4zero();
      if (gy >= 0 && gy < a.x_h && gx >= 0 && gx < a.x_w) v = ld4(xn + ((int64_t)gy * a.x_w + gx) * a.x_ld + q * 4);
      st4(s_x + (int64_t)pxl * KC + q * 4, v);
    }
    // dy tile = A*gm + B*y + C (zero outside)
    for (int i = tid; i < OTH * OTW; i += NT) {
      const int r = i / OTW, c = i - r * OTW;
      const int oy = by0 * SO + r, ox = bx0 * SO + c;
      float4 lo = f4zero(), hi = f4zero();
      if (oy < a.o_h && ox < a.o_w && by0 + r / SO < a.base_h && bx0 + c / SO < a.base_w) {
        const int64_t pix = (int64_t)oy * a.o_w + ox;
        const float4 glo = ld4(gmn + pix * 8), ghi = ld4(gmn + pix * 8 + 4);
        const float4 ylo = ld4(yn + pix * a.y_ld), yhi = ld4(yn + pix * a.y_ld + 4);
        const float *A = a.coefA + n * 8, *B = a.coefB + n * 8, *C = a.coefC + n * 8;
        lo.x = A[0] * glo.x + B[0] * ylo.x + C[0], lo.y = A[1] * glo.y + B[1] * ylo.y + C[1];
        lo.z = A[2] * glo.z + B[2] * ylo.z + C[2], lo.w = A[3] * glo.w + B[3] * ylo.w + C[3];
        hi.x = A[4] * ghi.x + B[4] * yhi.x + C[4], hi.y = A[5] * ghi.y + B[5] * yhi.y + C[5];
        hi.z = A[6] * ghi.z + B[6] * yhi.z + C[6], hi.w = A[7] * ghi.w + B[7] * yhi.w + C[7];
      }
      s_lo[i] = lo, s_hi[i] = hi;
    }
    __syncthreads();
    for (int ty = ps; ty < TH; ty += PS) {
      const float *xrow[K];
#pragma unroll
      for (int j = 0; j < K; ++j) xrow[j] = s_x + (int64_t)((ty * SI + trow[j]) * XC) * KC + xoff[j];
      const float4 *dlo = s_lo + ty * SO * OTW, *dhi = s_hi + ty * SO * OTW;
#pragma unroll
      for (int tx = 0; tx < TW; ++tx) {
        float4 lo1 = f4zero(), hi1 = f4zero();
        if (SO == 1) lo1 = dlo[tx], hi1 = dhi[tx];  // stride-1 output: dy of this pixel is the same for every tap
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const float xv = xrow[j][tx * SI * KC];
          const float4 lo = SO == 1 ? lo1 : dlo[doff[j] + tx * SO], hi = SO == 1 ? hi1 : dhi[doff[j] + tx * SO];
          acc[j][0] = fmaf(xv, lo.x, acc[j][0]), acc[j][1] = fmaf(xv, lo.y, acc[j][1]);
          acc[j][2] = fmaf(xv, lo.z, acc[j][2]), acc[j][3] = fmaf(xv, lo.w, acc[j][3]);
          acc[j][4] = fmaf(xv, hi.x, acc[j][4]), acc[j][5] = fmaf(xv, hi.y, acc[j][5]);
          acc[j][6] = fmaf(xv, hi.z, acc[j][6]), acc[j][7] = fmaf(xv, hi.w, acc[j][7]);
        }
      }
    }
  }
  if (PS > 1) {  // combine the pixel splits (lanes ci + 8*ps)
#pragma unroll
    for (int m = KC; m < 32; m <<= 1)
#pragma unroll
      for (int j = 0; j < K; ++j)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[j][c] += __shfl_xor_sync(0xffffffffu, acc[j][c], m);
  }
  if (ps == 0) {
    float *out = a.partials + (int64_t)blockIdx.x * a.taps.n * KC * 8;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      float *o = out + ((int64_t)tslot[j] * KC + ci) * 8;
      st4(o, make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]));
      st4(o + 4, make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient of the 8 -> 8 node-edge convolutions (dil_2 / dil_3_conv_5, stride 1) on the tensor cores: warp-level
// mma.sync m16n8k8 TF32 with the PIXELS as GEMM-K (bf16 mode; conv_wgrad2_kernel<8, 5, 1, 1> is the exact fp32 version and
// is bound by its shared-memory reads: ncu, round 2, L1 83 % / FMA 29 %).  Two taps share one MMA:
//   A (16 x 8, row = (tap pair member, ci), k = pixel)   a0 = x[q0][px t][ci g]   a1 = x[q1][px t][ci g]
//                                                        a2 = x[q0][px t + 4][g]  a3 = x[q1][px t + 4][g]
//   B ( 8 x 8, k = pixel, n = co)                        b0 = dy[px t][co g]      b1 = dy[px t + 4][co g]
//   D (16 x 8)   d0, d1 = dW[q0][ci g][co 2t, 2t + 1]    d2, d3 = dW[q1][ci g][co 2t, 2t + 1]
// (g = lane >> 2, t = lane & 3), i.e. 13 MMAs + 54 conflict-free LDS per 8 pixels for all 25 taps, against 25 x 8 x 8
// FMAs + 27 LDS per pixel.  Same tile (8 x 16 output pixels, x tile with halo and dy = A gm + B y + C staged once), same
// persistent grid and the same partial layout [block][tap][ci][co] as conv_wgrad2_kernel, so wgrad_reduce_kernel follows
// unchanged.  Operands are rounded to TF32 when staged (2e-2 gate of the mode).
// ------------------------------------------------------------------------------------------------
constexpr int kWmTH = 8, kWmTW = 16, kWmThreads = 128, kWmTaps = 25;
__global__ void __launch_bounds__(kWmThreads) conv_wgrad_mma8_kernel(WgradArgs a) {
  SENAS_DYN_SMEM(float, smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int span_y = a.taps.max_dy - a.taps.min_dy, span_x = a.taps.max_dx - a.taps.min_dx;
  const int XR = kWmTH + span_y, XC = kWmTW + span_x;
  float *s_x = smem;                             // [XR][XC][8]
  float *s_dy = s_x + XR * XC * 8;               // [TH * TW][8]
  float *s_red = s_dy + kWmTH * kWmTW * 8;       // [4 warps][25 * 64]
  int off[kWmTaps + 1];
#pragma unroll
  for (int q = 0; q < kWmTaps + 1; ++q)
    off[q] = q < a.taps.n ? ((a.taps.dy[q] - a.taps.min_dy) * XC + (a.taps.dx[q] - a.taps.min_dx)) * 8 : 0;
  float acc[(kWmTaps + 1) / 2][4];
#pragma unroll
  for (int p = 0; p < (kWmTaps + 1) / 2; ++p) acc[p][0] = acc[p][1] = acc[p][2] = acc[p][3] = 0.f;
  const int tiles = a.tiles_x * a.tiles_y, total = tiles * a.batch;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int n = item / tiles, tile = item - n * tiles;
    const int by0 = (tile / a.tiles_x) * kWmTH, bx0 = (tile % a.tiles_x) * kWmTW;
    const float *gmn = a.gm + (int64_t)n * a.o_h * a.o_w * 8;
    const float *yn = a.y + (int64_t)n * a.o_h * a.o_w * a.y_ld;
    const float *xn = a.x + (int64_t)n * a.x_h * a.x_w * a.x_ld;
    __syncthreads();
    const int gy0 = by0 + a.taps.min_dy, gx0 = bx0 + a.taps.min_dx;
    for (int i = tid; i < XR * XC * 2; i += kWmThreads) {
      const int q = i & 1, pxl = i >> 1;
      const int r = pxl / XC, c = pxl - r * XC;
      const int gy = gy0 + r, gx = gx0 + c;
      float4 v = f4zero();
      if (gy >= 0 && gy < a.x_h && gx >= 0 && gx < a.x_w) v = ld4(xn + ((int64_t)gy * a.x_w + gx) * a.x_ld + q * 4);
      st4(s_x + (int64_t)pxl * 8 + q * 4, make_float4(senas_tf32(v.x), senas_tf32(v.y), senas_tf32(v.z), senas_tf32(v.w)));
    }
    for (int i = tid; i < kWmTH * kWmTW * 2; i += kWmThreads) {
      const int q = i & 1, pxl = i >> 1;
      const int r = pxl / kWmTW, c = pxl - r * kWmTW;
      const int oy = by0 + r, ox = bx0 + c;
      float4 v = f4zero();
      if (oy < a.o_h && ox < a.o_w) {
        const int64_t pix = (int64_t)oy * a.o_w + ox;
        const float4 gv = ld4(gmn + pix * 8 + q * 4), yv = ld4(yn + pix * a.y_ld + q * 4);
        const float *A = a.coefA + n * 8 + q * 4, *B = a.coefB + n * 8 + q * 4, *C = a.coefC + n * 8 + q * 4;
        v.x = A[0] * gv.x + B[0] * yv.x + C[0], v.y = A[1] * gv.y + B[1] * yv.y + C[1];
        v.z = A[2] * gv.z + B[2] * yv.z + C[2], v.w = A[3] * gv.w + B[3] * yv.w + C[3];
      }
      st4(s_dy + (int64_t)pxl * 8 + q * 4, make_float4(senas_tf32(v.x), senas_tf32(v.y), senas_tf32(v.z), senas_tf32(v.w)));
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int row = 2 * warp + rr;
#pragma unroll
      for (int ks = 0; ks < kWmTW / 8; ++ks) {
        const float *dyp = s_dy + (row * kWmTW + ks * 8) * 8 + g;
        const float bf[2] = {dyp[t * 8], dyp[(t + 4) * 8]};
        const float *xb = s_x + (row * XC + ks * 8) * 8 + g;
#pragma unroll
        for (int p = 0; p < (kWmTaps + 1) / 2; ++p) {
          const int q0 = 2 * p, q1 = 2 * p + 1;
          float af[4];
          af[0] = xb[off[q0] + t * 8], af[2] = xb[off[q0] + (t + 4) * 8];
          if (q1 < kWmTaps) af[1] = xb[off[q1] + t * 8], af[3] = xb[off[q1] + (t + 4) * 8];
          else af[1] = af[3] = 0.f;
          senas_mma_tf32(acc[p], af, bf);
        }
      }
    }
  }
  __syncthreads();  // (tiles done: the staging area is not read any more; s_red is its own region)
#pragma unroll
  for (int p = 0; p < (kWmTaps + 1) / 2; ++p) {
    const int q0 = 2 * p, q1 = 2 * p + 1;
    float *r0 = s_red + warp * (kWmTaps * 64) + (q0 * 8 + g) * 8 + 2 * t;
    r0[0] = acc[p][0], r0[1] = acc[p][1];
    if (q1 < kWmTaps) {
      float *r1 = s_red + warp * (kWmTaps * 64) + (q1 * 8 + g) * 8 + 2 * t;
      r1[0] = acc[p][2], r1[1] = acc[p][3];
    }
  }
  __syncthreads();
  float *out = a.partials + (int64_t)blockIdx.x * kWmTaps * 64;
  for (int o = tid; o < kWmTaps * 64; o += kWmThreads)
    out[o] = (s_red[o] + s_red[kWmTaps * 64 + o]) + (s_red[2 * kWmTaps * 64 + o] + s_red[3 * kWmTaps * 64 + o]);
}

// depthwise convolution with a sliding register window (stride-1 output grids: NORM / DOWN forward, NORM data
// gradient).  thread = (channel, strip): the K x K input window of the thread's channel slides along x in registers,
// K*SI global loads + 1 store per output instead of K*K loads (the first version was L1-bandwidth bound).
//   out[b][c] = sum_{j,i} in[b*SI + (j - PAD), b*SI + (i - PAD)][c] * w[c][wj(j), wi(i)]     FLIP: w index mirrored (dgrad)
// STATS: per-channel sum / sum of squares partials (forward);  ACC: out += (data gradient into a shared dx).
template <int C, int K, int SI, bool FLIP, bool STATS>
__global__ void __launch_bounds__(256) dw_sw_kernel(const float *in, int64_t in_ld, int in_h, int in_w, float *out,
                                                    int64_t out_ld, int base_h, int base_w, const float *w, int accumulate,
                                                    float *partials, int rows_per_block) {
  constexpr int NS = 256 / C, T = K * K, PAD = K / 2, L = 64, WC = K + SI - 1;
  __shared__ float s_red[STATS ? NS : 1][2 * C + 1];
  const int tid = threadIdx.x, c = tid % C, strip = tid / C, n = blockIdx.y;
  const int segs = (base_w + L - 1) / L;
  const int row0 = blockIdx.x * rows_per_block, rows = min(rows_per_block, base_h - row0);
  const float *inn = in + (int64_t)n * in_h * in_w * in_ld + c;
  float *outn = out + (int64_t)n * base_h * base_w * out_ld + c;
  float wk[T];
#pragma unroll
  for (int t = 0; t < T; ++t) wk[t] = __ldg(w + c * T + (FLIP ? T - 1 - t : t));
  float ssum = 0.f, ssq = 0.f;
  for (int item = strip; item < rows * segs; item += NS) {
    const int by = row0 + item / segs, bx0 = (item % segs) * L, bx1 = min(bx0 + L, base_w);
    float win[K][WC];
    const float *xr[K];
    bool rok[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int iy = by * SI + j - PAD;
      rok[j] = iy >= 0 && iy < in_h;
      xr[j] = inn + (int64_t)(rok[j] ? iy : 0) * in_w * in_ld;
#pragma unroll
      for (int i = SI; i < WC; ++i) {
        const int ix = bx0 * SI + (i - SI) - PAD;
        win[j][i] = (rok[j] && ix >= 0 && ix < in_w) ? __ldg(xr[j] + (int64_t)ix * in_ld) : 0.f;
      }
    }
    for (int bx = bx0; bx < bx1; ++bx) {
      float r = 0.f;
#pragma unroll
      for (int j = 0; j < K; ++j) {
#pragma unroll
        for (int i = 0; i + SI < WC; ++i) win[j][i] = win[j][i + SI];
#pragma unroll
        for (int i = WC - SI; i < WC; ++i) {
          const int ix = bx * SI + i - PAD;
          win[j][i] = (rok[j] && ix >= 0 && ix < in_w) ? __ldg(xr[j] + (int64_t)ix * in_ld) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < K; ++i) r = fmaf(win[j][i], wk[j * K + i], r);
      }
      float *o = outn + ((int64_t)by * base_w + bx) * out_ld;
      if (accumulate) r += *o;
      *o = r;
      if (STATS) ssum += r, ssq += r * r;
    }
  }
  if (STATS) {
    s_red[strip][c] = ssum, s_red[strip][C + c] = ssq;
    __syncthreads();
    if (tid < 2 * C) {
      float r = 0.f;
      for (int q = 0; q < NS; ++q) r += s_red[q][tid];
      partials[((int64_t)n * gridDim.x + blockIdx.x) * 2 * C + tid] = r;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Depthwise convolution, stride-1 geometry (NORM edges), several convolutions of ONE input per launch.
//
// The dep-sep candidates of the edges that read the same state (Cell edges 0/2/5 <- in0, 4/7 <- node 0, ...; k = 3
// and k = 5 each) share their input: a block owns one spatial tile and walks the list of convolutions, so x is read
// from HBM once (the re-reads hit L1/L2) and, for the data gradient, the shared dx tile is accumulated by ONE block
// in list order (fixed order => bit-reproducible; one HBM write instead of a read-modify-write per candidate).
//
// thread = (channel quad, pair of adjacent columns); it walks down the rows of its tile keeping the K output rows
// that are "in flight" in registers (slot = output row mod K, compile-time under a K-fold unroll), the K x K weights
// of its 4 channels in registers, and loads each input row once: K + 1 LDG.128 for 2 * K * K * 4 FMA.
//   out[o][bx][c] = sum_{ky,kx} in[o + ky - P][bx + kx - P][c] * w[c][ky*K + kx]        (FLIP: w index mirrored)
// ------------------------------------------------------------------------------------------------
constexpr int kDwMaxItems = 8;
struct DwItem {
  const float *in;    // [B][H][W][in_ld]
  float *out;         // [B][H][W][out_ld]
  const float *w;     // [C][K*K]
  float *partials;    // forward: [B][gridDim.x][2C] sums / sums of squares;  weight gradient: [B][gridDim.x][C*K*K]
  const float *in2;   // weight gradient: dz [B][H][W][C]
  int64_t in_ld, out_ld;
  int32_t k, flip, accumulate, in2_ld;  // in2_ld: pixel stride of in2 (0 = C)
  int32_t in_bf, out_bf, in2_bf, pad_;  // != 0: that tensor is the bf16-stored dep-sep intermediate (z / dz)
};
struct DwMultiArgs {
  DwItem it[kDwMaxItems];
  int32_t n, H, W, tiles_x, tile_rows, pad_;
};

SENAS_DEVFN void fma4(float4 &a, const float4 &x, const float4 &w) {
  a.x = fmaf(x.x, w.x, a.x), a.y = fmaf(x.y, w.y, a.y), a.z = fmaf(x.z, w.z, a.z), a.w = fmaf(x.w, w.w, a.w);
}

// The first version kept the K x K x 4 weights in registers (222-254 registers, 8 warps per SM): ncu showed it
// latency-bound (issue slots 21 % active, FMA pipe 13 %) while filling the register file, so nothing could run beside
// it.  Weights now sit in shared memory as [tap][C] (one conflict-free LDS.128 per tap and row, 8 FMAs each).
// ncu (profiles/): a 128-bit shared-memory load costs ~2.6 LSU wavefronts even when the quarter-warps read the same 128
// bytes, so with 2 columns per thread the 25 weight loads of a 5x5 row step were 65 % of the LSU traffic and the kernel
// sat at 54 % of the L1/LSU pipe with the FMA pipe at 26 %.  NCOL = 4 columns per thread halves the weight traffic per
// FMA; the read-modify-write of a shared dx (data gradient) is issued before the FMAs of the step that completes the row.
template <int C, int K, int NCOL, bool STATS>
SENAS_DEVFN void dw_tile_rows(const DwItem &it, const float *s_w, int n, int H, int W, int by0, int by1, int bx, int q,
                              float *st) {
  constexpr int P = K / 2, NX = K + NCOL - 1;
  float4 acc[K][NCOL];
#pragma unroll
  for (int s = 0; s < K; ++s)
#pragma unroll
    for (int j = 0; j < NCOL; ++j) acc[s][j] = f4zero();
  bool cok[NX];
#pragma unroll
  for (int j = 0; j < NX; ++j) cok[j] = bx - P + j >= 0 && bx - P + j < W;
  const int64_t in0 = (int64_t)n * H * W * it.in_ld + q * 4, out0 = (int64_t)n * H * W * it.out_ld + q * 4;
  const float *wq = s_w + q * 4;
  const int R = by1 - by0, niter = R + K - 1, r_first = by0 - P;
  const int64_t in_ld = it.in_ld;
  const bool rmw = it.accumulate != 0;
  const int in_bf = it.in_bf, out_bf = it.out_bf;
  for (int i0 = 0; i0 < niter; i0 += K) {
#pragma unroll
    for (int u = 0; u < K; ++u) {
      const int i = i0 + u, rr = r_first + i;
      if (i < niter) {
        const int o = by0 + i - (K - 1);  // output row completed by this step (when i >= K - 1)
        float4 old[NCOL];
        if (rmw && i >= K - 1) {
#pragma unroll
          for (int j = 0; j < NCOL; ++j)
            old[j] = bx + j < W ? ldx4(it.out, out0 + ((int64_t)o * W + bx + j) * it.out_ld, out_bf) : f4zero();
        }
        if (rr >= 0 && rr < H) {
          float4 xv[NX];
          const int64_t rowe = in0 + ((int64_t)rr * W + (bx - P)) * in_ld;
#pragma unroll
          for (int j = 0; j < NX; ++j) xv[j] = cok[j] ? ldx4(it.in, rowe + (int64_t)j * in_ld, in_bf) : f4zero();
#pragma unroll
          for (int ky = 0; ky < K; ++ky) {
            if ((unsigned)(i - ky) < (unsigned)R) {  // output row by0 + i - ky is inside the tile
              const int s = (u - ky + K) % K;
#pragma unroll
              for (int kx = 0; kx < K; ++kx) {
                const float4 wv = ld4(wq + (ky * K + kx) * C);
#pragma unroll
                for (int j = 0; j < NCOL; ++j) fma4(acc[s][j], xv[kx + j], wv);
              }
            }
          }
        }
        const int sc = (u + 1) % K;
        if (i >= K - 1) {
#pragma unroll
          for (int j = 0; j < NCOL; ++j) {
            if (bx + j < W) {
              float4 v = acc[sc][j];
              if (rmw) v.x += old[j].x, v.y += old[j].y, v.z += old[j].z, v.w += old[j].w;
              stx4(it.out, out0 + ((int64_t)o * W + bx + j) * it.out_ld, v, out_bf);
              if (STATS) {
                v = roundx4(v, out_bf);
                st[0] += v.x, st[1] += v.y, st[2] += v.z, st[3] += v.w;
                st[4] += v.x * v.x, st[5] += v.y * v.y, st[6] += v.z * v.z, st[7] += v.w * v.w;
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < NCOL; ++j) acc[sc][j] = f4zero();
      }
    }
  }
}

// grid = (tiles_x * tiles_y, B), block = 128:  Q = C/4 channel quads x (128/Q) column groups of kDwCols => 512/Q columns
// per tile, tile_rows rows (chosen by the host so that the grid fills the 148 SMs several times over)
constexpr int kDwCols = 2;
template <int C, bool STATS>
__global__ void __launch_bounds__(128, 4) dw_multi_kernel(DwMultiArgs a) {
  constexpr int Q = C / 4, SLOTS = 128 / Q, NCB = kDwCols * SLOTS;
  __shared__ float s_red[STATS ? 128 : 1][8];
  __shared__ float4 s_w4[25 * C / 4];
  float *s_w = reinterpret_cast<float *>(s_w4);
  const int tid = threadIdx.x, q = tid % Q, slot = tid / Q, n = blockIdx.y;
  const int tx = blockIdx.x % a.tiles_x, ty = blockIdx.x / a.tiles_x;
  const int bx = tx * NCB + kDwCols * slot, by0 = ty * a.tile_rows, by1 = min(by0 + a.tile_rows, a.H);
  const bool active = bx < a.W;
  for (int m = 0; m < a.n; ++m) {
    const DwItem &it = a.it[m];
    const int T = it.k * it.k;
    __syncthreads();  // previous item done with s_w / s_red
    for (int i = tid; i < T * C; i += 128) {
      const int c = i % C, t = i / C;
      s_w[i] = __ldg(it.w + c * T + (it.flip ? T - 1 - t : t));
    }
    __syncthreads();
    float st[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (active) {
      if (it.k == 5) dw_tile_rows<C, 5, kDwCols, STATS>(it, s_w, n, a.H, a.W, by0, by1, bx, q, st);
      else dw_tile_rows<C, 3, kDwCols, STATS>(it, s_w, n, a.H, a.W, by0, by1, bx, q, st);
    }
    if (STATS) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s_red[tid][j] = st[j];
      __syncthreads();
      if (tid < 2 * C) {
        const int c = tid % C, which = tid / C;
        float r = 0.f;
        for (int s = 0; s < SLOTS; ++s) r += s_red[s * Q + c / 4][which * 4 + (c & 3)];
        it.partials[((int64_t)n * gridDim.x + blockIdx.x) * 2 * C + tid] = r;
      }
    }
  }
}

// weight gradients of several depthwise convolutions of one input (stride-1 geometry):
//   dW[c][ky*K + kx] = sum_{o, bx} x[o + ky - P][bx + kx - P][c] * dz[o][bx][c]
// thread = (channel quad, column); same row walk, with the K most recent dz values of its column in registers; the
// K*K*4 accumulators of a thread are combined over the tile's columns with shuffles + shared memory (fixed order).
template <int C, int K>
SENAS_DEVFN void dw_wgrad_rows(const DwItem &it, int n, int H, int W, int by0, int by1, int bx, int q, bool active,
                               float *s_part, int nblk_idx) {
  constexpr int P = K / 2, T = K * K, Q = C / 4;
  float4 acc[T];
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t] = f4zero();
  if (active) {
    float4 dzv[K];
#pragma unroll
    for (int s = 0; s < K; ++s) dzv[s] = f4zero();
    bool cok[K];
#pragma unroll
    for (int j = 0; j < K; ++j) cok[j] = bx - P + j >= 0 && bx - P + j < W;
    const float *inb = it.in + (int64_t)n * H * W * it.in_ld + q * 4;
    const int64_t dz0 = (int64_t)n * H * W * C + q * 4;
    const int in2_bf = it.in2_bf;
    const int R = by1 - by0, niter = R + K - 1, r_first = by0 - P;
    const int64_t in_ld = it.in_ld;
    for (int i0 = 0; i0 < niter; i0 += K) {
#pragma unroll
      for (int u = 0; u < K; ++u) {
        const int i = i0 + u, rr = r_first + i;
        if (i < niter) {
          // dz row entering the window: output row by0 + i (slot u)
          dzv[u] = i < R ? ldx4(it.in2, dz0 + ((int64_t)(by0 + i) * W + bx) * C, in2_bf) : f4zero();
          if (rr >= 0 && rr < H) {
            float4 xv[K];
            const float *rowp = inb + ((int64_t)rr * W + (bx - P)) * in_ld;
#pragma unroll
            for (int j = 0; j < K; ++j) xv[j] = cok[j] ? ld4(rowp + (int64_t)j * in_ld) : f4zero();
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
              if ((unsigned)(i - ky) < (unsigned)R) {
                const int s = (u - ky + K) % K;
#pragma unroll
                for (int kx = 0; kx < K; ++kx) fma4(acc[ky * K + kx], xv[kx], dzv[s]);
              }
            }
          }
        }
      }
    }
  }
  // combine: lanes of a warp that hold the same quad (lane % Q), then the 4 warps through shared memory
#pragma unroll
  for (int t = 0; t < T; ++t) {
#pragma unroll
    for (int m = Q; m < 32; m <<= 1) {
      acc[t].x += __shfl_xor_sync(0xffffffffu, acc[t].x, m), acc[t].y += __shfl_xor_sync(0xffffffffu, acc[t].y, m);
      acc[t].z += __shfl_xor_sync(0xffffffffu, acc[t].z, m), acc[t].w += __shfl_xor_sync(0xffffffffu, acc[t].w, m);
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // s_part free
  if (lane < Q) {
#pragma unroll
    for (int t = 0; t < T; ++t) {
      float *o = s_part + warp * (C * 25) + (lane * 4) * T + t;
      o[0] = acc[t].x, o[T] = acc[t].y, o[2 * T] = acc[t].z, o[3 * T] = acc[t].w;
    }
  }
  __syncthreads();
  float *out = it.partials + (int64_t)nblk_idx * C * T;
  for (int o = threadIdx.x; o < C * T; o += 128)
    out[o] = (s_part[o] + s_part[C * 25 + o]) + (s_part[2 * C * 25 + o] + s_part[3 * C * 25 + o]);
}

// grid = (tiles_x * tiles_y, B), block = 128: Q quads x 128/Q columns, tile_rows rows
template <int C>
__global__ void __launch_bounds__(128) dw_wgrad_multi_kernel(DwMultiArgs a) {
  constexpr int Q = C / 4, SLOTS = 128 / Q;
  __shared__ float s_part[4 * C * 25];
  const int tid = threadIdx.x, q = tid % Q, slot = tid / Q, n = blockIdx.y;
  const int tx = blockIdx.x % a.tiles_x, ty = blockIdx.x / a.tiles_x;
  const int bx = tx * SLOTS + slot, by0 = ty * a.tile_rows, by1 = min(by0 + a.tile_rows, a.H);
  const bool active = bx < a.W;
  const int nblk_idx = n * gridDim.x + blockIdx.x;
  for (int m = 0; m < a.n; ++m) {
    const DwItem &it = a.it[m];
    if (it.k == 5) dw_wgrad_rows<C, 5>(it, n, a.H, a.W, by0, by1, bx, q, active, s_part, nblk_idx);
    else dw_wgrad_rows<C, 3>(it, n, a.H, a.W, by0, by1, bx, q, active, s_part, nblk_idx);
  }
}



// ------------------------------------------------------------------------------------------------
// Depthwise convolutions of NORM edges, "lane = channel" layout (round 2).  With NHWC and C = 32 one warp-wide load is
// the 128 contiguous bytes of one pixel; the K*K weights of a lane's channel sit in registers, a warp walks down a strip
// of WT columns with the K in-flight output rows in registers (slot = row mod K under a K-fold unroll) and the next input
// row is prefetched while the current one is consumed: K*K*WT FMAs per (WT + K - 1) loads, no shared memory, statistics
// and weight-gradient sums lane-local.  Micro-benchmark on B200 (scripts/ubench/dwbench.cu, 16 x 32 x 256 x 256, k3 + k5):
// forward with statistics 0.160 ms against 0.207 ms for dw_multi_kernel (float4 quads, weights in shared memory), weight
// gradient 0.127 ms against 0.262 ms for dw_wgrad_multi_kernel.  C = 8: four strips side by side in a warp.
// Same argument struct, item list, partial layouts and fixed summation order per block as the kernels they replace.
// ------------------------------------------------------------------------------------------------
template <int C>
struct DwLane {
  static constexpr int STRIPS = 32 / C;              // strips per warp
  static constexpr int WT = C == 32 ? 8 : 4;         // columns per thread
  static constexpr int TILE_W = 4 * STRIPS * WT;     // columns per block (4 warps)
};

// forward (STATS: z = dw(x), statistics) and data gradient (flipped weights come in through it.flip; optional += into out)
template <int C, int K, bool STATS>
SENAS_DEVFN void dwl_walk(const DwItem &it, int n, int H, int W, int c0, int r0, int r1, int ch, float &s0, float &s1) {
  constexpr int P = K / 2, WT = DwLane<C>::WT, NX = WT + K - 1, T = K * K;
  float wr[T];
#pragma unroll
  for (int t = 0; t < T; ++t) wr[t] = __ldg(it.w + ch * T + (it.flip ? T - 1 - t : t));
  bool cok[NX];
#pragma unroll
  for (int j = 0; j < NX; ++j) cok[j] = c0 - P + j >= 0 && c0 - P + j < W;
  float acc[K][WT];
#pragma unroll
  for (int s = 0; s < K; ++s)
#pragma unroll
    for (int j = 0; j < WT; ++j) acc[s][j] = 0.f;
  const float *xb = it.in + (int64_t)n * H * W * it.in_ld + ch;
  float *ob = it.out + (int64_t)n * H * W * it.out_ld + ch;
  const int64_t in_ld = it.in_ld, out_ld = it.out_ld;
  const int R = r1 - r0, niter = R + K - 1, r_first = r0 - P;
  const bool rmw = it.accumulate != 0;
  float xn[NX];
  auto load_row = [&](int rr, float *dst) {
    if (rr >= 0 && rr < H) {
      const float *rowp = xb + ((int64_t)rr * W + (c0 - P)) * in_ld;
#pragma unroll
      for (int j = 0; j < NX; ++j) dst[j] = cok[j] ? rowp[(int64_t)j * in_ld] : 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < NX; ++j) dst[j] = 0.f;
    }
  };
  load_row(r_first, xn);
  for (int i0 = 0; i0 < niter; i0 += K) {
#pragma unroll
    for (int u = 0; u < K; ++u) {
      const int i = i0 + u, rr = r_first + i;
      if (i < niter) {
        float xv[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) xv[j] = xn[j];
        if (i + 1 < niter) load_row(rr + 1, xn);
        const int o = r0 + i - (K - 1);  // output row completed by this step (when i >= K - 1)
        float old[WT];
        if (rmw && i >= K - 1) {
#pragma unroll
          for (int j = 0; j < WT; ++j) old[j] = c0 + j < W ? ob[((int64_t)o * W + c0 + j) * out_ld] : 0.f;
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
#pragma unroll
          for (int j = 0; j < WT; ++j) {
            if (c0 + j < W) {
              float v = acc[sc_][j];
              if (rmw) v += old[j];
              ob[((int64_t)o * W + c0 + j) * out_ld] = v;
              if (STATS) s0 += v, s1 += v * v;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < WT; ++j) acc[sc_][j] = 0.f;
      }
    }
  }
}

// grid = (tiles_x * tiles_y, B), block = 128; a.tiles_x = ceil(W / DwLane<C>::TILE_W)
template <int C, bool STATS>
__global__ void __launch_bounds__(128) dwl_multi_kernel(DwMultiArgs a) {
  constexpr int STRIPS = DwLane<C>::STRIPS;
  __shared__ float s_red[STATS ? 4 : 1][2][32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, n = blockIdx.y;
  const int tx = blockIdx.x % a.tiles_x, ty = blockIdx.x / a.tiles_x;
  const int c0 = tx * DwLane<C>::TILE_W + (warp * STRIPS + lane / C) * DwLane<C>::WT;
  const int r0 = ty * a.tile_rows, r1 = min(r0 + a.tile_rows, a.H), ch = lane % C;
  for (int m = 0; m < a.n; ++m) {
    const DwItem &it = a.it[m];
    float s0 = 0.f, s1 = 0.f;
    if (c0 < a.W) {
      if (it.k == 5) dwl_walk<C, 5, STATS>(it, n, a.H, a.W, c0, r0, r1, ch, s0, s1);
      else dwl_walk<C, 3, STATS>(it, n, a.H, a.W, c0, r0, r1, ch, s0, s1);
    }
    if (STATS) {
      __syncthreads();  // s_red free (previous item)
      s_red[warp][0][lane] = s0, s_red[warp][1][lane] = s1;
      __syncthreads();
      if (tid < 2 * C) {
        const int c = tid % C, which = tid / C;
        float r = 0.f;
        for (int wv = 0; wv < 4; ++wv)
          for (int s = 0; s < STRIPS; ++s) r += s_red[wv][which][s * C + c];
        it.partials[((int64_t)n * gridDim.x + blockIdx.x) * 2 * C + tid] = r;
      }
    }
  }
}

// weight gradient: dW[c][ky][kx] = sum x[o + ky - P][col + kx - P][c] * dz[o][col][c]   (in = x, in2 = dz)
template <int C, int K>
SENAS_DEVFN void dwl_wgrad_walk(const DwItem &it, int n, int H, int W, int c0, int r0, int r1, int ch, float *acc) {
  constexpr int P = K / 2, WT = DwLane<C>::WT, NX = WT + K - 1;
  bool cok[NX];
#pragma unroll
  for (int j = 0; j < NX; ++j) cok[j] = c0 - P + j >= 0 && c0 - P + j < W;
  float dzr[K][WT];  // dz rows in flight (slot = row mod K)
#pragma unroll
  for (int s = 0; s < K; ++s)
#pragma unroll
    for (int j = 0; j < WT; ++j) dzr[s][j] = 0.f;
  const float *xb = it.in + (int64_t)n * H * W * it.in_ld + ch;
  const float *dzb = it.in2 + (int64_t)n * H * W * C + ch;
  const uint16_t *dzh = reinterpret_cast<const uint16_t *>(it.in2) + (int64_t)n * H * W * C + ch;  // bf16-stored dz
  const int dz_bf = it.in2_bf;
  const int64_t in_ld = it.in_ld;
  const int R = r1 - r0, niter = R + K - 1, r_first = r0 - P;
  for (int i0 = 0; i0 < niter; i0 += K) {
#pragma unroll
    for (int u = 0; u < K; ++u) {
      const int i = i0 + u, rr = r_first + i;
      if (i < niter) {
#pragma unroll
        for (int j = 0; j < WT; ++j)  // dz row entering the window: output row r0 + i (slot u)
          dzr[u][j] = (i < R && c0 + j < W) ? (dz_bf ? senas_bits_f((uint32_t)dzh[((int64_t)(r0 + i) * W + c0 + j) * C] << 16)
                                                     : dzb[((int64_t)(r0 + i) * W + c0 + j) * C])
                                            : 0.f;
        if (rr >= 0 && rr < H) {
          float xv[NX];
          const float *rowp = xb + ((int64_t)rr * W + (c0 - P)) * in_ld;
#pragma unroll
          for (int j = 0; j < NX; ++j) xv[j] = cok[j] ? rowp[(int64_t)j * in_ld] : 0.f;
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
}

template <int C>
__global__ void __launch_bounds__(128) dwl_wgrad_multi_kernel(DwMultiArgs a) {
  constexpr int STRIPS = DwLane<C>::STRIPS;
  __shared__ float s_red[4][25][32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, n = blockIdx.y;
  const int tx = blockIdx.x % a.tiles_x, ty = blockIdx.x / a.tiles_x;
  const int c0 = tx * DwLane<C>::TILE_W + (warp * STRIPS + lane / C) * DwLane<C>::WT;
  const int r0 = ty * a.tile_rows, r1 = min(r0 + a.tile_rows, a.H), ch = lane % C;
  const int64_t nblk_idx = (int64_t)n * gridDim.x + blockIdx.x;
  for (int m = 0; m < a.n; ++m) {
    const DwItem &it = a.it[m];
    const int T = it.k * it.k;
    float acc[25];
#pragma unroll
    for (int t = 0; t < 25; ++t) acc[t] = 0.f;
    if (c0 < a.W) {
      if (it.k == 5) dwl_wgrad_walk<C, 5>(it, n, a.H, a.W, c0, r0, r1, ch, acc);
      else dwl_wgrad_walk<C, 3>(it, n, a.H, a.W, c0, r0, r1, ch, acc);
    }
    __syncthreads();  // s_red free
#pragma unroll
    for (int t = 0; t < 25; ++t) s_red[warp][t][lane] = acc[t];
    __syncthreads();
    float *out = it.partials + nblk_idx * C * T;  // [C][T]
    for (int o = tid; o < C * T; o += 128) {
      const int c = o / T, t = o - c * T;
      float r = 0.f;
      for (int wv = 0; wv < 4; ++wv)
        for (int s = 0; s < STRIPS; ++s) r += s_red[wv][t][s * C + c];
      out[o] = r;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fused depthwise-separable candidate (dep_sep_conv_3 / dep_sep_conv_5, utils/operations.py:107-115) for NORM edges:
// the depthwise output z is NEVER written to HBM.  Every sweep recomputes z = dw(x) from the input tile (9 / 25 FMAs per
// element, far cheaper than a 128 B/pixel round trip of a 32-channel fp32 tensor) and keeps it in registers:
//   DS_FWD_STATS : z                                  -> BatchNorm-1 partial sums (sum z, sum z^2 per channel)
//   DS_FWD_Y     : z -> BN1 -> ReLU -> 1x1 (C -> 8)   -> y (8 channels) + BatchNorm-2 partial sums
//   DS_BWD_STATS : z, dy = A g + B y + C              -> sum du, sum du zhat, dW_pw partials ([10C] per block)
//   DS_BWD_DZ    : z, dy                              -> dz = BN1'(relu'(W_pw^T dy))  (fp32, or bf16 in bf16 mode: a plain
//                                                        gradient value, no ReLU decision hangs on its rounding)
// so the only per-candidate tensors in HBM are y (32 B/pixel) and, during backward, dz.  The arithmetic is exact fp32.
//
// Layout: lane = channel.  With NHWC and C = 32 one warp-wide load is the 128 contiguous bytes of one pixel, the
// depthwise weights of a lane's channel (K*K floats) and every BatchNorm-1 quantity sit in registers, and all the
// per-channel reductions (statistics, dW_pw columns) are lane-local: no shuffles until the end of the tile.  A warp walks
// down a strip of kDsCols columns with the K in-flight output rows in registers (slot = row mod K under a K-fold
// unroll): K*K*kDsCols FMAs per (kDsCols + K - 1) loads.  C = 8: four strips side by side in one warp (lane = strip, ch).
// The 1x1 reduces over the channel lanes with a fixed-order transpose-reduce (8 values -> one output channel per lane).
// grid = (tiles_x * tiles_y, B), block = 128 = 4 warps.
// ------------------------------------------------------------------------------------------------
enum { DS_FWD_STATS = 0, DS_FWD_Y = 1, DS_BWD_STATS = 2, DS_BWD_DZ = 3 };
constexpr int kDsMaxItems = 6, kDsCols = 4;
struct DsItem {
  const float *w;                        // depthwise weight [C][K*K]
  const float *mean1, *istd1, *g1, *b1;  // BatchNorm-1 (modes 1..3)
  const float *wpw;                      // pointwise weight [8][C] (modes 1..3)
  float *y;                              // [B][HW][8]: written by mode 1, read by modes 2, 3
  const float *gm;                       // [B][HW][8] node gradient after the ReLU mask (modes 2, 3)
  const float *coef;                     // [3][B][8] dy coefficients A, B, C (modes 2, 3)
  const float *bn1_coef;                 // [3][C] a1, dbeta1/M, dgamma1/M (mode 3)
  float *dz;                             // [B][HW][C] (mode 3)
  float *partials;                       // mode 0: [B][grid.x][2C]; mode 1: [B][grid.x][16]; mode 2: [B][grid.x][10C]
  int32_t k, dz_bf, training, flip;
};
struct DsArgs {
  const float *x;  // [B][H][W][x_ld]
  int64_t x_ld;
  int32_t H, W, tiles_x, tile_rows, n, batch;
  DsItem it[kDsMaxItems];
};
template <int C>
struct DsGeo {
  static constexpr int STRIPS = 32 / C;               // strips per warp
  static constexpr int TILE_W = 4 * STRIPS * kDsCols;  // columns per block
};

// Sum v[0..8) over the C channel lanes of a strip (fixed order); on return v[0] of lane `ch` holds output channel
// ch / (C / 8) (complete in every lane of that group of C / 8 lanes).
template <int C>
SENAS_DEVFN void ds_reduce8(float *v, int lane) {
#pragma unroll
  for (int m = C / 2, nk = 4; nk >= 1; m >>= 1, nk >>= 1) {
    const bool hi = (lane & m) != 0;
#pragma unroll
    for (int j = 0; j < nk; ++j) {
      const float send = hi ? v[j] : v[j + nk], keep = hi ? v[j + nk] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, m);
    }
  }
#pragma unroll
  for (int m = C / 16; m >= 1; m >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], m);
}

template <int C, int K, int MODE>
SENAS_DEVFN void ds_walk(const DsArgs &a, const DsItem &it, int n, int c0, int r0, int r1, int lane, float *acc_out) {
  constexpr int P = K / 2, WT = kDsCols, NX = WT + K - 1, T = K * K, G = C / 8;
  const int ch = lane % C;
  const int H = a.H, W = a.W;
  float w[T];
#pragma unroll
  for (int t = 0; t < T; ++t) w[t] = __ldg(it.w + ch * T + (it.flip ? T - 1 - t : t));
  float sc = 1.f, sh = 0.f, mean1 = 0.f, istd1 = 1.f, wp[8], k0 = 0.f, k1 = 0.f, k2 = 0.f;
#pragma unroll
  for (int co = 0; co < 8; ++co) wp[co] = 0.f;
  if (MODE != DS_FWD_STATS) {
    mean1 = it.mean1[ch], istd1 = it.istd1[ch];
    sc = it.g1[ch] * istd1, sh = it.b1[ch] - mean1 * sc;
#pragma unroll
    for (int co = 0; co < 8; ++co) wp[co] = __ldg(it.wpw + co * C + ch);
  }
  if (MODE == DS_BWD_DZ) k0 = it.bn1_coef[ch], k1 = it.bn1_coef[C + ch], k2 = it.bn1_coef[2 * C + ch];
  float cA[8], cB[8], cC[8];
  if (MODE >= DS_BWD_STATS) {
    const float *cf = it.coef + n * 8;
    const int tot = a.batch * 8;
#pragma unroll
    for (int co = 0; co < 8; ++co) cA[co] = cf[co], cB[co] = cf[tot + co], cC[co] = cf[2 * tot + co];
  }
  bool cok[NX];
#pragma unroll
  for (int j = 0; j < NX; ++j) cok[j] = c0 - P + j >= 0 && c0 - P + j < W;
  float acc[K][WT];
#pragma unroll
  for (int s = 0; s < K; ++s)
#pragma unroll
    for (int j = 0; j < WT; ++j) acc[s][j] = 0.f;
  const float *xb = a.x + (int64_t)n * H * W * a.x_ld + ch;
  const int64_t x_ld = a.x_ld;
  const int R = r1 - r0, niter = R + K - 1, r_first = r0 - P;
  const int co_mine = ch / G;            // output channel this lane holds after ds_reduce8
  const bool owner = (ch % G) == 0;      // one lane of the group stores it
  for (int i0 = 0; i0 < niter; i0 += K) {
#pragma unroll
    for (int u = 0; u < K; ++u) {
      const int i = i0 + u, rr = r_first + i;
      if (i < niter) {
        if (rr >= 0 && rr < H) {
          float xv[NX];
          const float *rowp = xb + ((int64_t)rr * W + (c0 - P)) * x_ld;
#pragma unroll
          for (int j = 0; j < NX; ++j) xv[j] = cok[j] ? rowp[(int64_t)j * x_ld] : 0.f;
#pragma unroll
          for (int ky = 0; ky < K; ++ky) {
            if ((unsigned)(i - ky) < (unsigned)R) {  // output row r0 + i - ky is inside the tile
              const int s = (u - ky + K) % K;
#pragma unroll
              for (int kx = 0; kx < K; ++kx)
#pragma unroll
                for (int j = 0; j < WT; ++j) acc[s][j] = fmaf(xv[kx + j], w[ky * K + kx], acc[s][j]);
            }
          }
        }
        const int sc_ = (u + 1) % K;  // slot of the output row completed by this step
        if (i >= K - 1) {
          const int o = r0 + i - (K - 1);
#pragma unroll
          for (int j = 0; j < WT; ++j) {
            const int col = c0 + j;
            const bool ok = col < W;  // (uniform over the channel lanes of a strip)
            const float z = acc[sc_][j];
            const int64_t pix = ((int64_t)n * H + o) * W + (ok ? col : 0);
            if (MODE == DS_FWD_STATS) {
              if (ok) acc_out[0] += z, acc_out[1] += z * z;
            } else if (MODE == DS_FWD_Y) {
              const float r = fmaxf(fmaf(z, sc, sh), 0.f);
              float v[8];
#pragma unroll
              for (int co = 0; co < 8; ++co) v[co] = r * wp[co];
              ds_reduce8<C>(v, lane);
              if (ok && owner) {
                it.y[pix * 8 + co_mine] = v[0];
                acc_out[0] += v[0], acc_out[1] += v[0] * v[0];
              }
            } else {
              const float4 glo = ld4(it.gm + pix * 8), ghi = ld4(it.gm + pix * 8 + 4);
              const float4 ylo = ld4(it.y + pix * 8), yhi = ld4(it.y + pix * 8 + 4);
              const float g8[8] = {glo.x, glo.y, glo.z, glo.w, ghi.x, ghi.y, ghi.z, ghi.w};
              const float y8[8] = {ylo.x, ylo.y, ylo.z, ylo.w, yhi.x, yhi.y, yhi.z, yhi.w};
              float dy[8], dr = 0.f;
#pragma unroll
              for (int co = 0; co < 8; ++co) {
                dy[co] = fmaf(cA[co], g8[co], fmaf(cB[co], y8[co], cC[co]));
                dr = fmaf(dy[co], wp[co], dr);
              }
              const float uu = fmaf(z, sc, sh);
              const float du = (uu > 0.f && ok) ? dr : 0.f;
              const float zhat = (z - mean1) * istd1;
              if (MODE == DS_BWD_STATS) {
                const float r = ok ? fmaxf(uu, 0.f) : 0.f;
                acc_out[0] += du, acc_out[1] += du * zhat;
#pragma unroll
                for (int co = 0; co < 8; ++co) acc_out[2 + co] = fmaf(dy[co], r, acc_out[2 + co]);
              } else {
                const float dzv = it.training ? k0 * (du - k1 - zhat * k2) : k0 * du;
                if (it.dz_bf) {  // pack two channels per 32-bit store
                  const float nb = __shfl_down_sync(0xffffffffu, dzv, 1);
                  if (ok && !(lane & 1))
                    reinterpret_cast<uint32_t *>(it.dz)[(pix * C + ch) >> 1] = senas_pack_bf2(dzv, nb);
                } else if (ok) {
                  it.dz[pix * C + ch] = dzv;
                }
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < WT; ++j) acc[sc_][j] = 0.f;
      }
    }
  }
}

template <int C, int MODE>
__global__ void __launch_bounds__(128, 3) ds_norm_kernel(DsArgs a) {
  constexpr int STRIPS = DsGeo<C>::STRIPS, NACC = MODE == DS_BWD_STATS ? 10 : 2;
  __shared__ float s_red[4][NACC][32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, n = blockIdx.y;
  const int tx = blockIdx.x % a.tiles_x, ty = blockIdx.x / a.tiles_x;
  const int strip = warp * STRIPS + lane / C;
  const int c0 = tx * DsGeo<C>::TILE_W + strip * kDsCols;
  const int r0 = ty * a.tile_rows, r1 = min(r0 + a.tile_rows, a.H);
  for (int m = 0; m < a.n; ++m) {
    const DsItem &it = a.it[m];
    float acc[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) acc[j] = 0.f;
    // (strips that start beyond the right edge still walk: their lanes take part in the shuffles, all stores are masked)
    if (it.k == 5) ds_walk<C, 5, MODE>(a, it, n, c0, r0, r1, lane, acc);
    else ds_walk<C, 3, MODE>(a, it, n, c0, r0, r1, lane, acc);
    if (MODE == DS_BWD_DZ) continue;
    __syncthreads();  // s_red free (previous item)
#pragma unroll
    for (int j = 0; j < NACC; ++j) s_red[warp][j][lane] = acc[j];
    __syncthreads();
    float *out = it.partials + ((int64_t)n * gridDim.x + blockIdx.x) * (MODE == DS_FWD_STATS ? 2 * C : (MODE == DS_FWD_Y ? 16 : 10 * C));
    if (MODE == DS_FWD_Y) {  // acc[0..1] = sum y, sum y^2 of output channel co in the owner lanes (ch % (C/8) == 0)
      if (tid < 16) {
        const int co = tid & 7, which = tid >> 3;
        float r = 0.f;
        for (int wv = 0; wv < 4; ++wv)
          for (int s = 0; s < STRIPS; ++s) r += s_red[wv][which][s * C + co * (C / 8)];
        out[tid] = r;
      }
    } else {  // per-channel values: combine the strips of a warp and the 4 warps in fixed order
      for (int o = tid; o < NACC * C; o += 128) {
        const int j = o / C, c = o - j * C;
        float r = 0.f;
        for (int wv = 0; wv < 4; ++wv)
          for (int s = 0; s < STRIPS; ++s) r += s_red[wv][j][s * C + c];
        // layouts: [sum z | sum z^2] and [sum du | sum du zhat | dW_pw[co][c]]
        out[j * C + c] = r;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// dep-sep pointwise backward, quad layout (same as pw_fwd_kernel): LP = C/4 lanes share a pixel, lane l owns channels
// 4l..4l+3 of z (one coalesced float4 per lane), the 8 dy values of the pixel are computed once by the group and
// broadcast with shuffles, W_pw[:, 4l..4l+3] lives in registers.
//   PASS 1: per-channel sums of du and du*zhat and dW_pw partials ([10C] per block);   PASS 2: dz in place over z.
// ------------------------------------------------------------------------------------------------
template <int C, int PASS, int UN = (PASS == 1 ? 2 : 4)>  // UN pixel steps in flight; pass 1 carries 40 accumulators
__global__ void __launch_bounds__(256, 2) pw_bwd_q_kernel(PwBwdArgs a, int px_per_block, int training) {
  constexpr int LP = C / 4, PPW = 32 / LP, NDY = 8 / LP;
  __shared__ float s_part[PASS == 1 ? 8 : 1][PASS == 1 ? 10 * C : 1];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, n = blockIdx.y;
  const int l = lane % LP, sub = lane / LP;
  float sc[4], sh[4], mean1[4], istd1[4], wq[8][4], k0[4], k1[4], k2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = l * 4 + j;
    mean1[j] = a.mean1[c], istd1[j] = a.istd1[c];
    sc[j] = a.g1[c] * istd1[j], sh[j] = a.b1[c] - mean1[j] * sc[j];
#pragma unroll
    for (int co = 0; co < 8; ++co) wq[co][j] = __ldg(a.wpw + co * C + c);
    k0[j] = k1[j] = k2[j] = 0.f;
    if (PASS == 2) k0[j] = a.bn1_coef[c], k1[j] = a.bn1_coef[C + c], k2[j] = a.bn1_coef[2 * C + c];
  }
  float cA[NDY], cB[NDY], cC[NDY];
#pragma unroll
  for (int j = 0; j < NDY; ++j) {
    const int co = l * NDY + j;
    cA[j] = a.coefA[n * 8 + co], cB[j] = a.coefB[n * 8 + co], cC[j] = a.coefC[n * 8 + co];
  }
  float s_du[4], s_duz[4], dw[8][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s_du[j] = s_duz[j] = 0.f;
#pragma unroll
    for (int co = 0; co < 8; ++co) dw[co][j] = 0.f;
  }
  const int p_begin = blockIdx.x * px_per_block, p_end = min(p_begin + px_per_block, a.hw);
  const int per_warp = (px_per_block + 7) / 8;
  const int w_begin = p_begin + warp * per_warp, w_end = min(w_begin + per_warp, p_end);
  const int64_t z0 = (int64_t)n * a.hw * C + l * 4;
  const int z_bf = a.z_bf;
  const float *gn = a.gm + (int64_t)n * a.hw * 8 + l * NDY, *yn = a.y + (int64_t)n * a.hw * 8 + l * NDY;
  for (int p0 = w_begin; p0 < w_end; p0 += PPW * UN) {
    float4 zv[UN];
    float dyl[UN][NDY];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int p = p0 + u * PPW + sub;
      const bool ok = p < w_end;
      const int pp = ok ? p : w_begin;
      zv[u] = ldx4(a.z, z0 + (int64_t)pp * C, z_bf);
      if (NDY == 4) {
        const float4 g4 = ld4(gn + (int64_t)pp * 8), y4 = ld4(yn + (int64_t)pp * 8);
        dyl[u][0] = cA[0] * g4.x + cB[0] * y4.x + cC[0];
        dyl[u][1 % NDY] = cA[1 % NDY] * g4.y + cB[1 % NDY] * y4.y + cC[1 % NDY];
        dyl[u][2 % NDY] = cA[2 % NDY] * g4.z + cB[2 % NDY] * y4.z + cC[2 % NDY];
        dyl[u][3 % NDY] = cA[3 % NDY] * g4.w + cB[3 % NDY] * y4.w + cC[3 % NDY];
      } else {
        dyl[u][0] = cA[0] * gn[(int64_t)pp * 8] + cB[0] * yn[(int64_t)pp * 8] + cC[0];
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int p = p0 + u * PPW + sub;
      const bool ok = p < w_end;
      float dy[8];
#pragma unroll
      for (int co = 0; co < 8; ++co) dy[co] = __shfl_sync(0xffffffffu, dyl[u][co % NDY], sub * LP + co / NDY);
      const float z[4] = {zv[u].x, zv[u].y, zv[u].z, zv[u].w};
      float dzo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float uu = fmaf(z[j], sc[j], sh[j]);
        float dr = 0.f;
#pragma unroll
        for (int co = 0; co < 8; ++co) dr = fmaf(dy[co], wq[co][j], dr);
        const float du = (uu > 0.f && ok) ? dr : 0.f;
        const float zhat = (z[j] - mean1[j]) * istd1[j];
        if (PASS == 1) {
          const float r = ok ? fmaxf(uu, 0.f) : 0.f;
          s_du[j] += du, s_duz[j] += du * zhat;
#pragma unroll
          for (int co = 0; co < 8; ++co) dw[co][j] = fmaf(dy[co], r, dw[co][j]);
        } else {
          dzo[j] = training ? k0[j] * (du - k1[j] - zhat * k2[j]) : k0[j] * du;
        }
      }
      if (PASS == 2 && ok) stx4(a.z, z0 + (int64_t)p * C, make_float4(dzo[0], dzo[1], dzo[2], dzo[3]), z_bf);
    }
  }
  if (PASS == 1) {
    // combine the PPW pixel groups of a warp (lanes with equal l), then the 8 warps, in fixed order
#pragma unroll
    for (int m = LP; m < 32; m <<= 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s_du[j] += __shfl_xor_sync(0xffffffffu, s_du[j], m), s_duz[j] += __shfl_xor_sync(0xffffffffu, s_duz[j], m);
#pragma unroll
        for (int co = 0; co < 8; ++co) dw[co][j] += __shfl_xor_sync(0xffffffffu, dw[co][j], m);
      }
    }
    if (lane < LP) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = l * 4 + j;
        s_part[warp][c] = s_du[j], s_part[warp][C + c] = s_duz[j];
#pragma unroll
        for (int co = 0; co < 8; ++co) s_part[warp][2 * C + co * C + c] = dw[co][j];
      }
    }
    __syncthreads();
    float *out = a.partials + ((int64_t)n * gridDim.x + blockIdx.x) * 10 * C;
    for (int o = tid; o < 10 * C; o += 256) {
      float r = 0.f;
      for (int wv = 0; wv < 8; ++wv) r += s_part[wv][o];
      out[o] = r;
    }
  }
}


// ------------------------------------------------------------------------------------------------
// AdapterBlock with a 1x1 conv on 32 channels (identity on NORM edges, up_sample on UP edges), quad layout.
// up_sample: the 1x1 conv commutes with the bilinear interpolation (both linear), so the conv runs on the LOW-resolution
// grid (u = W.x, pw_fwd_kernel<C, true>) and only 8 channels are interpolated (up8_fwd_kernel: y, statistics); backward
// mirrors it: du = bilinear^T(dy) once (up8_bwd_kernel), then dx += W^T.du and dW = sum du (x) x in ONE sweep over the
// low-resolution grid (lin_bwd_q_kernel) -- the first version gathered the 4x4 dy window twice per input pixel.
// ------------------------------------------------------------------------------------------------
struct Up8Args {
  const float *u;  // [B][h][w][8] low resolution
  float *y;        // [B][2h][2w][8]
  int32_t h, w;
  float *partials;  // [B][gridDim.x][16]
};
__global__ void __launch_bounds__(128) up8_fwd_kernel(Up8Args a) {
  const int n = blockIdx.y, p = blockIdx.x * 128 + threadIdx.x, oh = 2 * a.h, ow = 2 * a.w;
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = 0.f;
  if (p < oh * ow) {
    const int oy = p / ow, ox = p - oy * ow;
    int y0, y1, x0, x1;
    float ly, lx;
    up_src(oy, a.h, y0, y1, ly);
    up_src(ox, a.w, x0, x1, lx);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const float *un = a.u + (int64_t)n * a.h * a.w * 8;
    const float *p00 = un + ((int64_t)y0 * a.w + x0) * 8, *p01 = un + ((int64_t)y0 * a.w + x1) * 8;
    const float *p10 = un + ((int64_t)y1 * a.w + x0) * 8, *p11 = un + ((int64_t)y1 * a.w + x1) * 8;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; c += 4) {
      const float4 q0 = ld4(p00 + c), q1 = ld4(p01 + c), q2 = ld4(p10 + c), q3 = ld4(p11 + c);
      acc[c] = w00 * q0.x + w01 * q1.x + w10 * q2.x + w11 * q3.x;
      acc[c + 1] = w00 * q0.y + w01 * q1.y + w10 * q2.y + w11 * q3.y;
      acc[c + 2] = w00 * q0.z + w01 * q1.z + w10 * q2.z + w11 * q3.z;
      acc[c + 3] = w00 * q0.w + w01 * q1.w + w10 * q2.w + w11 * q3.w;
    }
    float *yp = a.y + ((int64_t)n * oh * ow + p) * 8;
    st4(yp, make_float4(acc[0], acc[1], acc[2], acc[3]));
    st4(yp + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = acc[j], v[8 + j] = acc[j] * acc[j];
  }
  block_sum_store<16>(v, a.partials + ((int64_t)n * gridDim.x + blockIdx.x) * 16);
}

// du[B][x_h][x_w][8] = bilinear^T(dy), dy = A*gm + B*y + C on the output grid.  thread = low-resolution pixel.
__global__ void __launch_bounds__(128) up8_bwd_kernel(AdapterBwdArgs a, float *du) {
  const int n = blockIdx.y, p = blockIdx.x * 128 + threadIdx.x;
  if (p >= a.x_h * a.x_w) return;
  const int iy = p / a.x_w, ix = p - iy * a.x_w;
  float s[8];
  adapter_gather_dy<AD_UP>(a, n, iy, ix, s);
  float *o = du + ((int64_t)n * a.x_h * a.x_w + p) * 8;
  st4(o, make_float4(s[0], s[1], s[2], s[3]));
  st4(o + 4, make_float4(s[4], s[5], s[6], s[7]));
}

struct LinBwdArgs {
  const float *x;  // [B][hw][x_ld]
  int64_t x_ld;
  const float *gm, *y;  // affine: dy = A*gm + B*y + C (gm, y: [B][hw][8]);  plain (coefA == null): dy = gm
  const float *coefA, *coefB, *coefC;
  const float *w;  // [8][C]
  float *dx;       // [B][hw][dx_ld], (+)=; null to skip
  int64_t dx_ld;
  int32_t accumulate, hw;
  float *partials;  // [B][gridDim.x][8C] (co-major), null to skip the weight gradient
};
template <int C>
__global__ void __launch_bounds__(256, 2) lin_bwd_q_kernel(LinBwdArgs a, int px_per_block) {
  constexpr int LP = C / 4, PPW = 32 / LP, NDY = 8 / LP, UN = 2;
  __shared__ float s_part[8][8 * C];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, n = blockIdx.y;
  const int l = lane % LP, sub = lane / LP;
  const bool affine = a.coefA != nullptr;
  float wq[8][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int co = 0; co < 8; ++co) wq[co][j] = __ldg(a.w + co * C + l * 4 + j);
  float cA[NDY], cB[NDY], cC[NDY];
#pragma unroll
  for (int j = 0; j < NDY; ++j) {
    const int co = l * NDY + j;
    cA[j] = affine ? a.coefA[n * 8 + co] : 1.f, cB[j] = affine ? a.coefB[n * 8 + co] : 0.f;
    cC[j] = affine ? a.coefC[n * 8 + co] : 0.f;
  }
  float dw[8][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int co = 0; co < 8; ++co) dw[co][j] = 0.f;
  const int p_begin = blockIdx.x * px_per_block, p_end = min(p_begin + px_per_block, a.hw);
  const int per_warp = (px_per_block + 7) / 8;
  const int w_begin = p_begin + warp * per_warp, w_end = min(w_begin + per_warp, p_end);
  const float *xn = a.x + (int64_t)n * a.hw * a.x_ld + l * 4;
  const float *gn = a.gm + (int64_t)n * a.hw * 8 + l * NDY;
  const float *yn = affine ? a.y + (int64_t)n * a.hw * 8 + l * NDY : gn;
  float *dxn = a.dx ? a.dx + (int64_t)n * a.hw * a.dx_ld + l * 4 : nullptr;
  for (int p0 = w_begin; p0 < w_end; p0 += PPW * UN) {
    float4 xv[UN], ov[UN];
    float dyl[UN][NDY];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int p = p0 + u * PPW + sub;
      const bool ok = p < w_end;
      const int pp = ok ? p : w_begin;
      xv[u] = ld4(xn + (int64_t)pp * a.x_ld);
      ov[u] = (dxn && a.accumulate) ? ld4(dxn + (int64_t)pp * a.dx_ld) : f4zero();
      if (NDY == 4) {
        const float4 g4 = ld4(gn + (int64_t)pp * 8), y4 = affine ? ld4(yn + (int64_t)pp * 8) : f4zero();
        dyl[u][0] = cA[0] * g4.x + cB[0] * y4.x + cC[0];
        dyl[u][1 % NDY] = cA[1 % NDY] * g4.y + cB[1 % NDY] * y4.y + cC[1 % NDY];
        dyl[u][2 % NDY] = cA[2 % NDY] * g4.z + cB[2 % NDY] * y4.z + cC[2 % NDY];
        dyl[u][3 % NDY] = cA[3 % NDY] * g4.w + cB[3 % NDY] * y4.w + cC[3 % NDY];
      } else {
        dyl[u][0] = cA[0] * gn[(int64_t)pp * 8] + (affine ? cB[0] * yn[(int64_t)pp * 8] : 0.f) + cC[0];
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int p = p0 + u * PPW + sub;
      const bool ok = p < w_end;
      float dy[8];
#pragma unroll
      for (int co = 0; co < 8; ++co) dy[co] = __shfl_sync(0xffffffffu, dyl[u][co % NDY], sub * LP + co / NDY);
      const float xx[4] = {ok ? xv[u].x : 0.f, ok ? xv[u].y : 0.f, ok ? xv[u].z : 0.f, ok ? xv[u].w : 0.f};
      float dr[4] = {ov[u].x, ov[u].y, ov[u].z, ov[u].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int co = 0; co < 8; ++co) {
          dr[j] = fmaf(dy[co], wq[co][j], dr[j]);
          dw[co][j] = fmaf(dy[co], xx[j], dw[co][j]);
        }
      }
      if (dxn && ok) st4(dxn + (int64_t)p * a.dx_ld, make_float4(dr[0], dr[1], dr[2], dr[3]));
    }
  }
  if (a.partials != nullptr) {
#pragma unroll
    for (int m = LP; m < 32; m <<= 1)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int co = 0; co < 8; ++co) dw[co][j] += __shfl_xor_sync(0xffffffffu, dw[co][j], m);
    if (lane < LP) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int co = 0; co < 8; ++co) s_part[warp][co * C + l * 4 + j] = dw[co][j];
    }
    __syncthreads();
    float *out = a.partials + ((int64_t)n * gridDim.x + blockIdx.x) * 8 * C;
    for (int o = tid; o < 8 * C; o += 256) {
      float r = 0.f;
      for (int wv = 0; wv < 8; ++wv) r += s_part[wv][o];
      out[o] = r;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Depthwise transposed convolution (UP edges: ConvTranspose2d k, stride 2, pad k/2, output_padding 1, groups = C),
// several convolutions of ONE input per launch, on the input grid.  Per axis kernel index kk reaches output parity
// p(kk) = (P + kk) & 1 from input offset d(kk) = (p + P - kk) / 2, i.e. output o = 2 i + kk - P:
//   forward : z[2i + py(ky)][2j + px(kx)] += x[i + d(ky)][j + d(kx)] * w[ky][kx]     (4 output pixels per input pixel)
//   dx      : dx[i][j] (+)= sum_{ky,kx} dz[2i + ky - P][2j + kx - P] * w[ky][kx]     (stride-2 gather)
//   dW      : dW[ky][kx] = sum_{i,j} x[i + d(ky)][j + d(kx)] * dz[2i + py(ky)][2j + px(kx)]
// thread = (channel quad, input column); same tiling, smem weight layout and item list as dw_multi_kernel.
// ------------------------------------------------------------------------------------------------
template <int K>
SENAS_DEVFN constexpr int up_par(int kk) { return (K / 2 + kk) & 1; }
template <int K>
SENAS_DEVFN constexpr int up_off(int kk) { return (up_par<K>(kk) + K / 2 - kk) / 2; }  // in {-1, 0, 1}

// x window rows i-1, i, i+1 and columns j-1, j, j+1 (zero outside the image) -> xw[3][3]
SENAS_DEVFN void up_load_window(const float *in, int64_t in0, int bf, int64_t in_ld, int H, int W, int i, int j,
                                float4 (&xw)[3][3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int iy = i + r - 1;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int ix = j + c - 1;
      xw[r][c] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? ldx4(in, in0 + ((int64_t)iy * W + ix) * in_ld, bf) : f4zero();
    }
  }
}

template <int C, int K, bool STATS>
SENAS_DEVFN void dw_up_fwd_rows(const DwItem &it, const float *s_w, int n, int H, int W, int by0, int by1, int j, int q,
                                float *st) {
  const int64_t in0 = (int64_t)n * H * W * it.in_ld + q * 4;
  const int64_t out0 = (int64_t)n * 4 * H * W * it.out_ld + q * 4;  // [2H][2W][out_ld]
  const float *wq = s_w + q * 4;
  const int out_bf = it.out_bf;
  for (int i = by0; i < by1; ++i) {
    float4 xw[3][3];
    up_load_window(it.in, in0, it.in_bf, it.in_ld, H, W, i, j, xw);
    float4 acc[2][2];
    acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = f4zero();
#pragma unroll
    for (int ky = 0; ky < K; ++ky)
#pragma unroll
      for (int kx = 0; kx < K; ++kx)
        fma4(acc[up_par<K>(ky)][up_par<K>(kx)], xw[up_off<K>(ky) + 1][up_off<K>(kx) + 1], ld4(wq + (ky * K + kx) * C));
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        float4 v = acc[py][px];
        const int64_t oe = out0 + ((int64_t)(2 * i + py) * (2 * W) + 2 * j + px) * it.out_ld;
        if (it.accumulate) {
          const float4 old = ldx4(it.out, oe, out_bf);
          v.x += old.x, v.y += old.y, v.z += old.z, v.w += old.w;
        }
        stx4(it.out, oe, v, out_bf);
        if (STATS) {
          v = roundx4(v, out_bf);
          st[0] += v.x, st[1] += v.y, st[2] += v.z, st[3] += v.w;
          st[4] += v.x * v.x, st[5] += v.y * v.y, st[6] += v.z * v.z, st[7] += v.w * v.w;
        }
      }
  }
}

template <int C, int K, bool STATS>
SENAS_DEVFN void dw_up_dx_rows(const DwItem &it, const float *s_w, int n, int H, int W, int by0, int by1, int j, int q,
                               float *st) {
  constexpr int P = K / 2;
  const int64_t dz0 = (int64_t)n * 4 * H * W * it.in_ld + q * 4;  // high-resolution input [2H][2W][in_ld]
  const int64_t out0 = (int64_t)n * H * W * it.out_ld + q * 4;
  const float *wq = s_w + q * 4;
  const int OH = 2 * H, OW = 2 * W;
  const int in_bf = it.in_bf, out_bf = it.out_bf;
  for (int i = by0; i < by1; ++i) {
    const int64_t oe = out0 + ((int64_t)i * W + j) * it.out_ld;
    float4 acc = it.accumulate ? ldx4(it.out, oe, out_bf) : f4zero();
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      const int oy = 2 * i + ky - P;
      if (oy < 0 || oy >= OH) continue;
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int ox = 2 * j + kx - P;
        if (ox < 0 || ox >= OW) continue;
        fma4(acc, ldx4(it.in, dz0 + ((int64_t)oy * OW + ox) * it.in_ld, in_bf), ld4(wq + (ky * K + kx) * C));
      }
    }
    stx4(it.out, oe, acc, out_bf);
    if (STATS) {
      acc = roundx4(acc, out_bf);
      st[0] += acc.x, st[1] += acc.y, st[2] += acc.z, st[3] += acc.w;
      st[4] += acc.x * acc.x, st[5] += acc.y * acc.y, st[6] += acc.z * acc.z, st[7] += acc.w * acc.w;
    }
  }
}

// SCATTER: low -> high resolution (4 output pixels per low-resolution pixel), else the stride-2 gather high -> low.
//   UP   forward <C, true, true>,  data gradient <C, false, false>
//   DOWN forward <C, false, true> (z[o] = sum x[2o + k - P] w[k]),  data gradient <C, true, false> (+= into dx)
// a.H, a.W = the LOW-resolution grid.  grid = (tiles_x * tiles_y, B), block = 128 = Q quads x 128/Q columns
template <int C, bool SCATTER, bool STATS>
__global__ void __launch_bounds__(128) dw_up_multi_kernel(DwMultiArgs a) {
  constexpr int Q = C / 4, SLOTS = 128 / Q;
  __shared__ float s_red[STATS ? 128 : 1][8];
  __shared__ float4 s_w4[25 * C / 4];
  float *s_w = reinterpret_cast<float *>(s_w4);
  const int tid = threadIdx.x, q = tid % Q, slot = tid / Q, n = blockIdx.y;
  const int tx = blockIdx.x % a.tiles_x, ty = blockIdx.x / a.tiles_x;
  const int j = tx * SLOTS + slot, by0 = ty * a.tile_rows, by1 = min(by0 + a.tile_rows, a.H);
  const bool active = j < a.W;
  for (int m = 0; m < a.n; ++m) {
    const DwItem &it = a.it[m];
    const int T = it.k * it.k;
    __syncthreads();
    for (int i = tid; i < T * C; i += 128) {
      const int c = i % C, t = i / C;
      s_w[i] = __ldg(it.w + c * T + t);
    }
    __syncthreads();
    float st[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (active) {
      if (SCATTER) {
        if (it.k == 5) dw_up_fwd_rows<C, 5, STATS>(it, s_w, n, a.H, a.W, by0, by1, j, q, st);
        else dw_up_fwd_rows<C, 3, STATS>(it, s_w, n, a.H, a.W, by0, by1, j, q, st);
      } else {
        if (it.k == 5) dw_up_dx_rows<C, 5, STATS>(it, s_w, n, a.H, a.W, by0, by1, j, q, st);
        else dw_up_dx_rows<C, 3, STATS>(it, s_w, n, a.H, a.W, by0, by1, j, q, st);
      }
    }
    if (STATS) {
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) s_red[tid][jj] = st[jj];
      __syncthreads();
      if (tid < 2 * C) {
        const int c = tid % C, which = tid / C;
        float r = 0.f;
        for (int s = 0; s < SLOTS; ++s) r += s_red[s * Q + c / 4][which * 4 + (c & 3)];
        it.partials[((int64_t)n * gridDim.x + blockIdx.x) * 2 * C + tid] = r;
      }
    }
  }
}

// weight gradient of the same group (in = low-resolution operand, in2 = high-resolution operand: UP x / dz, DOWN dz / x);
// partial layout [C][K*K] per block like dw_wgrad_multi_kernel
template <int C, int K>
SENAS_DEVFN void dw_up_wgrad_rows(const DwItem &it, int n, int H, int W, int by0, int by1, int j, int q, bool active,
                                  float *s_part, int nblk_idx) {
  constexpr int T = K * K, Q = C / 4;
  float4 acc[T];
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t] = f4zero();
  if (active) {
    const int64_t ld2 = it.in2_ld ? it.in2_ld : C;
    const int64_t in0 = (int64_t)n * H * W * it.in_ld + q * 4, dz0 = (int64_t)n * 4 * H * W * ld2 + q * 4;
    const int in2_bf = it.in2_bf;
    for (int i = by0; i < by1; ++i) {
      float4 xw[3][3], dz[2][2];
      up_load_window(it.in, in0, it.in_bf, it.in_ld, H, W, i, j, xw);
#pragma unroll
      for (int py = 0; py < 2; ++py)
#pragma unroll
        for (int px = 0; px < 2; ++px)
          dz[py][px] = ldx4(it.in2, dz0 + ((int64_t)(2 * i + py) * (2 * W) + 2 * j + px) * ld2, in2_bf);
#pragma unroll
      for (int ky = 0; ky < K; ++ky)
#pragma unroll
        for (int kx = 0; kx < K; ++kx)
          fma4(acc[ky * K + kx], xw[up_off<K>(ky) + 1][up_off<K>(kx) + 1], dz[up_par<K>(ky)][up_par<K>(kx)]);
    }
  }
#pragma unroll
  for (int t = 0; t < T; ++t) {
#pragma unroll
    for (int m = Q; m < 32; m <<= 1) {
      acc[t].x += __shfl_xor_sync(0xffffffffu, acc[t].x, m), acc[t].y += __shfl_xor_sync(0xffffffffu, acc[t].y, m);
      acc[t].z += __shfl_xor_sync(0xffffffffu, acc[t].z, m), acc[t].w += __shfl_xor_sync(0xffffffffu, acc[t].w, m);
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane < Q) {
#pragma unroll
    for (int t = 0; t < T; ++t) {
      float *o = s_part + warp * (C * 25) + (lane * 4) * T + t;
      o[0] = acc[t].x, o[T] = acc[t].y, o[2 * T] = acc[t].z, o[3 * T] = acc[t].w;
    }
  }
  __syncthreads();
  float *out = it.partials + (int64_t)nblk_idx * C * T;
  for (int o = threadIdx.x; o < C * T; o += 128)
    out[o] = (s_part[o] + s_part[C * 25 + o]) + (s_part[2 * C * 25 + o] + s_part[3 * C * 25 + o]);
}

template <int C>
__global__ void __launch_bounds__(128) dw_up_wgrad_multi_kernel(DwMultiArgs a) {
  constexpr int Q = C / 4, SLOTS = 128 / Q;
  __shared__ float s_part[4 * C * 25];
  const int tid = threadIdx.x, q = tid % Q, slot = tid / Q, n = blockIdx.y;
  const int tx = blockIdx.x % a.tiles_x, ty = blockIdx.x / a.tiles_x;
  const int j = tx * SLOTS + slot, by0 = ty * a.tile_rows, by1 = min(by0 + a.tile_rows, a.H);
  const bool active = j < a.W;
  const int nblk_idx = n * gridDim.x + blockIdx.x;
  for (int m = 0; m < a.n; ++m) {
    const DwItem &it = a.it[m];
    if (it.k == 5) dw_up_wgrad_rows<C, 5>(it, n, a.H, a.W, by0, by1, j, q, active, s_part, nblk_idx);
    else dw_up_wgrad_rows<C, 3>(it, n, a.H, a.W, by0, by1, j, q, active, s_part, nblk_idx);
  }
}

// ------------------------------------------------------------------------------------------------
// AvgPool2d(3, stride 2, padding 1, count_include_pad=False) on NHWC fp32, forward and backward: the pooling of the
// down cells' `preprocess0` (build_rectify, utils/operations.py:141-152).  PyTorch 2.11's CUDA backward of
// this op is wrong for channels_last tensors and its NCHW fallback cost 1.7 ms per step (8 launches) plus two layout
// conversions each.  thread = (pixel, channel quad).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const float *x, int64_t x_ld, float *y, int B, int H, int W, int C) {
  const int Q = C / 4, Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x, total = (int64_t)B * Ho * Wo * Q;
  if (i >= total) return;
  const int q = (int)(i % Q);
  const int64_t pix = i / Q;
  const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), n = (int)(pix / ((int64_t)Wo * Ho));
  const float *xn = x + (int64_t)n * H * W * x_ld + q * 4;
  float4 acc = f4zero();
  int cnt = 0;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int iy = 2 * oy + dy, ix = 2 * ox + dx;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      ++cnt;
      const float4 v = ld4(xn + ((int64_t)iy * W + ix) * x_ld);
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
  const float inv = 1.f / (float)cnt;
  st4(y + pix * C + q * 4, make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv));
}

__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const float *gy, float *gx, int B, int H, int W, int C) {
  const int Q = C / 4, Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x, total = (int64_t)B * H * W * Q;
  if (i >= total) return;
  const int q = (int)(i % Q);
  const int64_t pix = i / Q;
  const int ix = (int)(pix % W), iy = (int)((pix / W) % H), n = (int)(pix / ((int64_t)W * H));
  const float *gn = gy + (int64_t)n * Ho * Wo * C + q * 4;
  float4 acc = f4zero();
  for (int oy = iy >> 1; oy <= (iy + 1) >> 1; ++oy) {  // output windows that contain input row iy
    if (oy >= Ho) continue;
    for (int ox = ix >> 1; ox <= (ix + 1) >> 1; ++ox) {
      if (ox >= Wo) continue;
      const float inv = 1.f / (float)(pool_cnt(oy, H) * pool_cnt(ox, W));
      const float4 v = ld4(gn + ((int64_t)oy * Wo + ox) * C);
      acc.x = fmaf(v.x, inv, acc.x), acc.y = fmaf(v.y, inv, acc.y), acc.z = fmaf(v.z, inv, acc.z), acc.w = fmaf(v.w, inv, acc.w);
    }
  }
  st4(gx + pix * C + q * 4, acc);
}

// avg_pool AdapterBlock on 32 channels (DOWN edges): the 1x1 conv commutes with the pooling as it does with the bilinear
// interpolation, so u = W.x runs on the input grid in quad layout (pw_fwd_kernel<32, true>) and only 8 channels are
// pooled; backward: du = pool^T(dy) once (8 channels on the input grid), then lin_bwd_q_kernel.
struct Pool8Args {
  const float *u;  // [B][h][w][8] input resolution
  float *y;        // [B][oh][ow][8], oh = ceil(h/2)
  int32_t h, w, oh, ow;
  float *partials;  // [B][gridDim.x][16]
};
__global__ void __launch_bounds__(128) pool8_fwd_kernel(Pool8Args a) {
  const int n = blockIdx.y, p = blockIdx.x * 128 + threadIdx.x;
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = 0.f;
  if (p < a.oh * a.ow) {
    const int oy = p / a.ow, ox = p - oy * a.ow;
    const float *un = a.u + (int64_t)n * a.h * a.w * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int cnt = 0;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int iy = 2 * oy + dy, ix = 2 * ox + dx;
        if (iy < 0 || iy >= a.h || ix < 0 || ix >= a.w) continue;
        ++cnt;
        const float *q = un + ((int64_t)iy * a.w + ix) * 8;
        const float4 lo = ld4(q), hi = ld4(q + 4);
        acc[0] += lo.x, acc[1] += lo.y, acc[2] += lo.z, acc[3] += lo.w;
        acc[4] += hi.x, acc[5] += hi.y, acc[6] += hi.z, acc[7] += hi.w;
      }
    const float inv = 1.f / (float)cnt;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    float *yp = a.y + ((int64_t)n * a.oh * a.ow + p) * 8;
    st4(yp, make_float4(acc[0], acc[1], acc[2], acc[3]));
    st4(yp + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = acc[j], v[8 + j] = acc[j] * acc[j];
  }
  block_sum_store<16>(v, a.partials + ((int64_t)n * gridDim.x + blockIdx.x) * 16);
}

// du[B][x_h][x_w][8] = pool^T(dy), dy = A*gm + B*y + C on the output grid.  thread = input pixel.
__global__ void __launch_bounds__(128) pool8_bwd_kernel(AdapterBwdArgs a, float *du) {
  const int n = blockIdx.y, p = blockIdx.x * 128 + threadIdx.x;
  if (p >= a.x_h * a.x_w) return;
  const int iy = p / a.x_w, ix = p - iy * a.x_w;
  float s[8];
  adapter_gather_dy<AD_POOL>(a, n, iy, ix, s);
  float *o = du + ((int64_t)n * a.x_h * a.x_w + p) * 8;
  st4(o, make_float4(s[0], s[1], s[2], s[3]));
  st4(o + 4, make_float4(s[4], s[5], s[6], s[7]));
}
