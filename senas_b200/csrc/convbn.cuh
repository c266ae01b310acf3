// convbn.cuh -- SURVEY.md section 8f row f1: the Cell's pre / post blocks
//   ShrinkBlock  (utils/operations.py:206-218)  ReLU -> Conv2d 3x3 (c_in0 in {32, 64, 96, 128} -> 32, pad 1, no bias) -> BatchNorm2d
//   RectifyBlock (utils/operations.py:221-232)           Conv2d 3x3 (24 -> 32) -> BatchNorm2d
// as ONE library op, forward and backward, on the tcgen05 kernels of conv_tc.cuh (bf16 operands, fp32 accumulation):
// N = 32 output channels is the shape those kernels were built for (4 "terms" of 8 channels = one full N = 32 tile).
//
//   forward : x (NHWC fp32, any pixel stride -- the concat buffer is consumed in place) --(ReLU, cast)--> bf16 slices of
//             32 channels (24 -> padded with zeros) --conv_tc_fwd, one launch per slice accumulating--> y (pre-BN, fp32, NHWC
//             32) with the BatchNorm statistics out of the last launch's epilogue --finalize--> scale / shift, running
//             statistics --apply--> out = y * scale + shift.
//   backward: one sweep over (g, y) for sum g, sum g * yhat --finalize--> d gamma, d beta and the per-channel coefficients of
//             dy = A g + B y + C --pack--> bf16 dy --conv_tc_fwd mode 1 per slice--> dx (masked by x > 0 for the ReLU)
//             and conv_tc_wgrad per slice --> dW.
// Replaces cuDNN's fprop / dgrad / stream-K wgrad + two BatchNorm kernels + ReLU + threshold_backward of every such block
// (the stock blocks ran ALONE on the step's serial spine: profiles/r1_timeline_final.log).
// Maps must be a multiple of 64 pixels wide (the tcgen05 strip); the host side (senas_b200/ops.py) keeps PyTorch's own
// modules for narrower maps and for the exact fp32 mode.
#pragma once
#ifndef SENAS_EMU
#include "conv_tc.cuh"

// x [npix][x_ld] fp32 -> dst [slice][npix][32] bf16 (channels >= c_in are 0); thread = (pixel, slice, plane of 8 channels)
__global__ void __launch_bounds__(256) cbn_cast_kernel(const float *x, int64_t x_ld, int c_in, int relu, __nv_bfloat16 *dst,
                                                       int64_t npix, int nslices) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= npix * nslices * 4) return;
  const int pl = (int)(i & 3);
  const int64_t r = i >> 2, pix = r % npix;
  const int sl = (int)(r / npix), c0 = sl * 32 + pl * 8;
  __align__(16) __nv_bfloat16 v[8];
  if (c0 < c_in) {  // (c_in is a multiple of 8)
    float4 lo = ld4(x + pix * x_ld + c0), hi = ld4(x + pix * x_ld + c0 + 4);
    if (relu) {
      lo.x = fmaxf(lo.x, 0.f), lo.y = fmaxf(lo.y, 0.f), lo.z = fmaxf(lo.z, 0.f), lo.w = fmaxf(lo.w, 0.f);
      hi.x = fmaxf(hi.x, 0.f), hi.y = fmaxf(hi.y, 0.f), hi.z = fmaxf(hi.z, 0.f), hi.w = fmaxf(hi.w, 0.f);
    }
    v[0] = __float2bfloat16(lo.x), v[1] = __float2bfloat16(lo.y), v[2] = __float2bfloat16(lo.z), v[3] = __float2bfloat16(lo.w);
    v[4] = __float2bfloat16(hi.x), v[5] = __float2bfloat16(hi.y), v[6] = __float2bfloat16(hi.z), v[7] = __float2bfloat16(hi.w);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16(0.f);
  }
  *reinterpret_cast<uint4 *>(dst + ((int64_t)sl * npix + pix) * 32 + pl * 8) = *reinterpret_cast<const uint4 *>(v);
}

// BatchNorm forward finalize.  sums: [4 terms][16] = per term (sum y[8], sum y^2[8]) over the whole batch.
// stats (saved for backward): [0:32) mean, [32:64) istd, [64:96) scale = gamma * istd, [96:128) shift = beta - mean * scale
__global__ void __launch_bounds__(32) cbn_finalize_kernel(const float *sums, float M, const float *gamma, const float *beta,
                                                          float *rmean, float *rvar, int64_t *nbt, float momentum, float eps,
                                                          int training, float *stats) {
  const int c = threadIdx.x, g = c >> 3, j = c & 7;
  float mean, var;
  if (training) {
    mean = sums[g * 16 + j] / M;
    var = fmaxf(sums[g * 16 + 8 + j] / M - mean * mean, 0.f);
    if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
    if (rvar) rvar[c] = (1.f - momentum) * rvar[c] + momentum * var * (M > 1.f ? M / (M - 1.f) : 1.f);
    if (nbt && c == 0) *nbt += 1;
  } else {
    mean = rmean[c], var = rvar[c];
  }
  const float istd = rsqrtf(var + eps), sc = gamma[c] * istd;
  stats[c] = mean, stats[32 + c] = istd, stats[64 + c] = sc, stats[96 + c] = beta[c] - mean * sc;
}

// out = y * scale + shift, NHWC 32 channels; thread = (pixel, channel quad)
__global__ void __launch_bounds__(256) cbn_apply_kernel(const float *y, const float *stats, float *out, int64_t npix) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= npix * 8) return;
  const int c = (int)(i & 7) * 4;
  const float4 v = ld4(y + i * 4), sc = ld4(stats + 64 + c), sh = ld4(stats + 96 + c);
  st4(out + i * 4, make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w)));
}

// backward statistics: partials[block][64] = (sum g[32], sum g * yhat[32]) over the block's pixels (fixed order).
// block = 256 threads = 32 pixels x 8 channel quads per step.
__global__ void __launch_bounds__(256) cbn_bwd_stats_kernel(const float *g, int64_t g_ld, const float *y, const float *stats,
                                                            int64_t npix, int px_per_block, float *partials) {
  __shared__ float s_acc[32][65];
  const int q = threadIdx.x & 7, pl = threadIdx.x >> 3, c = q * 4;
  const float4 mean = ld4(stats + c), istd = ld4(stats + 32 + c);
  const int64_t p0 = (int64_t)blockIdx.x * px_per_block, p1 = min(p0 + (int64_t)px_per_block, npix);
  float s[4] = {0.f, 0.f, 0.f, 0.f}, t[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t p = p0 + pl; p < p1; p += 32) {
    const float4 gv = ld4(g + p * g_ld + c), yv = ld4(y + p * 32 + c);
    s[0] += gv.x, s[1] += gv.y, s[2] += gv.z, s[3] += gv.w;
    t[0] += gv.x * (yv.x - mean.x) * istd.x, t[1] += gv.y * (yv.y - mean.y) * istd.y;
    t[2] += gv.z * (yv.z - mean.z) * istd.z, t[3] += gv.w * (yv.w - mean.w) * istd.w;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) s_acc[pl][c + j] = s[j], s_acc[pl][32 + c + j] = t[j];
  __syncthreads();
  if (threadIdx.x < 64) {
    float r = 0.f;
    for (int k = 0; k < 32; ++k) r += s_acc[k][threadIdx.x];
    partials[(int64_t)blockIdx.x * 64 + threadIdx.x] = r;
  }
}

// sums [64] -> d gamma, d beta, coefficients [3][32] of dy = A g + B y + C
__global__ void __launch_bounds__(32) cbn_bwd_finalize_kernel(const float *sums, float M, const float *stats, int training,
                                                              float *dgamma, float *dbeta, float *coef) {
  const int c = threadIdx.x;
  const float sg = sums[c], sgy = sums[32 + c];
  const float mean = stats[c], istd = stats[32 + c], sc = stats[64 + c];
  dgamma[c] = sgy, dbeta[c] = sg;
  if (training) {  // dy = scale * (g - sg / M - yhat * sgy / M),  yhat = (y - mean) * istd
    coef[c] = sc, coef[32 + c] = -sc * istd * sgy / M, coef[64 + c] = -sc * sg / M + sc * istd * mean * sgy / M;
  } else {
    coef[c] = sc, coef[32 + c] = 0.f, coef[64 + c] = 0.f;
  }
}

// dy = A g + B y + C -> dense bf16 [npix][32]; thread = (pixel, plane of 8 channels)
__global__ void __launch_bounds__(256) cbn_pack_dy_kernel(const float *g, int64_t g_ld, const float *y, const float *coef,
                                                          __nv_bfloat16 *dst, int64_t npix) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= npix * 4) return;
  const int c = (int)(i & 3) * 8;
  const int64_t p = i >> 2;
  __align__(16) __nv_bfloat16 v[8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float4 gv = ld4(g + p * g_ld + c + 4 * h), yv = ld4(y + p * 32 + c + 4 * h);
    const float4 A = ld4(coef + c + 4 * h), B = ld4(coef + 32 + c + 4 * h), C = ld4(coef + 64 + c + 4 * h);
    v[4 * h + 0] = __float2bfloat16(fmaf(A.x, gv.x, fmaf(B.x, yv.x, C.x)));
    v[4 * h + 1] = __float2bfloat16(fmaf(A.y, gv.y, fmaf(B.y, yv.y, C.y)));
    v[4 * h + 2] = __float2bfloat16(fmaf(A.z, gv.z, fmaf(B.z, yv.z, C.z)));
    v[4 * h + 3] = __float2bfloat16(fmaf(A.w, gv.w, fmaf(B.w, yv.w, C.w)));
  }
  *reinterpret_cast<uint4 *>(dst + p * 32 + c) = *reinterpret_cast<const uint4 *>(v);
}

// ---- workspace layout (floats) ---------------------------------------------------------------------------------
// saved  : y [npix][32] | stats [128] | xb bf16 [nslices][npix][32] (= nslices * npix * 16 floats)
// scratch: tc statistics partials [4][B * ctas][16] | sums [64] | coef [96] | bwd partials [nblk][64] | dy bf16 [npix][32]
//          | wgrad partials [B * ctas][9][1024]
struct CbnGeo {
  int B, H, W, c_in, nslices, rows, chunks, ctas;
  int64_t npix, y_off, stats_off, xb_off, saved_floats;
  int64_t part_off, sums_off, coef_off, bpart_off, dy_off, wpart_off, scratch_floats;
  int bwd_px, bwd_blocks;
};
static int cbn_geo(int B, int H, int W, int c_in, CbnGeo *g) {
  if (B < 1 || H < 1 || W < 1 || tc_strip(W) == 0) return 1;
  if (c_in != 24 && c_in != 32 && c_in != 64 && c_in != 96 && c_in != 128) return 1;
  g->B = B, g->H = H, g->W = W, g->c_in = c_in, g->nslices = (c_in + 31) / 32;
  g->npix = (int64_t)B * H * W;
  g->rows = tc_rows(H, W, B, 2);
  g->chunks = (H + g->rows - 1) / g->rows;
  g->ctas = (W / tc_strip(W)) * g->chunks;
  int64_t o = 0;
  g->y_off = o, o += g->npix * 32;
  g->stats_off = o, o += 128;
  g->xb_off = o, o += (int64_t)g->nslices * g->npix * 16;
  g->saved_floats = o;
  g->bwd_px = 2048;
  g->bwd_blocks = (int)((g->npix + g->bwd_px - 1) / g->bwd_px);
  o = 0;
  g->part_off = o, o += (int64_t)4 * B * g->ctas * 16;
  g->sums_off = o, o += 64;
  g->coef_off = o, o += 96;
  g->bpart_off = o, o += (int64_t)g->bwd_blocks * 64;
  g->dy_off = o, o += g->npix * 16;
  g->wpart_off = o, o += (int64_t)B * g->ctas * 9 * 1024;
  g->scratch_floats = o;
  return 0;
}
#endif  // SENAS_EMU
