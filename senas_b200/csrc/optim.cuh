// optim.cuh -- SURVEY.md section 8 rows f4 / f3: the parts of the search step that sit between the cells.
//
// f4  `clip_grad_norm_` + `SGD.step` and `Adam.step` (experiments/search_arc.py:282-293; architect step of
//     search/senas_search.py) over FLAT fp32 buffers.  The supernet has 3 367 parameter tensors; PyTorch's foreach
//     optimizers turn them into a few hundred multi-tensor launches plus ~3.4 k tensor views of host bookkeeping.  Here
//     the parameters, their gradients and the momentum live in one arena each (senas_b200/optim.py re-homes `p.data`),
//     so the whole update is TWO launches: a fixed-order sum of squares (bit-reproducible, no atomics) and one sweep
//     that derives the clip coefficient from the partial sums and applies weight decay, momentum and the step.  The
//     learning rate is read from device memory: a scheduler changes it without re-capturing the CUDA graph.
// f3  the gamma-weighted skip mix + concat that feeds every ShrinkBlock (search/senas_search.py:96-107):
//     out[:, slot*C:(slot+1)*C] = g0 * a (+ g1 * b) written straight into the NHWC concat buffer, and its backward.
//
// Same arithmetic as the PyTorch ops they replace (torch.nn.utils.clip_grad_norm_: coef = min(1, max_norm / (norm +
// 1e-6)); torch.optim.SGD: g += wd p, buf = mom buf + g, p -= lr buf; torch.optim.Adam with L2 weight decay), checked
// against them in tests/.
#pragma once
#include "kernels.cuh"

constexpr int kOptBlocks = 296;  // 2 per SM: the partial count of the norm, independent of n (fixed summation order)

// partials[b] = sum of g[i]^2 over the float4 chunks b, b + G, b + 2G, ... (block-reduced in a fixed order)
__global__ void __launch_bounds__(256) opt_sqnorm_kernel(const float *g, int64_t n, float *partials) {
  __shared__ float s_w[8];
  const int64_t n4 = n >> 2;
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 v = ld4(g + 4 * i);
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = n4 * 4; i < n; ++i) s += g[i] * g[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_w[w];
    partials[blockIdx.x] = t;
  }
}

// clip (in place on g, like clip_grad_norm_) + SGD with momentum and L2 weight decay.  max_norm <= 0: no clipping.
__global__ void __launch_bounds__(256) opt_sgd_kernel(float *p, float *g, float *m, int64_t n, const float *lr_dev, float mom,
                                                      float wd, float max_norm, const float *partials, int npart,
                                                      float *norm_out) {
  __shared__ float s_coef;
  if (threadIdx.x == 0) {
    float coef = 1.f;
    if (max_norm > 0.f) {
      float t = 0.f;
      for (int i = 0; i < npart; ++i) t += partials[i];
      const float norm = sqrtf(t);
      coef = fminf(max_norm / (norm + 1e-6f), 1.f);
      if (blockIdx.x == 0 && norm_out) *norm_out = norm;
    }
    s_coef = coef;
  }
  __syncthreads();
  const float coef = s_coef, lr = *lr_dev;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    float4 gv = ld4(g + 4 * i), pv = ld4(p + 4 * i), mv = ld4(m + 4 * i);
    gv.x *= coef, gv.y *= coef, gv.z *= coef, gv.w *= coef;
    st4(g + 4 * i, gv);
    mv.x = mom * mv.x + (gv.x + wd * pv.x), mv.y = mom * mv.y + (gv.y + wd * pv.y);
    mv.z = mom * mv.z + (gv.z + wd * pv.z), mv.w = mom * mv.w + (gv.w + wd * pv.w);
    st4(m + 4 * i, mv);
    pv.x -= lr * mv.x, pv.y -= lr * mv.y, pv.z -= lr * mv.z, pv.w -= lr * mv.w;
    st4(p + 4 * i, pv);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = n4 * 4; i < n; ++i) {
      const float gv = g[i] * coef;
      g[i] = gv;
      m[i] = mom * m[i] + (gv + wd * p[i]);
      p[i] -= lr * m[i];
    }
}

// torch.optim.Adam (amsgrad off, L2 weight decay added to the gradient), `step` is a device scalar incremented here
// by thread 0 of block 0 AFTER every thread has read it (one block: the arch tables are a few hundred floats).
__global__ void __launch_bounds__(256) opt_adam_kernel(float *p, const float *g, float *ea, float *es, float *step, int64_t n,
                                                       const float *lr_dev, float b1, float b2, float eps, float wd) {
  const float t = *step + 1.f, lr = *lr_dev;
  const float bc1 = 1.f - powf(b1, t), bc2s = sqrtf(1.f - powf(b2, t));
  __syncthreads();
  for (int64_t i = threadIdx.x; i < n; i += 256) {
    const float gv = g[i] + wd * p[i];
    const float a = ea[i] + (gv - ea[i]) * (1.f - b1);  // exp_avg.lerp_(g, 1 - beta1)
    const float s = b2 * es[i] + (1.f - b2) * gv * gv;
    ea[i] = a, es[i] = s;
    p[i] -= (lr / bc1) * a / (sqrtf(s) / bc2s + eps);
  }
  if (threadIdx.x == 0) *step = t;
}

// ---- f3: gamma-weighted skip mix written into a channel slice of the NHWC concat buffer -------------------------
// out[pix][c0 + c] = w[0] * a[pix][c] + (b ? w[1] * b[pix][c] : 0);   w = two device floats (a softmax pair of gamma)
struct MixArgs {
  const float *a, *b, *w;
  float *out;
  int64_t npix, a_ld, b_ld, out_ld;
  int32_t C, c0;
};
__global__ void __launch_bounds__(256) mix_fwd_kernel(MixArgs q) {
  const int Q = q.C >> 2;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= q.npix * Q) return;
  const int64_t pix = i / Q;
  const int c = (int)(i - pix * Q) * 4;
  const float w0 = q.w ? q.w[0] : 1.f;  // (slot 0 of the concat: the plain copy)
  float4 v = ld4(q.a + pix * q.a_ld + c);
  v.x *= w0, v.y *= w0, v.z *= w0, v.w *= w0;
  if (q.b) {
    const float w1 = q.w[1];
    const float4 u = ld4(q.b + pix * q.b_ld + c);
    v.x = fmaf(w1, u.x, v.x), v.y = fmaf(w1, u.y, v.y), v.z = fmaf(w1, u.z, v.z), v.w = fmaf(w1, u.w, v.w);
  }
  st4(q.out + pix * q.out_ld + q.c0 + c, v);
}
// backward: da = w0 * g, db = w1 * g (g = the slice of the concat gradient), partial sums of dw0 = <g, a>, dw1 = <g, b>
struct MixBwdArgs {
  const float *a, *b, *w, *g;
  float *da, *db, *partials;  // partials [gridDim.x][2]
  int64_t npix, a_ld, b_ld, g_ld;
  int32_t C, c0;
};
__global__ void __launch_bounds__(256) mix_bwd_kernel(MixBwdArgs q) {
  __shared__ float s_w[8][2];
  const int Q = q.C >> 2;
  const float w0 = q.w[0], w1 = q.b ? q.w[1] : 0.f;
  float d0 = 0.f, d1 = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < q.npix * Q; i += (int64_t)gridDim.x * 256) {
    const int64_t pix = i / Q;
    const int c = (int)(i - pix * Q) * 4;
    const float4 g = ld4(q.g + pix * q.g_ld + q.c0 + c);
    const float4 a = ld4(q.a + pix * q.a_ld + c);
    d0 += (g.x * a.x + g.y * a.y) + (g.z * a.z + g.w * a.w);
    if (q.da) st4(q.da + pix * q.C + c, make_float4(w0 * g.x, w0 * g.y, w0 * g.z, w0 * g.w));
    if (q.b) {
      const float4 b = ld4(q.b + pix * q.b_ld + c);
      d1 += (g.x * b.x + g.y * b.y) + (g.z * b.z + g.w * b.w);
      if (q.db) st4(q.db + pix * q.C + c, make_float4(w1 * g.x, w1 * g.y, w1 * g.z, w1 * g.w));
    }
  }
  d0 = warp_sum(d0), d1 = warp_sum(d1);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5][0] = d0, s_w[threadIdx.x >> 5][1] = d1;
  __syncthreads();
  if (threadIdx.x < 2) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_w[w][threadIdx.x];
    q.partials[(int64_t)blockIdx.x * 2 + threadIdx.x] = t;
  }
}
// dw[j] = sum over blocks of partials[b][j] (fixed order), one warp
__global__ void __launch_bounds__(32) mix_bwd_final_kernel(const float *partials, int nblk, float *dw) {
  if (threadIdx.x < 2) {
    float t = 0.f;
    for (int b = 0; b < nblk; ++b) t += partials[(int64_t)b * 2 + threadIdx.x];
    dw[threadIdx.x] = t;
  }
}

// gradient of one skip tensor T: dT[pix][c] = sum_{j < 3} coef_j * g[pix][off_j + c], where the (up to 3) slices of the
// concat gradient g that T reached are: the plain copy (slot 0, coef 1), `a` of the next slot (coef w[0] of that pair)
// and `b` of its own slot (coef w[1]).  w[j] == NULL with on[j] != 0 means coefficient 1.
struct MixDxArgs {
  const float *g;
  const float *w[3];
  int32_t on[3], off[3];
  float *out;  // dense [npix][C]
  int64_t npix, g_ld;
  int32_t C;
};
__global__ void __launch_bounds__(256) mix_dx_kernel(MixDxArgs q) {
  const int Q = q.C >> 2;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= q.npix * Q) return;
  const int64_t pix = i / Q;
  const int c = (int)(i - pix * Q) * 4;
  float4 v = f4zero();
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (!q.on[j]) continue;
    const float cf = q.w[j] ? *q.w[j] : 1.f;
    const float4 u = ld4(q.g + pix * q.g_ld + q.off[j] + c);
    v.x = fmaf(cf, u.x, v.x), v.y = fmaf(cf, u.y, v.y), v.z = fmaf(cf, u.z, v.z), v.w = fmaf(cf, u.w, v.w);
  }
  st4(q.out + pix * q.C + c, v);
}

// ---- f4: dice + cross-entropy segmentation loss (SegmentationLosses('dice_ce'), utils/loss/loss.py:45-70,124-159) -------
//   loss = mean_pixels CE(logits, target) * ce_scale + 1 - mean_{c >= 1} (2 tp_c + s) / (2 tp_c + fp_c + fn_c + s + 1e-8)
//   tp_c = sum p_c [t = c],  fp_c = sum p_c [t != c],  fn_c = sum (1 - p_c) [t = c],  p = softmax(logits) over the classes
// forward: one sweep (softmax, CE term and the 3C soft counts per block, fixed-order reduction) + a one-block finalize
// that also leaves the coefficients backward needs; backward: one sweep that recomputes the softmax.  ~15 elementwise /
// reduction launches of the PyTorch expression -> 3.  logits: element offset = n * sn + c * sc + pixel * sp.
constexpr int kLossMaxC = 8, kLossBlocks = 296;
struct LossArgs {
  const float *logits;
  const int64_t *target;  // [B][HW]
  int64_t sn, sc, sp, HW;
  int32_t B, C;
  float *partials;        // forward: [kLossBlocks][3C + 1]
  const float *coef;      // backward: [3C + 1] = ktp[C], kfp[C], kfn[C], ce weight
  const float *gout;      // backward: upstream gradient (device scalar)
  float *dlogits;         // backward: same strides as logits
};
SENAS_DEVFN float loss_softmax(const LossArgs &a, int64_t i, float *p, int *t) {  // returns -log p[t] (log-sum-exp form)
  const int64_t n = i / a.HW, px = i - n * a.HW;
  const float *l = a.logits + n * a.sn + px * a.sp;
  float m = -3.4e38f;
#pragma unroll
  for (int c = 0; c < kLossMaxC; ++c)
    if (c < a.C) p[c] = l[c * a.sc], m = fmaxf(m, p[c]);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kLossMaxC; ++c)
    if (c < a.C) p[c] = expf(p[c] - m), s += p[c];
  const float inv = 1.f / s;
#pragma unroll
  for (int c = 0; c < kLossMaxC; ++c)
    if (c < a.C) p[c] *= inv;
  const int64_t tv = a.target[i];
  if (tv < 0 || tv >= a.C) {  // not a class id: the pixel matches no class (no out-of-bounds read; PyTorch would assert)
    *t = -1;
    return 0.f;
  }
  *t = (int)tv;
  return (m + logf(s)) - l[tv * a.sc];
}
__global__ void __launch_bounds__(256) loss_fwd_kernel(LossArgs a) {
  __shared__ float s_w[8][3 * kLossMaxC + 1];
  float acc[3 * kLossMaxC + 1];
#pragma unroll
  for (int j = 0; j < 3 * kLossMaxC + 1; ++j) acc[j] = 0.f;
  const int64_t total = (int64_t)a.B * a.HW;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    float p[kLossMaxC];
    int t;
    acc[3 * kLossMaxC] += loss_softmax(a, i, p, &t);
#pragma unroll
    for (int c = 0; c < kLossMaxC; ++c)
      if (c < a.C) {
        const bool hit = t == c;
        acc[c] += hit ? p[c] : 0.f, acc[kLossMaxC + c] += hit ? 0.f : p[c], acc[2 * kLossMaxC + c] += hit ? 1.f - p[c] : 0.f;
      }
  }
#pragma unroll
  for (int j = 0; j < 3 * kLossMaxC + 1; ++j) {
    const float v = warp_sum(acc[j]);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3 * kLossMaxC + 1) {
    float r = 0.f;
    for (int w = 0; w < 8; ++w) r += s_w[w][threadIdx.x];
    a.partials[(int64_t)blockIdx.x * (3 * kLossMaxC + 1) + threadIdx.x] = r;
  }
}
// out[0] = loss; coef[0:C) = ktp, [C:2C) = kfp, [2C:3C) = kfn (d loss / d tp_c ...), coef[3C] = CE weight per pixel
__global__ void __launch_bounds__(32) loss_final_kernel(const float *partials, int nblk, int C, float npix, float ce_scale,
                                                        float smooth, float *out, float *coef) {
  __shared__ float s[3 * kLossMaxC + 1];
  if (threadIdx.x < 3 * kLossMaxC + 1) {
    float r = 0.f;
    for (int b = 0; b < nblk; ++b) r += partials[(int64_t)b * (3 * kLossMaxC + 1) + threadIdx.x];
    s[threadIdx.x] = r;
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    float dsum = 0.f;
    const float k = C > 1 ? 1.f / (float)(C - 1) : 0.f;
    for (int c = 0; c < C; ++c) {
      const float tp = s[c], fp = s[kLossMaxC + c], fn = s[2 * kLossMaxC + c];
      const float N = 2.f * tp + smooth, D = 2.f * tp + fp + fn + smooth + 1e-8f;
      if (c >= 1) dsum += N / D;
      coef[c] = c >= 1 ? -k * 2.f * (D - N) / (D * D) : 0.f;
      coef[C + c] = c >= 1 ? k * N / (D * D) : 0.f;
      coef[2 * C + c] = c >= 1 ? k * N / (D * D) : 0.f;
    }
    coef[3 * C] = ce_scale / npix;
    out[0] = s[3 * kLossMaxC] * ce_scale / npix + 1.f - k * dsum;
  }
}
__global__ void __launch_bounds__(256) loss_bwd_kernel(LossArgs a) {
  const int64_t total = (int64_t)a.B * a.HW;
  const float go = a.gout[0], cew = a.coef[3 * a.C];
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    float p[kLossMaxC], gp[kLossMaxC];
    int t;
    loss_softmax(a, i, p, &t);
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < kLossMaxC; ++c)
      if (c < a.C) {
        gp[c] = t == c ? a.coef[c] - a.coef[2 * a.C + c] : a.coef[a.C + c];  // d loss / d p_c (fn = sum (1 - p) [t = c])
        dot += gp[c] * p[c];
      }
    const int64_t n = i / a.HW, px = i - n * a.HW;
    float *d = a.dlogits + n * a.sn + px * a.sp;
#pragma unroll
    for (int c = 0; c < kLossMaxC; ++c)
      if (c < a.C) d[c * a.sc] = go * (p[c] * (gp[c] - dot) + cew * (p[c] - (t == c ? 1.f : 0.f)));
  }
}
