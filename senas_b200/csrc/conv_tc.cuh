// conv_tc.cuh -- tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation).
//
// Covers the dense 3x3 / dilated 5x5 candidates of the MixedOp (utils/operations.py:89-104 of the reference) for the
// stride-1 geometries: NORM (Conv2d) and UP (ConvTranspose2d as 4 output phases on the input grid),
// c_in = 32, and the *same candidate of up to 4 edges that share an input* in one GEMM (Cell edges 0/2/5 read in0,
// 1/3/6 read in1, search/cell.py:76-90), so N = 8 * edges (24 padded to 32) instead of 8.
//
//   D[128 pixels x 32] (TMEM, fp32) += A_tap[128 pixels x 32 ch] (smem, bf16) * W_tap[32 ch x 32] (smem, bf16)
//
// * One CTA owns a 128-pixel-wide column strip of `rows_per_cta` output rows of one image and walks down the rows.
//   Input rows live in a shared-memory ring; every input row is fetched ONCE per strip by TMA (5-D tensor map over
//   the NHWC bf16 tensor viewed as (c8, W, H, plane, N): box = one row x (128 + halo) pixels x 4 planes of 8 channels,
//   out-of-image coordinates are zero-filled by the TMA unit = the convolution's zero padding).
// * In shared memory a row block is [plane][pixel][8 ch] (16 B per pixel and plane): the canonical no-swizzle K-major
//   UMMA layout with SBO = 128 B (8 pixels) and LBO = plane stride.  A tap (dy, dx) is just a different start address
//   of the A descriptor: row slot (r + dy), pixel offset dx -- no im2col, no data movement.
// * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread, tcgen05.mma cta_group::1 kind::f16,
//   M = 128, N = 32, K = 16), warps 2-5 = epilogue (tcgen05.ld 32x32b, fp32 stores of the pre-BN outputs y_k and the
//   BatchNorm statistics).  Accumulators are double-buffered in TMEM so MMA of row r+1 overlaps the epilogue of row r.
#pragma once
#ifndef SENAS_EMU
#include <cuda.h>
#include <cuda_bf16.h>

#include "kernels.cuh"

constexpr int kTcM = 128;        // pixels per MMA
constexpr int kTcN = 32;         // padded output channels (<= 4 terms x 8)
constexpr int kTcThreads = 192;  // 6 warps
constexpr int kTcMaxTerms = 4;

struct TcConvArgs {
  const float *w[kTcMaxTerms];
  float *y[kTcMaxTerms];         // [B][Ho][Wo][8] fp32
  float *partials[kTcMaxTerms];  // [B][gridDim.x][16]
  int32_t nterms;
  int32_t mode;                  // 0: forward (per-term y + statistics); 1: data gradient (one 32-channel output)
  float *out32;                  // mode 1: dx [B][H][W][out_ld], += when accumulate
  int64_t out_ld;
  int32_t accumulate;
  int32_t ws_t, ws_k, ws_n;      // weight strides: tap, GEMM-K channel, GEMM-N channel
  int32_t H, W, Ho, Wo, so;
  int32_t rows_per_cta, row_chunks;
  int32_t P, S;                  // staged pixels per row (even), ring slots
  int32_t img_mul, img_add;      // image index of the TMA source = n * img_mul + img_add (phase-major packed dy of UP)
  int32_t M;                     // pixels per MMA = strip width: 128, or 64 for maps that are a multiple of 64 wide only
  int32_t acc_y, no_stats;       // mode 0 in several launches (DOWN: one per input phase): y += , statistics on the last
  // ConvBn blocks (convbn.cuh, SURVEY row f1: ShrinkBlock / RectifyBlock 3x3 convs, 32 output channels = 4 terms):
  int32_t y_ld;                  // mode 0: pixel stride of y (0 = 8: one dense tensor per term; 32: one NHWC tensor)
  int32_t cin_valid;             // input channels that exist in this 32-channel slice (0 = 32; RectifyBlock: 24): weights of
                                 // the others are staged as 0 and mode 1 does not store them
  const float *mask;             // mode 1: dx = mask > 0 ? dx : 0 (ReLU in front of the conv), same geometry as out32
  TapTable taps;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// no-swizzle K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
      "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

constexpr uint32_t kTcTmemCols = 256;  // 2 buffers x 4 phases x 32 columns
// instruction descriptor: D = F32, A = B = BF16, both K-major, N = 32, M = 128 (cute::UMMA::InstrDescriptor)
constexpr uint32_t kTcIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);

__global__ void __launch_bounds__(kTcThreads, 1) conv_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmap, TcConvArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.y;
  const int xseg = blockIdx.x / a.row_chunks, chunk = blockIdx.x - xseg * a.row_chunks;
  const int x0 = xseg * a.M;
  const int r0 = chunk * a.rows_per_cta, r_end = min(r0 + a.rows_per_cta, a.H);
  const int T = a.taps.n, NPH = a.taps.nphase;
  // instruction descriptor: D = F32, A = B = BF16, both K-major, N = 32, M = 128 or 64
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(a.M >> 4) << 24);
  const uint32_t row_bytes = (uint32_t)a.P * 64u;
  unsigned char *rows = smem;
  unsigned char *wsm = smem + (size_t)a.S * row_bytes;  // [T][4 k-chunks][32 n][8 k] bf16
  uint64_t *bars = reinterpret_cast<uint64_t *>(wsm + (size_t)T * 2048);
  uint64_t *full = bars, *empty = bars + a.S, *acc_full = bars + 2 * a.S, *acc_empty = acc_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
  __shared__ float s_red[4][64];

  // ---- weights: fp32 global (PyTorch layout) -> bf16 canonical K-major tiles in shared memory
  for (int i = tid; i < T * 4 * kTcN; i += kTcThreads) {
    const int nn = i % kTcN, kc = (i / kTcN) & 3, t = i / (4 * kTcN);
    // forward : K = input channel (kc*8+j), N = (term, out channel);  dgrad: K = (term = kc, out channel j), N = in channel
    const int g = a.mode == 0 ? (nn >> 3) : kc;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float f = 0.f;
      const int cin = a.mode == 0 ? kc * 8 + j : nn;  // input channel of this element
      if (g < a.nterms && (a.cin_valid == 0 || cin < a.cin_valid)) {
        const int kk = a.mode == 0 ? kc * 8 + j : j, cn = a.mode == 0 ? (nn & 7) : nn;
        f = __ldg(a.w[g] + (int64_t)a.taps.widx[t] * a.ws_t + (int64_t)kk * a.ws_k + (int64_t)cn * a.ws_n);
      }
      v[j] = __float2bfloat16(f);
    }
    *reinterpret_cast<uint4 *>(wsm + (size_t)i * 16) = *reinterpret_cast<const uint4 *>(v);
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.S; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    mbar_init(&acc_full[0], 1), mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], 4), mbar_init(&acc_empty[1], 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTcTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy weight stores -> visible to the MMA unit
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int first_in = r0 + a.taps.min_dy;
  const int last_in = r_end - 1 + a.taps.max_dy;
  float st_s[kTcN], st_q[kTcN];
#pragma unroll
  for (int i = 0; i < kTcN; ++i) st_s[i] = st_q[i] = 0.f;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int idx = 0;
      for (int y = first_in; y <= last_in; ++y, ++idx) {
        const int slot = idx & (a.S - 1), use = idx / a.S;
        if (use > 0) mbar_wait(&empty[slot], (uint32_t)((use - 1) & 1));
        mbar_expect_tx(&full[slot], row_bytes);
        tma_load_5d(rows + (size_t)slot * row_bytes, &tmap, &full[slot], 0, x0 + a.taps.min_dx, y, 0, n * a.img_mul + a.img_add);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====  The whole warp runs the loop with warp-uniform values (kernel parameters, loop counters)
    // so that the descriptors live in uniform registers -- UTCHMMA takes its operands from the uniform datapath; a
    // loop under `if (lane == 0)` makes the compiler wrap every MMA in an ELECT / R2UR.BROADCAST waterfall (~100
    // cycles per MMA, seen in the SASS of the first version).  Only the instruction itself is guarded by elect.sync.
    {
      const uint32_t lbo_a = (uint32_t)a.P * 16u;
      const uint32_t hi = (128u >> 4) | (1u << 14);                                  // SBO = 128 B, version 1
      const uint32_t a_lo_base = ((smem_u32(rows) >> 4) & 0x3FFF) | ((lbo_a >> 4) << 16);
      const uint32_t b_lo_base = ((smem_u32(wsm) >> 4) & 0x3FFF) | ((512u >> 4) << 16);
      const uint32_t a_khalf = (2u * lbo_a) >> 4, row16 = row_bytes >> 4;
      const uint32_t smask = (uint32_t)a.S - 1u;
      const bool leader = elect_one();
      int waited = 0, it = 0;
      for (int r = r0; r < r_end; ++r, ++it) {
        const int need = r + a.taps.max_dy - first_in + 1;
        while (waited < need) {
          mbar_wait(&full[waited & smask], (uint32_t)((waited / a.S) & 1));
          ++waited;
        }
        const int buf = it & 1;
        if (it >= 2) mbar_wait(&acc_empty[buf], (uint32_t)(((it >> 1) - 1) & 1));
        tc_fence_after();
        for (int ph = 0; ph < NPH; ++ph) {
          const uint32_t d = tmem_base + (uint32_t)((buf * NPH + ph) * kTcN);
          uint32_t acc = 0;
          const int t1 = a.taps.pstart[ph + 1];
          for (int t = a.taps.pstart[ph]; t < t1; ++t) {
            const uint32_t trow = (uint32_t)(a.taps.dy[t] - a.taps.min_dy), tcol = (uint32_t)(a.taps.dx[t] - a.taps.min_dx);
            const uint32_t alo = a_lo_base + (((uint32_t)it + trow) & smask) * row16 + tcol;
            const uint32_t blo = b_lo_base + (uint32_t)t * 128u;
            if (leader) {
              umma_bf16(d, ((uint64_t)hi << 32) | alo, ((uint64_t)hi << 32) | blo, idesc, acc);
              umma_bf16(d, ((uint64_t)hi << 32) | (alo + a_khalf), ((uint64_t)hi << 32) | (blo + 64u), idesc, 1u);
            }
            acc = 1;
          }
        }
        if (leader) {
          tc_commit(&acc_full[buf]);        // accumulators of this row are complete when all MMAs above retire
          tc_commit(&empty[it & smask]);    // ... and input row (r + min_dy) is no longer needed by any later row
        }
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> y (fp32) + statistics =====
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    // M = 128: accumulator row m sits in TMEM lane m.  M = 64: rows 16q..16q+15 sit in lanes 0..15 of quarter q.
    const bool lane_ok = a.M == 128 || lane < 16;
    const int px = x0 + (a.M == 128 ? q * 32 : q * 16) + lane;
    int it = 0;
    for (int r = r0; r < r_end; ++r, ++it) {
      const int buf = it & 1;
      mbar_wait(&acc_full[buf], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      for (int ph = 0; ph < NPH; ++ph) {
        float v[kTcN];
        if (a.taps.pstart[ph + 1] > a.taps.pstart[ph]) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * NPH + ph) * kTcN), v);
        } else {
#pragma unroll
          for (int i = 0; i < kTcN; ++i) v[i] = 0.f;  // phase without taps (UP dil_2_conv_5): exact zeros
        }
        const int oy = r * a.so + (ph >> 1), ox = px * a.so + (ph & 1);
        if (a.mode == 1) {
          if (!lane_ok) continue;
          float *o = a.out32 + (((int64_t)n * a.Ho + oy) * a.Wo + ox) * a.out_ld;
          const int nvalid = a.cin_valid ? a.cin_valid : kTcN;
#pragma unroll
          for (int j = 0; j < kTcN; j += 4) {
            if (j >= nvalid) continue;
            float4 u = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (a.accumulate) {
              const float4 w4 = ld4(o + j);
              u.x += w4.x, u.y += w4.y, u.z += w4.z, u.w += w4.w;
            }
            if (a.mask) {
              const float4 m4 = ld4(a.mask + (((int64_t)n * a.Ho + oy) * a.Wo + ox) * a.out_ld + j);
              u.x = m4.x > 0.f ? u.x : 0.f, u.y = m4.y > 0.f ? u.y : 0.f, u.z = m4.z > 0.f ? u.z : 0.f, u.w = m4.w > 0.f ? u.w : 0.f;
            }
            st4(o + j, u);
          }
        } else if (lane_ok && px < a.W && oy < a.Ho && ox < a.Wo) {
          const int64_t pix = ((int64_t)n * a.Ho + oy) * a.Wo + ox;
#pragma unroll
          for (int g = 0; g < kTcMaxTerms; ++g) {
            if (g < a.nterms) {
              float *o = a.y[g] + pix * (a.y_ld ? a.y_ld : 8);
              if (a.acc_y) {
                const float4 lo = ld4(o), hi = ld4(o + 4);
                v[g * 8] += lo.x, v[g * 8 + 1] += lo.y, v[g * 8 + 2] += lo.z, v[g * 8 + 3] += lo.w;
                v[g * 8 + 4] += hi.x, v[g * 8 + 5] += hi.y, v[g * 8 + 6] += hi.z, v[g * 8 + 7] += hi.w;
              }
              st4(o, make_float4(v[g * 8], v[g * 8 + 1], v[g * 8 + 2], v[g * 8 + 3]));
              st4(o + 4, make_float4(v[g * 8 + 4], v[g * 8 + 5], v[g * 8 + 6], v[g * 8 + 7]));
            }
          }
#pragma unroll
          for (int i = 0; i < kTcN; ++i) st_s[i] += v[i], st_q[i] += v[i] * v[i];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  // ---- statistics: 128 epilogue threads -> per-term partial sums (fixed order)
  __syncthreads();
  if (warp >= 2) {
    const int q = warp & 3;
#pragma unroll
    for (int i = 0; i < kTcN; ++i) {
      const float s = warp_sum(st_s[i]), sq = warp_sum(st_q[i]);
      if (lane == 0) s_red[q][i] = s, s_red[q][kTcN + i] = sq;
    }
  }
  __syncthreads();
  if (tid < 64 && a.mode == 0 && !a.no_stats) {
    const float r = s_red[0][tid] + s_red[1][tid] + s_red[2][tid] + s_red[3][tid];
    const int which = tid >> 5, col = tid & 31, g = col >> 3, c = col & 7;
    if (g < a.nterms) a.partials[g][((int64_t)n * gridDim.x + blockIdx.x) * 16 + which * 8 + c] = r;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTcTmemCols));
  }
}

// fp32 NHWC (any pixel stride) -> dense bf16 NHWC, 32 channels; thread = (pixel, 8-channel plane).
// ph_h > 0 (inputs of DOWN groups, H = 2 ph_h, W = 2 ph_w): the copy is PHASE-MAJOR, [B][4][ph_h][ph_w][32] with
// dst[n][2 py + px][i][j] = src[n][2i + py][2j + px], so that a stride-2 tap is a stride-1 tap inside one phase image.
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float *src, int64_t ld, __nv_bfloat16 *dst, int64_t npix,
                                                        int ph_h, int ph_w) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= npix * 4) return;
  const int64_t dpix = i >> 2;
  const int pl = (int)(i & 3);
  int64_t pix = dpix;
  if (ph_h > 0) {
    const int64_t hw = (int64_t)ph_h * ph_w, n = dpix / (4 * hw), r = dpix - n * 4 * hw;
    const int ph = (int)(r / hw), ij = (int)(r - ph * hw), ii = ij / ph_w, jj = ij - ii * ph_w;
    pix = (n * 2 * ph_h + 2 * ii + (ph >> 1)) * (2 * ph_w) + 2 * jj + (ph & 1);
  }
  const float4 lo = ld4(src + pix * ld + pl * 8), hi = ld4(src + pix * ld + pl * 8 + 4);
  __align__(16) __nv_bfloat16 v[8] = {__float2bfloat16(lo.x), __float2bfloat16(lo.y), __float2bfloat16(lo.z), __float2bfloat16(lo.w),
                                      __float2bfloat16(hi.x), __float2bfloat16(hi.y), __float2bfloat16(hi.z), __float2bfloat16(hi.w)};
  *reinterpret_cast<uint4 *>(dst + dpix * 32 + pl * 8) = *reinterpret_cast<const uint4 *>(v);
}

// dy of up to 4 grouped terms -> one dense bf16 NHWC tensor [B][hw][32] (channels g*8.. = dy of term g, rest 0),
// dy_g = A_g*gm_g + B_g*y_g + C_g with per-(sample, channel) coefficients.  thread = (pixel, term slot)
struct PackDyArgs {
  const float *gm[kTcMaxTerms], *y[kTcMaxTerms], *coef[kTcMaxTerms];  // coef: [3][B][8]
  int64_t y_ld[kTcMaxTerms];
  int32_t nterms, hw, batch;
  int32_t up_h, up_w;  // > 0: UP edges -- gm / y live on the (2 up_h) x (2 up_w) output grid and dst is phase-major
                       // [B][4 phases][up_h][up_w][32], each phase a dense image on the input grid (hw = up_h * up_w)
  __nv_bfloat16 *dst;
};
__global__ void __launch_bounds__(256) pack_dy_kernel(PackDyArgs a) {
  const int64_t phases = a.up_h > 0 ? 4 : 1;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x, total = (int64_t)a.batch * phases * a.hw * 4;
  if (i >= total) return;
  const int64_t dpix = i >> 2;  // destination pixel
  const int g = (int)(i & 3), n = (int)(dpix / (phases * a.hw));
  int64_t pix = dpix;           // source pixel (same image geometry unless UP)
  if (a.up_h > 0) {
    const int64_t r = dpix - (int64_t)n * 4 * a.hw;
    const int ph = (int)(r / a.hw), ij = (int)(r - (int64_t)ph * a.hw), ii = ij / a.up_w, jj = ij - ii * a.up_w;
    pix = ((int64_t)n * 2 * a.up_h + 2 * ii + (ph >> 1)) * (2 * a.up_w) + 2 * jj + (ph & 1);
  }
  __align__(16) __nv_bfloat16 v[8];
  if (g < a.nterms) {
    const float4 glo = ld4(a.gm[g] + pix * 8), ghi = ld4(a.gm[g] + pix * 8 + 4);
    const float4 ylo = ld4(a.y[g] + pix * a.y_ld[g]), yhi = ld4(a.y[g] + pix * a.y_ld[g] + 4);
    const float *A = a.coef[g] + n * 8, *B = A + a.batch * 8, *C = B + a.batch * 8;
    v[0] = __float2bfloat16(A[0] * glo.x + B[0] * ylo.x + C[0]), v[1] = __float2bfloat16(A[1] * glo.y + B[1] * ylo.y + C[1]);
    v[2] = __float2bfloat16(A[2] * glo.z + B[2] * ylo.z + C[2]), v[3] = __float2bfloat16(A[3] * glo.w + B[3] * ylo.w + C[3]);
    v[4] = __float2bfloat16(A[4] * ghi.x + B[4] * yhi.x + C[4]), v[5] = __float2bfloat16(A[5] * ghi.y + B[5] * yhi.y + C[5]);
    v[6] = __float2bfloat16(A[6] * ghi.z + B[6] * yhi.z + C[6]), v[7] = __float2bfloat16(A[7] * ghi.w + B[7] * yhi.w + C[7]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16(0.f);
  }
  *reinterpret_cast<uint4 *>(a.dst + dpix * 32 + g * 8) = *reinterpret_cast<const uint4 *>(v);
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradient on tcgen05:  dW_t[m = (term, co)][ci] = sum_pixels dy[pixel][m] * x[pixel + off_t][ci]
// GEMM-K is the pixel dimension (16 pixels per MMA), both operands are MN-major views of the same
// [plane][pixel][8 ch] row blocks the forward kernel stages:  A = packed dy row (M = 64: planes 0-3 real, 4-7 ignored),
// B = x row shifted by the tap (N = 32 input channels).  One launch handles up to 13 taps whose 64 x 32 fp32
// accumulators stay in TMEM (13 x 32 = 416 columns) for the whole strip; the epilogue writes one partial per CTA.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTcWTaps = 13;
constexpr int kTcDySlots = 4;
struct TcWgradArgs {
  int32_t H, W, rows_per_cta, row_chunks, P, S;
  int32_t Wt;                // strip width in pixels (= GEMM-K per row): 128 or 64
  int32_t t_begin, t_count;  // taps [t_begin, t_begin + t_count) of the table
  int32_t dy_img_mul, dy_img_add;  // image index of the dy rows = n * dy_img_mul + dy_img_add
  int32_t x_img_mul, x_img_add;    // same for the x rows (phase-major x of DOWN groups)
  float *partials;           // [B * gridDim.x][t_count][32 ci][32 m]
  TapTable taps;
};
// D = F32, A = B = BF16, both MN-major, N = 32, M = 64
constexpr uint32_t kTcIdescW = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);

__global__ void __launch_bounds__(kTcThreads, 1) conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                                      const __grid_constant__ CUtensorMap tmap_dy,
                                                                      TcWgradArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.y;
  const int xseg = blockIdx.x / a.row_chunks, chunk = blockIdx.x - xseg * a.row_chunks;
  const int x0 = xseg * a.Wt;
  const int r0 = chunk * a.rows_per_cta, r_end = min(r0 + a.rows_per_cta, a.H);
  const uint32_t row_bytes = (uint32_t)a.P * 64u, dy_bytes = (uint32_t)a.Wt * 64u;
  unsigned char *rows = smem;                                          // x ring: S slots
  unsigned char *dyr = smem + (size_t)a.S * row_bytes;                 // dy ring: kTcDySlots slots + 1 pad slot
  uint64_t *bars = reinterpret_cast<uint64_t *>(dyr + (size_t)(kTcDySlots + 1) * dy_bytes);
  uint64_t *full = bars, *empty = bars + a.S, *dfull = bars + 2 * a.S, *dempty = dfull + kTcDySlots, *done = dempty + kTcDySlots;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.S; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    for (int s = 0; s < kTcDySlots; ++s) mbar_init(&dfull[s], 1), mbar_init(&dempty[s], 1);
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int first_in = r0 + a.taps.min_dy, last_in = r_end - 1 + a.taps.max_dy;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer: x rows (with halo) and dy rows, interleaved in consumption order
      int idx = 0, y = first_in;
      for (int r = r0, it = 0; r < r_end; ++r, ++it) {
        for (; y <= min(r + a.taps.max_dy, last_in); ++y, ++idx) {
          const int slot = idx & (a.S - 1), use = idx / a.S;
          if (use > 0) mbar_wait(&empty[slot], (uint32_t)((use - 1) & 1));
          mbar_expect_tx(&full[slot], row_bytes);
          tma_load_5d(rows + (size_t)slot * row_bytes, &tmap_x, &full[slot], 0, x0 + a.taps.min_dx, y, 0, n * a.x_img_mul + a.x_img_add);
        }
        const int ds = it & (kTcDySlots - 1), duse = it / kTcDySlots;
        if (duse > 0) mbar_wait(&dempty[ds], (uint32_t)((duse - 1) & 1));
        mbar_expect_tx(&dfull[ds], dy_bytes);
        tma_load_5d(dyr + (size_t)ds * dy_bytes, &tmap_dy, &dfull[ds], 0, x0, r, 0, n * a.dy_img_mul + a.dy_img_add);
      }
    }
  } else if (warp == 1) {
    {  // ===== MMA issuer (warp-uniform loop, elected lane issues; see the forward kernel)
      const uint32_t hi_a = (((uint32_t)a.Wt * 16u) >> 4) | (1u << 14);     // SBO = dy plane stride (Wt px * 16 B)
      const int kks = a.Wt >> 4;                                            // MMAs per row: 16 pixels each
      const uint32_t hi_b = (((uint32_t)a.P * 16u) >> 4) | (1u << 14);      // SBO = x plane stride
      const uint32_t lbo = (128u >> 4) << 16;                               // LBO = 8 pixels * 16 B
      const uint32_t a_base = ((smem_u32(dyr) >> 4) & 0x3FFF) | lbo, b_base = ((smem_u32(rows) >> 4) & 0x3FFF) | lbo;
      const uint32_t row16 = row_bytes >> 4, smask = (uint32_t)a.S - 1u;
      const bool leader = elect_one();
      int waited = 0, it = 0;
      for (int r = r0; r < r_end; ++r, ++it) {
        const int need = r + a.taps.max_dy - first_in + 1;
        while (waited < need) {
          mbar_wait(&full[waited & smask], (uint32_t)((waited / a.S) & 1));
          ++waited;
        }
        const int ds = it & (kTcDySlots - 1);
        mbar_wait(&dfull[ds], (uint32_t)((it / kTcDySlots) & 1));
        tc_fence_after();
        const uint32_t alo0 = a_base + (uint32_t)ds * (dy_bytes >> 4);
        for (int t = 0; t < a.t_count; ++t) {
          const int tt = a.t_begin + t;
          const uint32_t trow = (uint32_t)(a.taps.dy[tt] - a.taps.min_dy), tcol = (uint32_t)(a.taps.dx[tt] - a.taps.min_dx);
          const uint32_t blo0 = b_base + (((uint32_t)it + trow) & smask) * row16 + tcol;
          const uint32_t d = tmem_base + (uint32_t)t * 32u;
          if (leader) {
#pragma unroll 4
            for (int kk = 0; kk < kks; ++kk)  // 16 pixels per MMA
              umma_bf16(d, ((uint64_t)hi_a << 32) | (alo0 + (uint32_t)kk * 16u), ((uint64_t)hi_b << 32) | (blo0 + (uint32_t)kk * 16u),
                        kTcIdescW, (it > 0 || kk > 0) ? 1u : 0u);
          }
        }
        if (leader) {
          tc_commit(&dempty[ds]);
          tc_commit(&empty[it & smask]);
        }
        __syncwarp();
      }
      if (leader) tc_commit(done);
    }
  }
  // ===== epilogue (after the whole strip): warps 2..5, TMEM lane quarters; rows m = 16*q' + lane for lane < 16
  if (warp >= 2) {
    const int q = warp & 3;
    mbar_wait(done, 0);
    tc_fence_after();
    if (q < 2) {
      float *out = a.partials + ((int64_t)n * gridDim.x + blockIdx.x) * a.t_count * 1024;
      for (int t = 0; t < a.t_count; ++t) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)t * 32u, v);
        if (lane < 16) {
          const int m = q * 16 + lane;
#pragma unroll
          for (int ci = 0; ci < 32; ++ci) out[((int64_t)t * 32 + ci) * 32 + m] = v[ci];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

// dW of term g: dst_g[widx_t*ws_t + ci*ws_ci + co*ws_co] = sum_cta partials[cta][t][ci][g*8 + co]
struct TcWgradReduceArgs {
  const float *partials;
  int32_t nrows, t_begin, t_count, nterms, ws_t, ws_ci, ws_co;
  int32_t cin_valid;  // input channels that exist (0 = 32)
  float *dst[kTcMaxTerms];
  TapTable taps;
};
__global__ void __launch_bounds__(256) tc_wgrad_reduce_kernel(TcWgradReduceArgs a) {
  const int V = a.t_count * 1024, col = blockIdx.x * 32 + (threadIdx.x & 31);
  const float s = block_rows_sum(a.partials, a.nrows, V, col);
  if (threadIdx.x < 32 && col < V) {
    const int m = col & 31, ci = (col >> 5) & 31, t = col >> 10, g = m >> 3, co = m & 7;
    if (g < a.nterms && (a.cin_valid == 0 || ci < a.cin_valid))
      a.dst[g][(int64_t)a.taps.widx[a.t_begin + t] * a.ws_t + (int64_t)ci * a.ws_ci + (int64_t)co * a.ws_co] = s;
  }
}

// ---- host side ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled tc_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (PFN_tmapEncodeTiled)p;
  }
  return fn;
}

static size_t tc_smem_bytes(const TcConvArgs &a) {
  return (size_t)a.S * a.P * 64 + (size_t)a.taps.n * 2048 + (size_t)(2 * a.S + 4) * 8 + 64;
}

// xb: dense bf16 NHWC [B][H][W][32].  Returns 0 on success, 1 when the geometry cannot run here (caller falls back
// to the exact fp32 kernel of the same family, still on the GPU), 2 on a driver error.
static int launch_conv_tc(const __nv_bfloat16 *xb, int B, TcConvArgs a, void *stream) {
  if (a.img_mul == 0) a.img_mul = 1;
  PFN_tmapEncodeTiled enc = tc_encode_fn();
  if (!enc) return 2;
  const int span_x = a.taps.max_dx - a.taps.min_dx, span_y = a.taps.max_dy - a.taps.min_dy;
  a.M = tc_strip(a.W);
  if (a.M == 0) return 1;
  a.P = (a.M + span_x + 1) & ~1;
  a.S = 16;  // power of two (cheap slot arithmetic in the single-thread issue loop)
  if (span_y + 4 > a.S) return 1;
  if (a.P > 256) return 1;
  const size_t smem = tc_smem_bytes(a);
  if (smem > 220 * 1024) return 1;
  CUtensorMap tmap;
  const cuuint64_t gdim[5] = {8, (cuuint64_t)a.W, (cuuint64_t)a.H, 4, (cuuint64_t)B * a.img_mul};
  const cuuint64_t gstr[4] = {64, (cuuint64_t)a.W * 64, 16, (cuuint64_t)a.H * a.W * 64};
  const cuuint32_t box[5] = {8, (cuuint32_t)a.P, 1, 4, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<__nv_bfloat16 *>(xb), gdim, gstr, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 2;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {  // (static shared memory of the kernel counts against the same 227 KB)
    if (cudaFuncSetAttribute(conv_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 2;
    attr_smem = smem;
  }
  dim3 grid((a.W / a.M) * a.row_chunks, B);
  SENAS_LAUNCH(conv_tc_fwd_kernel, grid, dim3(kTcThreads), smem, stream, tmap, a);
  return 0;
}
static int tc_encode(CUtensorMap *tmap, const __nv_bfloat16 *p, int B, int H, int W, int box_w) {
  PFN_tmapEncodeTiled enc = tc_encode_fn();
  if (!enc) return 2;
  const cuuint64_t gdim[5] = {8, (cuuint64_t)W, (cuuint64_t)H, 4, (cuuint64_t)B};
  const cuuint64_t gstr[4] = {64, (cuuint64_t)W * 64, 16, (cuuint64_t)H * W * 64};
  const cuuint32_t box[5] = {8, (cuuint32_t)box_w, 1, 4, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return enc(tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<__nv_bfloat16 *>(p), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : 2;
}

// wgrad of one group over taps [t0, t1) of `taps`: launches of <= 13 taps each + reductions.  partials: >= B*ctas*13*1024
// floats.  dy rows come from image n * dy_mul + dy_add of dyb (NORM: 1, 0; UP: 4, phase).
static int launch_conv_tc_wgrad(const __nv_bfloat16 *xb, const __nv_bfloat16 *dyb, int B, int H, int W, const TapTable &taps,
                                int t0, int t1, int dy_mul, int dy_add, float *partials, float *const *dst, int nterms,
                                int ws_ci, int ws_co, void *stream, int x_mul = 1, int x_add = 0, int cin_valid = 0) {
  TcWgradArgs a;
  memset(&a, 0, sizeof(a));
  a.H = H, a.W = W, a.rows_per_cta = tc_rows(H, W, B, taps.max_dy - taps.min_dy);
  a.row_chunks = (H + a.rows_per_cta - 1) / a.rows_per_cta, a.taps = taps, a.partials = partials;
  a.dy_img_mul = dy_mul, a.dy_img_add = dy_add, a.x_img_mul = x_mul, a.x_img_add = x_add;
  a.Wt = tc_strip(W);
  if (a.Wt == 0) return 1;
  a.P = (a.Wt + taps.max_dx - taps.min_dx + 1) & ~1, a.S = 16;
  const size_t smem = (size_t)a.S * a.P * 64 + (size_t)(kTcDySlots + 1) * 8192 + (size_t)(2 * a.S + 2 * kTcDySlots + 2) * 8 + 64;
  CUtensorMap mx, md;
  if (tc_encode(&mx, xb, B * x_mul, H, W, a.P) || tc_encode(&md, dyb, B * dy_mul, H, W, a.Wt)) return 2;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    if (cudaFuncSetAttribute(conv_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 2;
    attr_smem = smem;
  }
  dim3 grid((W / a.Wt) * a.row_chunks, B);
  for (int tb = t0; tb < t1; tb += kTcWTaps) {
    a.t_begin = tb, a.t_count = std::min(kTcWTaps, t1 - tb);
    SENAS_TAG("conv_tc_wgrad", 2.0 * B * H * W * a.t_count * 32 * 8 * nterms, 4.0 * B * H * W * 32);
    SENAS_LAUNCH(conv_tc_wgrad_kernel, grid, dim3(kTcThreads), smem, stream, mx, md, a);
    TcWgradReduceArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.partials = partials, ra.nrows = (int)(grid.x * B), ra.t_begin = tb, ra.t_count = a.t_count, ra.nterms = nterms;
    ra.ws_t = 1, ra.ws_ci = ws_ci, ra.ws_co = ws_co, ra.taps = taps, ra.cin_valid = cin_valid;
    for (int g = 0; g < nterms; ++g) ra.dst[g] = dst[g];
    SENAS_TAG("reduce", 0, 0);
    SENAS_LAUNCH(tc_wgrad_reduce_kernel, dim3((a.t_count * 1024 + 31) / 32), dim3(256), 0, stream, ra);
  }
  return 0;
}
#endif  // SENAS_EMU
