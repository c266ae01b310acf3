// graph.cu -- host side of libsenas_b200.so: edge-graph planning, kernel sequencing and the C ABI
// declared in include/senas_b200.h.  A graph is either one MixedOp (search/cell.py:32-43 of the
// reference) or the node loop + concat of one Cell (search/cell.py:95-110).
#include "../../include/senas_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "kernels.cuh"
// strip width (= MMA M) of the tcgen05 kernels for a map of width w: 128, else 64, else 0 (not on the tensor-core path)
static int tc_strip(int w) { return w % 128 == 0 ? 128 : (w % 64 == 0 ? 64 : 0); }
// rows of a strip per CTA of the tcgen05 kernels: a CTA loads (rows + span_y) input rows, the grid runs in waves of 148
// CTAs -- pick the row count whose (waves x rows loaded per CTA) is smallest (at most 592 CTAs: one partial each).
static int tc_rows(int H, int W, int B, int span_y) {
  const int strips = W / tc_strip(W);
  int best_rows = H, best_cost = 1 << 30;
  for (int chunks = 1; chunks <= H; ++chunks) {
    const int rows = (H + chunks - 1) / chunks;
    if ((rows - 1) * chunks >= H) continue;  // same as a smaller chunk count
    const int ctas = strips * chunks * B;
    if (ctas > 592 && chunks > 1) break;
    const int cost = ((ctas + 147) / 148) * (rows + span_y + 3);
    if (cost < best_cost) best_cost = cost, best_rows = rows;
  }
  return best_rows;
}
#include "conv_tc.cuh"
#include "comm.cuh"
#include "optim.cuh"
#include "convbn.cuh"

#ifdef SENAS_EMU
static void *dev_upload(const void *h, size_t n) {
  void *p = malloc(n);
  memcpy(p, h, n);
  return p;
}
static void dev_free(void *p) { free(p); }
#else
static void *dev_upload(const void *h, size_t n) {
  void *p = nullptr;
  if (cudaMalloc(&p, n) != cudaSuccess) return nullptr;
  cudaMemcpy(p, h, n, cudaMemcpyHostToDevice);
  return p;
}
static void dev_free(void *p) { cudaFree(p); }
#endif

static thread_local std::string g_err;
#define SENAS_FAIL(...)                           \
  do {                                            \
    char buf_[512];                               \
    snprintf(buf_, sizeof(buf_), __VA_ARGS__);    \
    g_err = buf_;                                 \
    return 1;                                     \
  } while (0)

// ------------------------------------------------------------------------------------------------
// tap tables (verified against torch in tests/test_oracle_golden.py)
// ------------------------------------------------------------------------------------------------
enum { DIR_FWD = 0, DIR_DGRAD = 1 };
struct Geo {
  TapTable taps;
  int si, so;
  bool base_is_out;  // base grid = the conv's output grid (else its input grid)
};

static Geo make_geo(int k, int dil, int op, int dir) {
  Geo g;
  memset(&g, 0, sizeof(g));
  const int pad = (k / 2) * dil;
  struct Tap {
    int dy, dx, w, ph;
  };
  std::vector<Tap> v;
  auto axis = [&](int kk, int &d, int &p) {  // per-axis offset and phase of kernel index kk
    const int off = dil * (kk - k / 2);
    if (dir == DIR_FWD) {
      if (op == SENAS_OP_UP) {
        p = (pad + dil * kk) & 1;
        d = (p + pad - dil * kk) / 2;
      } else {
        p = 0, d = off;
      }
    } else {
      if (op == SENAS_OP_NORM) {
        p = 0, d = -off;
      } else if (op == SENAS_OP_DOWN) {
        p = off & 1;
        d = (p - off) / 2;
      } else {
        p = 0, d = dil * kk - pad;
      }
    }
  };
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx) {
      int dy, py, dx, px;
      axis(ky, dy, py);
      axis(kx, dx, px);
      v.push_back({dy, dx, ky * k + kx, py * 2 + px});
    }
  const bool phased = (dir == DIR_FWD && op == SENAS_OP_UP) || (dir == DIR_DGRAD && op == SENAS_OP_DOWN);
  g.taps.n = (int)v.size();
  g.taps.nphase = phased ? 4 : 1;
  int idx = 0;
  g.taps.min_dy = g.taps.min_dx = 1000, g.taps.max_dy = g.taps.max_dx = -1000;
  for (int ph = 0; ph < 4; ++ph) {
    g.taps.pstart[ph] = idx;
    for (auto &t : v)
      if (t.ph == ph) {
        g.taps.dy[idx] = (int8_t)t.dy, g.taps.dx[idx] = (int8_t)t.dx, g.taps.widx[idx] = (int8_t)t.w;
        g.taps.phase[idx] = (int8_t)ph;
        ++idx;
        if (t.dy < g.taps.min_dy) g.taps.min_dy = t.dy;
        if (t.dy > g.taps.max_dy) g.taps.max_dy = t.dy;
        if (t.dx < g.taps.min_dx) g.taps.min_dx = t.dx;
        if (t.dx > g.taps.max_dx) g.taps.max_dx = t.dx;
      }
  }
  g.taps.pstart[4] = idx;
  if (!phased)
    for (int ph = 1; ph <= 4; ++ph) g.taps.pstart[ph] = idx;
  if (dir == DIR_FWD) {
    g.si = op == SENAS_OP_DOWN ? 2 : 1, g.so = op == SENAS_OP_UP ? 2 : 1;
    g.base_is_out = op != SENAS_OP_UP;
  } else {
    g.si = op == SENAS_OP_UP ? 2 : 1, g.so = op == SENAS_OP_DOWN ? 2 : 1;
    g.base_is_out = op == SENAS_OP_DOWN;  // here "out" = the forward conv's output grid (= gathered dy grid)
  }
  return g;
}

// A stride-2 (DOWN) convolution on a PHASE-MAJOR copy of its input: tap offset `off` (input pixels) reads phase
// (off_y & 1, off_x & 1) at offset (floor(off_y / 2), floor(off_x / 2)) of that phase image, so the conv is the sum over
// the 4 input phases of ordinary stride-1 convs on the output grid.  out[ph] = single-phase table of phase ph.
static void down_phase_tables(const TapTable &fwd, TapTable out[4]) {
  for (int ph = 0; ph < 4; ++ph) {
    TapTable &t = out[ph];
    memset(&t, 0, sizeof(t));
    t.nphase = 1;
    t.min_dy = t.min_dx = 1000, t.max_dy = t.max_dx = -1000;
    for (int q = 0; q < fwd.n; ++q) {
      const int oy = fwd.dy[q], ox = fwd.dx[q];
      if ((((oy & 1) << 1) | (ox & 1)) != ph) continue;
      const int dy = (oy - (oy & 1)) / 2, dx = (ox - (ox & 1)) / 2;  // floor division
      t.dy[t.n] = (int8_t)dy, t.dx[t.n] = (int8_t)dx, t.widx[t.n] = fwd.widx[q], t.phase[t.n] = 0;
      t.min_dy = std::min(t.min_dy, dy), t.max_dy = std::max(t.max_dy, dy);
      t.min_dx = std::min(t.min_dx, dx), t.max_dx = std::max(t.max_dx, dx);
      ++t.n;
    }
    if (t.n == 0) t.min_dy = t.max_dy = t.min_dx = t.max_dx = 0;
    t.pstart[0] = 0;
    for (int q = 1; q <= 4; ++q) t.pstart[q] = t.n;
  }
}
static int down_last_phase(const TapTable t[4]) {
  int last = 0;
  for (int ph = 0; ph < 4; ++ph)
    if (t[ph].n) last = ph;
  return last;
}

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
constexpr int kLanes = 16;                        // general lanes (each owns a slice of the tmp scratch)
constexpr int kDxLanes = 2 + SENAS_MAX_NODES;     // one per state that receives a data gradient
constexpr int kAllLanes = kLanes + kDxLanes;
static int env_flag(const char *name, int dflt) {
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}
static const int g_dw_per_edge = env_flag("SENAS_DW_EDGE", 1);  // depthwise backward groups per edge (1) or per state (0)
static const int g_split_lanes = env_flag("SENAS_SPLIT_LANES", 1);  // backward: chain lanes / weight-gradient lanes (A/B)
static const int g_dw_fwd_per_edge = env_flag("SENAS_DW_FWD_EDGE", 1);  // same for the forward (measured: 102.4 -> 101.4 ms)
static inline int64_t align4(int64_t v) { return (v + 3) & ~(int64_t)3; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static const int kDwChunk = 1024;  // base pixels per block in dw_wgrad_kernel
static const int kDwRows = 4;       // output rows per block in the sliding-window depthwise kernels
static int gather_pix(int NC, int si, int base_w);
// grouped depthwise kernels: columns per tile (forward / data gradient: kDwCols per thread; weight gradient: 1 per thread) and
// rows per tile, the largest of 32/16/8/4 that still gives the 148 SMs at least four blocks each
// lane = channel kernels (dwl_*): NORM edges whose map is at least half a tile wide (narrower maps would idle warps)
// SENAS_DW_LANE bit mask: 1 / 2 = weight gradient (C = 32 / 8), 4 / 8 = forward, 16 / 32 = data gradient.  Measured per bit
// on the head cell / a 128 x 128 up cell (scripts/lane_ab.sh, profiles/README.md): the weight gradient wins (1.56 -> 1.28
// ms, 0.54 -> 0.43 ms), forward and data gradient lose 5-20 % in place although the isolated forward kernel is faster
// (scripts/ubench/dwbench.cu) -- default 3.
enum { DWL_WGRAD = 0, DWL_FWD = 2, DWL_DX = 4 };
static int g_dw_lane = env_flag("SENAS_DW_LANE", 3);
// (round 2: a variant of dw_multi_kernel that prefetches the next input row while the current one is consumed -- 150 registers,
// 3 blocks per SM -- measured SLOWER in the step: dw_fwd 8.7 -> 9.6 ms, dw_dx 9.6 -> 11.1 ms, 88.1 -> 89.8 ms; removed.)
static int dwl_tile_w(int C) { return C == 32 ? DwLane<32>::TILE_W : DwLane<8>::TILE_W; }
static bool dw_lane_ok(int C, int w, int what) {
  return ((g_dw_lane >> what) & (C == 32 ? 1 : 2)) && 2 * w >= dwl_tile_w(C);
}
static int dw_tiles_x(int C, int w, bool wgrad, bool lane = false) {
  return lane ? cdiv(w, dwl_tile_w(C)) : cdiv(w, (wgrad ? 1 : kDwCols) * (128 / (C / 4)));
}
static int dw_rows(int C, int B, int h, int w, bool wgrad, bool lane = false) {
  const int tx = dw_tiles_x(C, w, wgrad, lane);
  for (int r = 32; r > 4; r /= 2)
    if (tx * cdiv(h, r) * B >= 4 * 148) return r;
  return 4;
}
static int dw_nblk(int C, int B, int h, int w, bool wgrad, bool lane = false) {
  return dw_tiles_x(C, w, wgrad, lane) * cdiv(h, dw_rows(C, B, h, w, wgrad, lane));
}

// fused dep-sep kernels (ds_norm_kernel): tile = DsGeo<C>::TILE_W columns x rows, the largest of 32/16/8 rows that still
// gives the 148 SMs a few blocks each
// NORM dep-sep candidates through the recompute kernels (z never stored).  Measured on B200 and OFF by default: the four
// sweeps cost 185 warp-instructions per pixel and candidate against ~85 for the spill path (recompute x3, the 1x1 as a
// cross-lane reduction, dy evaluated by every channel lane): head cell 8.1 ms vs 3.0 ms, search step 117.7 vs 92.9 ms
// (profiles/README.md, round 2).  Kept behind the switch with its emulator tests as the starting point for a version with
// the pointwise halves on tensor cores.
static int g_ds_fused = env_flag("SENAS_DS_FUSED", 0);
// bf16 mode: the depthwise output z of the grouped dep-sep chains (and the gradient dz written over it) is STORED as bf16
// (statistics, the ReLU mask and every consumer see the same rounded values, so forward and backward stay consistent).
static int g_z_bf16 = env_flag("SENAS_Z_BF16", 0);
static int ds_tile_w(int C) { return C == 32 ? DsGeo<32>::TILE_W : DsGeo<8>::TILE_W; }
static int ds_rows(int C, int B, int h, int w) {
  const int tx = cdiv(w, ds_tile_w(C));
  for (int r = 32; r > 8; r /= 2)
    if (tx * cdiv(h, r) * B >= 3 * 148) return r;
  return 8;
}
static int ds_nblk(int C, int B, int h, int w) { return cdiv(w, ds_tile_w(C)) * cdiv(h, ds_rows(C, B, h, w)); }

struct TermPlan {
  int kind = 0, k = 0, dil = 1;
  bool has_y = false, owns_y = false;
  int64_t y_off = -1, z_off = -1;
  int64_t mean_off = -1, istd_off = -1, mean1_off = -1, istd1_off = -1, ysum_off = -1, se_off = -1;
  int64_t part_off = -1, part1_off = -1, psum_off = -1, psum1_off = -1;
  int nblk = 0, nblk1 = 0;
  int64_t scale_off = -1, coef_off = -1;
  bool tc = false;  // forward runs in a tcgen05 group (bf16 operands)
  bool zb = false;  // fused dep-sep (recompute path) in bf16 mode: the gradient dz is stored as bf16
  bool fused = false;  // dep-sep on a NORM edge: z is recomputed from x in every sweep and never stored (ds_norm_kernel)
  bool dwg = false; // dep-sep: depthwise half runs in the grouped row-walk kernels (all but DOWN on odd-sized maps)
};
struct TcGroup {
  int src, op, kind, k, dil, nterms;
  int edge[4], cand[4];
};
struct EdgePlan {
  int in_h, in_w;
  TermPlan t[SENAS_MAX_CAND];
};
struct NodePlan {
  int64_t bias_off, gm_off, dnode_off, bpart_off, bsum_off;
  int nterms;
  int nblk;
  bool has_consumer;
};
struct Plan {
  int B, in_h[2], in_w[2], out_h, out_w, hw;
  int64_t saved_floats = 0, scratch_floats = 0, tmp_off = 0, tmp_floats = 0;
  std::vector<EdgePlan> edges;
  std::vector<NodePlan> nodes;
  std::vector<TcGroup> tc_groups;
  int64_t xb_off[2] = {-1, -1};  // saved: dense bf16 NHWC copy of an input state (floats offset)
  std::vector<int64_t> dyb_off;  // scratch: packed bf16 dy of each NORM group (backward)
  std::vector<BnDesc *> d_bnA, d_bnB;  // per stage
  std::vector<int> n_bnA, n_bnB;
  NodeDesc *d_nodes = nullptr;
};

struct senas_graph {
  senas_graph_desc_t d;
  std::map<std::tuple<int, int, int, int, int>, Plan *> plans;
};

static int state_stage(const senas_graph_desc_t &d, int s) { return s < d.n_inputs ? 0 : s - d.n_inputs + 1; }

static int out_dim(int op, int v) { return op == SENAS_OP_NORM ? v : (op == SENAS_OP_DOWN ? (v + 1) / 2 : 2 * v); }

static int build_plan(senas_graph *g, int B, const int32_t *ih, const int32_t *iw, Plan **outp) {
  const senas_graph_desc_t &d = g->d;
  auto key = std::make_tuple(B, (int)ih[0], (int)iw[0], d.n_inputs > 1 ? (int)ih[1] : 0, d.n_inputs > 1 ? (int)iw[1] : 0);
  auto it = g->plans.find(key);
  if (it != g->plans.end()) {
    *outp = it->second;
    return 0;
  }
  if (B < 1 || B > 128) SENAS_FAIL("batch %d unsupported (1..128 per GPU)", B);
  Plan *p = new Plan();
  p->B = B;
  for (int i = 0; i < 2; ++i) p->in_h[i] = i < d.n_inputs ? ih[i] : 0, p->in_w[i] = i < d.n_inputs ? iw[i] : 0;
  p->out_h = p->out_w = -1;
  p->edges.resize(d.n_edges);
  for (int e = 0; e < d.n_edges; ++e) {  // geometry: resolve in edge order (node edges come after their node's inputs)
    const senas_edge_desc_t &ed = d.edge[e];
    int h, w;
    if (ed.src < d.n_inputs) {
      h = ih[ed.src], w = iw[ed.src];
    } else {
      if (p->out_h < 0) SENAS_FAIL("edge %d reads a node before any input edge defined the node size", e);
      h = p->out_h, w = p->out_w;
    }
    if (h < 1 || w < 1) SENAS_FAIL("edge %d: empty input %dx%d", e, h, w);
    const int oh = out_dim(ed.op_type, h), ow = out_dim(ed.op_type, w);
    if (p->out_h < 0) p->out_h = oh, p->out_w = ow;
    if (oh != p->out_h || ow != p->out_w)
      SENAS_FAIL("edge %d produces %dx%d but the nodes are %dx%d", e, oh, ow, p->out_h, p->out_w);
    p->edges[e].in_h = h, p->edges[e].in_w = w;
  }
  p->hw = p->out_h * p->out_w;
  const int HW = p->hw;
  int64_t sv = 0, sc = 0;
  auto take = [](int64_t &cur, int64_t n) {
    int64_t o = cur;
    cur = align4(cur + n);
    return o;
  };
  const int nblk_px = cdiv(HW, 128);
  int64_t tmp_need = 0;
#ifndef SENAS_EMU
  if (d.reserved & 1) {  // SENAS_FLAG_TC_BF16: group equal candidates of edges that share an input
    for (int e = 0; e < d.n_edges; ++e) {
      const senas_edge_desc_t &ed = d.edge[e];
      if (ed.src >= d.n_inputs || ed.c_in != 32) continue;
      const bool down = ed.op_type == SENAS_OP_DOWN;  // DOWN: strips on the output grid, even-sized inputs only
      if (down && (p->edges[e].in_h != 2 * p->out_h || p->edges[e].in_w != 2 * p->out_w)) continue;
      const int strip = tc_strip(down ? p->out_w : p->edges[e].in_w);
      if (strip == 0) continue;
      for (int k = 0; k < SENAS_MAX_CAND; ++k) {
        if (ed.kind[k] != SENAS_KIND_CONV && ed.kind[k] != SENAS_KIND_SE_CONV) continue;
        Geo geo = make_geo(ed.ksize[k], ed.dilation[k], ed.op_type, DIR_FWD);
        TcConvArgs probe;
        probe.taps = geo.taps;
        probe.P = (strip + geo.taps.max_dx - geo.taps.min_dx + 1) & ~1;
        probe.S = 16;
        if (geo.taps.max_dy - geo.taps.min_dy + 4 > probe.S || probe.P > 256 || tc_smem_bytes(probe) > 220 * 1024) continue;
        TcGroup *grp = nullptr;
        for (auto &g2 : p->tc_groups)
          if (g2.src == ed.src && g2.op == ed.op_type && g2.kind == ed.kind[k] && g2.k == ed.ksize[k] &&
              g2.dil == ed.dilation[k] && g2.nterms < kTcMaxTerms)
            grp = &g2;
        if (!grp) {
          p->tc_groups.push_back(TcGroup{ed.src, ed.op_type, ed.kind[k], ed.ksize[k], ed.dilation[k], 0, {0}, {0}});
          grp = &p->tc_groups.back();
        }
        grp->edge[grp->nterms] = e, grp->cand[grp->nterms] = k, grp->nterms++;
        p->edges[e].t[k].tc = true;
      }
    }
    for (auto &g2 : p->tc_groups) {
      if (p->xb_off[g2.src] < 0) p->xb_off[g2.src] = take(sv, (int64_t)B * ih[g2.src] * iw[g2.src] * 16);
      tmp_need = std::max<int64_t>(tmp_need, (int64_t)std::max(592, B * 4) * kTcWTaps * 1024);
    }
  }
#endif
  for (int e = 0; e < d.n_edges; ++e) {
    const senas_edge_desc_t &ed = d.edge[e];
    EdgePlan &ep = p->edges[e];
    const int C = ed.c_in;
    for (int k = 0; k < SENAS_MAX_CAND; ++k) {
      TermPlan &t = ep.t[k];
      t.kind = ed.kind[k], t.k = ed.ksize[k], t.dil = ed.dilation[k];
      t.mean_off = take(sv, 8), t.istd_off = take(sv, 8);
      t.scale_off = take(sc, B * 8), t.coef_off = take(sc, 3 * B * 8);
      const int T = t.k * t.k;
      switch (t.kind) {
        case SENAS_KIND_NONE:
          break;
        case SENAS_KIND_IDENTITY:
          t.has_y = true, t.owns_y = (C != 8);
          t.nblk = C == 32 ? cdiv(HW, kPwPx) : nblk_px;  // 32 channels: quad-layout 1x1 (pw_fwd_kernel<32, true>)
          if (C == 32) tmp_need = std::max<int64_t>(tmp_need, (int64_t)B * cdiv(HW, kPwPx) * 8 * C);
          if (C != 8) tmp_need = std::max<int64_t>(tmp_need, (int64_t)B * cdiv(HW, 128 * kPxTilesPerBlock) * 8 * C);
          break;
        case SENAS_KIND_AVG_POOL:
        case SENAS_KIND_UP_SAMPLE:
          t.has_y = t.owns_y = true, t.nblk = nblk_px;
          if (C == 32)  // u / du on the input grid (8 channels) + dW partials (up_sample: low, avg_pool: high resolution)
            tmp_need = std::max<int64_t>(tmp_need, (int64_t)B * ep.in_h * ep.in_w * 8 +
                                                       (int64_t)B * cdiv(ep.in_h * ep.in_w, kPwPx) * 8 * C);
          tmp_need = std::max<int64_t>(tmp_need, (int64_t)B * cdiv(std::max(HW, ep.in_h * ep.in_w), 128 * kPxTilesPerBlock) * 8 * C);
          break;
        case SENAS_KIND_CONV:
        case SENAS_KIND_SE_CONV: {
          t.has_y = t.owns_y = true;
          Geo geo = make_geo(t.k, t.dil, ed.op_type, DIR_FWD);
          const int bh = geo.base_is_out ? p->out_h : ep.in_h, bw = geo.base_is_out ? p->out_w : ep.in_w;
          t.nblk = cdiv(bh, kTileH) * cdiv(bw, kTileW * gather_pix(8, geo.si, bw));
          if (t.tc && ed.op_type == SENAS_OP_DOWN) {  // statistics come from the launch of the last non-empty phase
            TapTable pt[4];
            down_phase_tables(geo.taps, pt);
            const TapTable &lt = pt[down_last_phase(pt)];
            t.nblk = (p->out_w / tc_strip(p->out_w)) * cdiv(p->out_h, tc_rows(p->out_h, p->out_w, B, lt.max_dy - lt.min_dy));
          } else if (t.tc) {
            t.nblk = (ep.in_w / tc_strip(ep.in_w)) * cdiv(ep.in_h, tc_rows(ep.in_h, ep.in_w, B, geo.taps.max_dy - geo.taps.min_dy));
          }
          tmp_need = std::max<int64_t>(tmp_need, (int64_t)148 * 6 * T * C * 8);
          if (t.kind == SENAS_KIND_SE_CONV) t.ysum_off = take(sv, B * 8), t.se_off = take(sv, B * 17);
          break;
        }
        case SENAS_KIND_DEPSEP: {
          t.has_y = t.owns_y = true;
          Geo geo = make_geo(t.k, 1, ed.op_type, DIR_FWD);
          const int bh = geo.base_is_out ? p->out_h : ep.in_h, bw = geo.base_is_out ? p->out_w : ep.in_w;
          t.nblk1 = cdiv(bh * bw, 128 / (C / 4));
          t.dwg = ed.op_type != SENAS_OP_DOWN || (ep.in_h == 2 * p->out_h && ep.in_w == 2 * p->out_w && C == 32);
          if (ed.op_type == SENAS_OP_NORM) t.nblk1 = dw_nblk(C, B, bh, bw, false, dw_lane_ok(C, bw, DWL_FWD));  // dw(l)_multi_kernel grid
          if (ed.op_type == SENAS_OP_UP) t.nblk1 = dw_nblk(C, B, bh, bw, true);     // dw_up_multi_kernel grid (input grid)
          if (ed.op_type == SENAS_OP_DOWN && t.dwg) t.nblk1 = dw_nblk(C, B, p->out_h, p->out_w, true);  // (output grid)
          t.nblk = cdiv(HW, kPwPx);  // pw_fwd_kernel grid
          t.fused = g_ds_fused && ed.op_type == SENAS_OP_NORM;
          if (t.fused) {  // only dz lives in the buffer (backward); bf16 in bf16 mode: a gradient value, no mask depends on it
            t.nblk1 = t.nblk = ds_nblk(C, B, bh, bw);
            t.zb = (d.reserved & 1) != 0;
          } else if (g_z_bf16 && t.dwg && (d.reserved & 1) && (HW * C) % 2 == 0) {
            t.zb = true;
          }
          t.z_off = take(sv, t.zb ? ((int64_t)B * HW * C + 1) / 2 : (int64_t)B * HW * C);
          t.mean1_off = take(sv, C), t.istd1_off = take(sv, C);
          t.part1_off = take(sc, (int64_t)B * t.nblk1 * 2 * C), t.psum1_off = take(sc, (int64_t)B * 2 * C);
          const int64_t pw_tmp = t.fused ? 2 * ((int64_t)B * t.nblk * 10 * C + 16 * C)  // both candidates of an edge share a lane
                                         : (int64_t)B * cdiv(HW, 512) * 10 * C + 16 * C;
          const int64_t dw_tmp = (int64_t)B * std::max(cdiv(bh * bw, kDwChunk), cdiv(bh, 4)) * C * T;
          tmp_need = std::max<int64_t>(tmp_need, std::max(pw_tmp, dw_tmp));
          if (t.dwg)  // grouped weight gradient: one partial per block and convolution
            tmp_need = std::max<int64_t>(tmp_need, (int64_t)B * dw_nblk(C, B, ed.op_type == SENAS_OP_DOWN ? p->out_h : bh,
                                                                        ed.op_type == SENAS_OP_DOWN ? p->out_w : bw, true,
                                                                        ed.op_type == SENAS_OP_NORM && dw_lane_ok(C, bw, DWL_WGRAD)) * kDwMaxItems * C * 25);
          break;
        }
        default:
          delete p;
          SENAS_FAIL("edge %d candidate %d: unknown kind %d", e, k, t.kind);
      }
      if (t.owns_y) t.y_off = take(sv, (int64_t)B * HW * 8);
      if (t.has_y) t.part_off = take(sc, (int64_t)B * t.nblk * 16), t.psum_off = take(sc, (int64_t)B * 16);
    }
  }
  p->nodes.resize(d.n_nodes);
  for (int i = 0; i < d.n_nodes; ++i) {
    NodePlan &np = p->nodes[i];
    int nterms = 0;
    np.has_consumer = false;
    for (int e = 0; e < d.n_edges; ++e) {
      if (d.edge[e].dst == i) nterms += SENAS_MAX_CAND;
      if (d.edge[e].src == d.n_inputs + i) np.has_consumer = true;
    }
    if (nterms > kMaxTerms) {
      delete p;
      SENAS_FAIL("node %d has %d terms (max %d)", i, nterms, kMaxTerms);
    }
    np.nblk = nblk_px;  // node_bstats_kernel grid
    np.bias_off = take(sc, B * 8);
    np.gm_off = take(sc, (int64_t)B * HW * 8);
    np.dnode_off = np.has_consumer ? take(sc, (int64_t)B * HW * 8) : -1;
    np.bpart_off = take(sc, (int64_t)B * np.nblk * (1 + nterms) * 8);
    np.bsum_off = take(sc, (int64_t)B * (1 + nterms) * 8);
    np.nterms = nterms;
  }
  p->dyb_off.assign(p->tc_groups.size(), -1);
  for (size_t gi = 0; gi < p->tc_groups.size(); ++gi) {  // one packed-dy buffer per NORM group (groups run concurrently)
    const TcGroup &g2 = p->tc_groups[gi];  // UP: dy lives on the 2x grid, packed phase-major (4 images per sample)
    p->dyb_off[gi] = take(sc, (int64_t)B * ih[g2.src] * iw[g2.src] * 16 * (g2.op == SENAS_OP_UP ? 4 : 1));
  }
  tmp_need = align4(tmp_need);
  p->tmp_off = take(sc, tmp_need * kLanes);  // one slice per general lane
  p->tmp_floats = tmp_need;
  p->saved_floats = sv, p->scratch_floats = sc;

  // device descriptor tables
  const int nstage = d.n_nodes;
  p->d_bnA.assign(nstage, nullptr), p->d_bnB.assign(nstage, nullptr);
  p->n_bnA.assign(nstage, 0), p->n_bnB.assign(nstage, 0);
  for (int s = 0; s < nstage; ++s) {
    std::vector<BnDesc> A, Bv;
    for (int e = 0; e < d.n_edges; ++e) {
      const senas_edge_desc_t &ed = d.edge[e];
      if (state_stage(d, ed.src) != s) continue;
      const EdgePlan &ep = p->edges[e];
      for (int k = 0; k < SENAS_MAX_CAND; ++k) {
        const TermPlan &t = ep.t[k];
        BnDesc b;
        memset(&b, 0, sizeof(b));
        const bool ds = t.kind == SENAS_KIND_DEPSEP;
        const int bs = ds ? 7 : 1;  // slot of the final BN group
        b.C = 8, b.nblk = t.nblk, b.zero_input = t.kind == SENAS_KIND_NONE;
        b.count_per_sample = (float)HW;
        b.part_off = t.part_off, b.mean_off = t.mean_off, b.istd_off = t.istd_off, b.ysum_off = t.ysum_off;
        b.psum_off = t.psum_off;
        b.gamma = (float *)ed.param[k][bs], b.beta = (float *)ed.param[k][bs + 1];
        b.rmean = (float *)ed.param[k][bs + 2], b.rvar = (float *)ed.param[k][bs + 3];
        b.nbt = (int64_t *)ed.param[k][bs + 4];
        if (!b.gamma || !b.beta || !b.rmean || !b.rvar || !b.nbt) {
          delete p;
          SENAS_FAIL("edge %d candidate %d: missing BatchNorm parameter pointer", e, k);
        }
        if (ds) {
          Bv.push_back(b);
          BnDesc b1 = b;
          b1.C = ed.c_in, b1.nblk = t.nblk1, b1.part_off = t.part1_off, b1.mean_off = t.mean1_off;
          b1.psum_off = t.psum1_off;
          b1.istd_off = t.istd1_off, b1.ysum_off = -1;
          b1.gamma = (float *)ed.param[k][1], b1.beta = (float *)ed.param[k][2], b1.rmean = (float *)ed.param[k][3];
          b1.rvar = (float *)ed.param[k][4], b1.nbt = (int64_t *)ed.param[k][5];
          A.push_back(b1);
        } else {
          A.push_back(b);
        }
      }
    }
    p->n_bnA[s] = (int)A.size(), p->n_bnB[s] = (int)Bv.size();
    if (!A.empty()) p->d_bnA[s] = (BnDesc *)dev_upload(A.data(), A.size() * sizeof(BnDesc));
    if (!Bv.empty()) p->d_bnB[s] = (BnDesc *)dev_upload(Bv.data(), Bv.size() * sizeof(BnDesc));
  }
  std::vector<NodeDesc> nds(d.n_nodes);
  for (int i = 0; i < d.n_nodes; ++i) {
    NodeDesc &nd = nds[i];
    memset(&nd, 0, sizeof(nd));
    const NodePlan &np = p->nodes[i];
    nd.node = i, nd.bias_off = np.bias_off, nd.gm_off = np.gm_off, nd.dnode_off = np.dnode_off;
    nd.bpart_off = np.bpart_off, nd.bsum_off = np.bsum_off, nd.nblk = np.nblk, nd.hw = HW;
    for (int e = 0; e < d.n_edges; ++e) {
      const senas_edge_desc_t &ed = d.edge[e];
      if (ed.dst != i) continue;
      if (nd.nedges >= 4) {
        delete p;
        SENAS_FAIL("node %d has more than 4 incoming edges", i);
      }
      nd.edges[nd.nedges++] = e;
      for (int k = 0; k < SENAS_MAX_CAND; ++k) {
        const TermPlan &t = p->edges[e].t[k];
        TermDesc &td = nd.t[nd.nterms++];
        const int bs = t.kind == SENAS_KIND_DEPSEP ? 7 : 1;
        td.kind = t.kind, td.edge = e, td.cand = k, td.has_y = t.has_y;
        if (t.owns_y) {
          td.y = Ref{SP_SAVED, 0, t.y_off, 8};
        } else if (t.has_y) {  // identity 8->8: y is the edge input itself
          if (ed.src < d.n_inputs) td.y = Ref{SP_IN0 + ed.src, 0, 0, 0};
          else td.y = Ref{SP_OUT, 0, (int64_t)(ed.src - d.n_inputs) * 8, 0};
        } else {
          td.y = Ref{SP_NULL, 0, 0, 0};
        }
        td.mean_off = t.mean_off, td.istd_off = t.istd_off, td.ysum_off = t.ysum_off, td.se_off = t.se_off;
        td.scale_off = t.scale_off, td.coef_off = t.coef_off;
        td.gamma = (float *)ed.param[k][bs], td.beta = (float *)ed.param[k][bs + 1];
        td.w1 = (float *)ed.param[k][6], td.w2 = (float *)ed.param[k][7];
        td.g_gamma = ed.grad_off[k][bs], td.g_beta = ed.grad_off[k][bs + 1];
        td.g_w1 = ed.grad_off[k][6], td.g_w2 = ed.grad_off[k][7];
        td.hw = (float)HW;
        if (t.kind == SENAS_KIND_SE_CONV && (!td.w1 || !td.w2)) {
          delete p;
          SENAS_FAIL("edge %d candidate %d: missing SE weights", e, k);
        }
      }
    }
  }
  p->d_nodes = (NodeDesc *)dev_upload(nds.data(), nds.size() * sizeof(NodeDesc));
  g->plans[key] = p;
  *outp = p;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
static const int kPersistBlocks = 148 * 6;

// ------------------------------------------------------------------------------------------------
// lanes: independent candidate chains of a call are spread over side streams (fork / join with events), so that a
// captured search step is a DAG whose small kernels overlap instead of a 9 000-node chain (the cells at <= 64 x 64
// are latency-bound: ~300 dependent launches of a few microseconds each).  Accumulations into a shared gradient
// (dx of an input state) stay on ONE lane per state in issue order, so results are bit-identical to the serial order.
// ------------------------------------------------------------------------------------------------
static int g_lanes = -1;                          // -1: read SENAS_LANES on first use (default kLanes); 0: serial

#ifndef SENAS_EMU
constexpr int kEventPool = 2048;
struct LaneSet {
  cudaStream_t s[kAllLanes];
  cudaEvent_t ev[kEventPool];
  int next_ev = 0;
  int pending = 0;  // general lanes carry weight-gradient work of an earlier backward call that nobody has joined yet
};
static std::vector<LaneSet *> g_lane_sets;
static int g_defer = 0;  // senas_set_defer: leave the weight-gradient lanes of a backward call running (joined later)
// one lane set per (device, slot): cells that the host runs concurrently on different streams (independent cells of one
// level of the UNet++ triangle, senas_b200/supernet.py) select different slots (senas_set_slot) so that they do not
// queue behind each other on shared lanes.  Slots, not stream handles, so that warm-up and capture share the sets.
static int g_slot = 0;
static LaneSet *lanes_for(int slot) {
  static std::map<std::pair<int, int>, LaneSet *> sets;
  int dev = 0;
  cudaGetDevice(&dev);
  const auto key = std::make_pair(dev, slot);
  auto it = sets.find(key);
  if (it != sets.end()) return it->second;
  LaneSet *ls = new LaneSet();
  // backward: lanes [0, n/2) carry the chains that feed a data gradient (statistics -> dz, packed dy) and the dx lanes
  // carry the data gradients themselves: both are on the critical path of the next node / the caller and get the high
  // stream priority; lanes [n/2, n) carry weight gradients that nothing waits for (low priority).  Priorities are
  // captured into the graph's kernel nodes.
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  for (int i = 0; i < kAllLanes; ++i)
    cudaStreamCreateWithPriority(&ls->s[i], cudaStreamNonBlocking, (i >= kLanes / 2 && i < kLanes) ? prio_lo : prio_hi);
  for (int i = 0; i < kEventPool; ++i) cudaEventCreateWithFlags(&ls->ev[i], cudaEventDisableTiming);
  sets[key] = ls;
  g_lane_sets.push_back(ls);
  return ls;
}
#endif

struct Sched {
  void *main = nullptr;
  int n = 0;  // 0 = serial: every lane is the main stream
  int rr = 0;
#ifndef SENAS_EMU
  LaneSet *ls = nullptr;
#endif
  void init(void *main_stream) {
    main = main_stream;
#ifndef SENAS_EMU
    if (g_lanes < 0) {
      const char *e = getenv("SENAS_LANES");
      g_lanes = e ? std::max(0, std::min(kLanes, atoi(e))) : kLanes;
    }
    n = g_prof_on ? 0 : g_lanes;  // per-kernel timing wants serial launches
    if (n > 0) {
      ls = lanes_for(g_slot);
      if (ls->pending) {  // deferred weight-gradient work of the previous backward call in this slot: order it before us
        for (int i = 0; i < kLanes; ++i) dep(i, -1);
        ls->pending = 0;
      }
    }
#endif
  }
  // end of a backward call: the caller needs the data gradients (dx lanes) now; the general lanes only finish weight
  // gradients that nothing reads before the optimizer.  With senas_set_defer(1) they keep running beside whatever the
  // caller enqueues next and are joined by the next call of the slot or by senas_flush(); the host keeps the buffers
  // they read alive until then (senas_b200/fused.py).
  void join_backward() {
#ifndef SENAS_EMU
    if (n == 0) return;
    if (!g_defer) {
      join();
      return;
    }
    for (int i = kLanes; i < kAllLanes; ++i) dep(i, -1);
    ls->pending = 1;
#endif
  }
  void *stream(int lane) const {
#ifndef SENAS_EMU
    if (n > 0 && lane >= 0) return (void *)ls->s[lane];
#endif
    (void)lane;
    return main;
  }
  // next general lane (round robin); its index doubles as the tmp-scratch slot
  int pick() {
    if (n == 0) return -1;
    const int l = rr;
    rr = (rr + 1) % n;
    return l;
  }
  // backward: chain lanes (first half) / background lanes for weight gradients (second half)
  int rr_chain = 0, rr_bg = 0;
  int pick_chain() {
    if (n == 0) return -1;
    if (n < 2 || !g_split_lanes) return pick();
    const int l = rr_chain;
    rr_chain = (rr_chain + 1) % (n / 2);
    return l;
  }
  int pick_bg() {
    if (n == 0) return -1;
    if (n < 2 || !g_split_lanes) return pick();
    const int l = n / 2 + rr_bg;
    rr_bg = (rr_bg + 1) % (n - n / 2);
    return l;
  }
  int dx_lane(int state) const { return n == 0 ? -1 : kLanes + state; }
  int tmp_slot(int lane) const { return lane < 0 ? 0 : lane; }
  void dep(int from, int to) {  // work enqueued on `to` from now on waits for everything enqueued on `from` so far
#ifndef SENAS_EMU
    if (n == 0 || from == to) return;
    cudaEvent_t e = ls->ev[ls->next_ev];
    ls->next_ev = (ls->next_ev + 1) % kEventPool;
    cudaEventRecord(e, (cudaStream_t)stream(from));
    cudaStreamWaitEvent((cudaStream_t)stream(to), e, 0);
#else
    (void)from, (void)to;
#endif
  }
  void fork() {  // every lane waits for the main stream's tail
#ifndef SENAS_EMU
    if (n == 0) return;
    cudaEvent_t e = ls->ev[ls->next_ev];
    ls->next_ev = (ls->next_ev + 1) % kEventPool;
    cudaEventRecord(e, (cudaStream_t)main);
    for (int i = 0; i < n; ++i) cudaStreamWaitEvent(ls->s[i], e, 0);
    for (int i = kLanes; i < kAllLanes; ++i) cudaStreamWaitEvent(ls->s[i], e, 0);
#endif
  }
  void join() {  // the main stream waits for every lane
#ifndef SENAS_EMU
    if (n == 0) return;
    for (int i = 0; i < kAllLanes; ++i) {
      if (i >= n && i < kLanes) continue;
      dep(i, -1);
    }
#endif
  }
};

// Opt a kernel in to more than the default 48 KB of shared memory.  The attribute is per FUNCTION and only ever raised:
// a launch recorded in a CUDA graph is replayed with whatever value the function has at replay time, so a later, smaller
// launch of the same kernel must not lower it (found in round 2: replayed launches of a captured step failed silently).
// Static shared memory counts against the default as well: everything above 46 KB opts in.
template <typename K>
static void allow_smem(K kern, size_t bytes) {
#ifndef SENAS_EMU
  static std::map<const void *, size_t> granted;
  if (bytes <= 46 * 1024) return;
  size_t &cur = granted[(const void *)kern];
  if (bytes > cur) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    cur = bytes;
  }
#else
  (void)kern, (void)bytes;
#endif
}

// base pixels per thread of gather_mac_kernel (tile = 8 rows x 16*PIX columns): as many as the map width and the
// shared-memory footprint of the gathered tile allow
static int gather_pix(int NC, int si, int base_w) {
  const int want = (NC == 8 && si == 1) ? 4 : 2;
  int pix = 1;
  while (pix < want && base_w >= 32 * pix) pix *= 2;
  return pix;
}
static size_t gather_smem(const Geo &g, int NC, int pix) {
  const int R = (kTileH - 1) * g.si + (g.taps.max_dy - g.taps.min_dy) + 1;
  const int Cc = (kTileW * pix - 1) * g.si + (g.taps.max_dx - g.taps.min_dx) + 1;
  return (size_t)2 * R * Cc * 16 + (size_t)g.taps.n * 8 * NC * 4;
}

template <int KC, int NC, int NPH, int PIX>
static void launch_gather(GatherArgs &a, const Geo &g, int B, void *stream, bool mma) {
  // mma: bf16 mode -- the same tile and staging with the inner product on tensor cores (gather_mma_kernel, TF32 mma.sync)
  const size_t smem = gather_smem(g, mma ? GatherMma<NC>::NCP : NC, PIX);
  a.tiles_x = cdiv(a.base_w, kTileW * PIX);
  dim3 grid(a.tiles_x * cdiv(a.base_h, kTileH), B);
  if (mma) {
    auto kern = gather_mma_kernel<KC, NC, NPH, PIX>;
    allow_smem(kern, smem);
    SENAS_LAUNCH(kern, grid, dim3(kTileThreads), smem, stream, a);
  } else {
    auto kern = gather_mac_kernel<KC, NC, NPH, PIX>;
    allow_smem(kern, smem);
    SENAS_LAUNCH(kern, grid, dim3(kTileThreads), smem, stream, a);
  }
}

// bf16 mode: the convolutions that are not on the tcgen05 path through mma.sync TF32 (gather_mma_kernel).  OFF by default:
// measured 38 vs 29 TFLOP/s on the 8 -> 8 edges of the head cell but only 94.1 -> 92.1 ms per search step (the kernel is
// bound by its serial stage -> sync -> MMA -> store structure, not by the FMA pipe), and the extra TF32 rounding on the node
// edges flipped a near-tie of the fixed-seed genotype gate.  Opt in with SENAS_GATHER_MMA=1 / senas_set_gather_mma(1).
static int g_gather_mma = env_flag("SENAS_GATHER_MMA", 0);
// bf16 mode: weight gradient of the 8 -> 8 node edges as mma.sync TF32 (conv_wgrad_mma8_kernel) instead of the fp32
// CUDA-core kernel.  A parameter gradient (a sum over all pixels): the TF32 rounding of its operands averages out and no
// forward value or ReLU decision depends on it.
static int g_wgrad_mma = env_flag("SENAS_WGRAD_MMA", 1);
// bf16 mode: data gradients of the non-tcgen05 convs through gather_mma_kernel.  Measured (r2k): the dgrad families drop
// from 12.0 to 9.9 ms of kernel time per step, the step does not move (88.1 vs 87.5 ms, inside the run-to-run noise) -- OFF.
static int g_dgrad_mma = env_flag("SENAS_DGRAD_MMA", 0);
static int launch_gather_any(GatherArgs &a, const Geo &g, int KC, int NC, int B, void *stream, bool mma = false) {
  const int nph = g.taps.nphase, pix = gather_pix(NC, g.si, a.base_w);
  // forward (a.partials != nullptr: statistics epilogue) only with the opt-in; data gradients (no forward value, no ReLU
  // decision depends on them: the TF32 rounding stays inside the 2e-2 gate of the mode) also with SENAS_DGRAD_MMA
  mma = mma && (g_gather_mma || (g_dgrad_mma && a.partials == nullptr));
#define SENAS_GATHER(KC_, NC_, NPH_)                                                       \
  if (KC == KC_ && NC == NC_ && nph == NPH_) {                                             \
    if (pix == 4 && NC_ == 8) launch_gather<KC_, NC_, NPH_, (NC_ == 8 ? 4 : 2)>(a, g, B, stream, mma); \
    else if (pix >= 2) launch_gather<KC_, NC_, NPH_, 2>(a, g, B, stream, mma);                  \
    else launch_gather<KC_, NC_, NPH_, 1>(a, g, B, stream, mma);                                \
    return 0;                                                                              \
  }
  SENAS_GATHER(32, 8, 1)
  SENAS_GATHER(32, 8, 4)
  SENAS_GATHER(8, 8, 1)
  SENAS_GATHER(8, 8, 4)
  SENAS_GATHER(8, 32, 1)
  SENAS_GATHER(8, 32, 4)
#undef SENAS_GATHER
  SENAS_FAIL("no gather kernel for KC=%d NC=%d phases=%d", KC, NC, nph);
}

struct Call {  // per-call resolved pointers
  const senas_graph_desc_t *d;
  Plan *p;
  Bases bases;
  float *saved, *scratch;
  const float *in[2];
  int64_t in_ld[2];
  float *out;
  int64_t out_ld;
  void *stream;
  int B;
  Sched S;
  float *tmp(int lane) const { return scratch + p->tmp_off + (int64_t)S.tmp_slot(lane) * p->tmp_floats; }
};
static const float *state_ptr(const Call &c, int s, int64_t *ld) {
  if (s < c.d->n_inputs) {
    *ld = c.in_ld[s];
    return c.in[s];
  }
  *ld = c.out_ld;
  return c.out + (s - c.d->n_inputs) * 8;
}

static void conv_weight_strides(int op, int c_in, int T, int dir, int *ws_k, int *ws_n) {
  // forward: k = ci, n = co;  dgrad: k = co, n = ci
  const int s_ci = op == SENAS_OP_UP ? 8 * T : T, s_co = op == SENAS_OP_UP ? T : c_in * T;
  if (dir == DIR_FWD) *ws_k = s_ci, *ws_n = s_co;
  else *ws_k = s_co, *ws_n = s_ci;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// dep-sep candidates of NORM edges: the depthwise halves of every edge that reads `src` go out as ONE launch
// (dw_multi_kernel: the input tile is read from HBM once for up to 6 convolutions).
static int forward_dw_group(Call &c, int src, int only_edge) {
  const senas_graph_desc_t &d = *c.d;
  const Plan &p = *c.p;
  DwMultiArgs a;
  memset(&a, 0, sizeof(a));
  int C = 0, h = 0, w = 0, nblk = 0;
  bool up = false, down = false;  // every edge that reads a state has the same op type (cell.py:76-90)
  int64_t x_ld = 0;
  const float *x = state_ptr(c, src, &x_ld);
  for (int e = 0; e < d.n_edges; ++e) {
    const senas_edge_desc_t &ed = d.edge[e];
    if (ed.src != src || (only_edge >= 0 && e != only_edge)) continue;
    up = ed.op_type == SENAS_OP_UP, down = ed.op_type == SENAS_OP_DOWN;
    for (int k = 0; k < SENAS_MAX_CAND; ++k) {
      const TermPlan &t = p.edges[e].t[k];
      if (t.kind != SENAS_KIND_DEPSEP || !t.dwg || t.fused) continue;
      if (a.n == kDwMaxItems) SENAS_FAIL("more than %d dep-sep candidates read state %d", kDwMaxItems, src);
      DwItem &it = a.it[a.n++];
      it.in = x, it.in_ld = x_ld, it.out = c.saved + t.z_off, it.out_ld = ed.c_in, it.out_bf = t.zb;
      it.w = (const float *)ed.param[k][0], it.partials = c.scratch + t.part1_off, it.k = t.k;
      C = ed.c_in, h = down ? p.out_h : p.edges[e].in_h, w = down ? p.out_w : p.edges[e].in_w, nblk = t.nblk1;
    }
  }
  if (a.n == 0) return 0;
  const bool col1 = up || down;  // one column per thread, tiles on the low-resolution grid
  const bool lane = !up && !down && dw_lane_ok(C, w, DWL_FWD);
  a.H = h, a.W = w, a.tiles_x = dw_tiles_x(C, w, col1, lane), a.tile_rows = dw_rows(C, c.B, h, w, col1, lane);
  double taps = 0;
  for (int m = 0; m < a.n; ++m) taps += a.it[m].k * a.it[m].k;
  void *st = c.S.stream(c.S.pick());
  SENAS_TAG("dw_fwd", 2.0 * c.B * h * w * taps * C, 4.0 * c.B * h * w * C * (1 + a.n * (up ? 4 : 1)));
  dim3 grid(nblk, c.B);
  if (up) {
    if (C != 32) SENAS_FAIL("UP depthwise: c_in %d unsupported", C);
    auto kern = dw_up_multi_kernel<32, true, true>;
    SENAS_LAUNCH(kern, grid, dim3(128), 0, st, a);
  } else if (down) {  // z[o] = sum x[2o + k - P] w[k]: the stride-2 gather, with statistics
    auto kern = dw_up_multi_kernel<32, false, true>;
    SENAS_LAUNCH(kern, grid, dim3(128), 0, st, a);
  } else if (lane && C == 32) {
    auto kern = dwl_multi_kernel<32, true>;
    SENAS_LAUNCH(kern, grid, dim3(128), 0, st, a);
  } else if (lane) {
    auto kern = dwl_multi_kernel<8, true>;
    SENAS_LAUNCH(kern, grid, dim3(128), 0, st, a);
  } else if (C == 32) {
    auto kern = dw_multi_kernel<32, true>;
    SENAS_LAUNCH(kern, grid, dim3(128), 0, st, a);
  } else {
    auto kern = dw_multi_kernel<8, true>;
    SENAS_LAUNCH(kern, grid, dim3(128), 0, st, a);
  }
  return 0;
}

// fused dep-sep candidates of the NORM edges that read `src`: ONE launch per sweep for all of them (the input tile is
// read once for up to 6 convolutions).  mode = DS_FWD_STATS (stage A) or DS_FWD_Y (stage B).
static int forward_ds_group(Call &c, int src, int mode) {
  const senas_graph_desc_t &d = *c.d;
  const Plan &p = *c.p;
  DsArgs a;
  memset(&a, 0, sizeof(a));
  int C = 0, h = 0, w = 0, nblk = 0;
  a.x = state_ptr(c, src, &a.x_ld);
  double taps = 0;
  for (int e = 0; e < d.n_edges; ++e) {
    const senas_edge_desc_t &ed = d.edge[e];
    if (ed.src != src) continue;
    for (int k = 0; k < SENAS_MAX_CAND; ++k) {
      const TermPlan &t = p.edges[e].t[k];
      if (t.kind != SENAS_KIND_DEPSEP || !t.fused) continue;
      if (a.n == kDsMaxItems) SENAS_FAIL("more than %d fused dep-sep candidates read state %d", kDsMaxItems, src);
      DsItem &it = a.it[a.n++];
      it.w = (const float *)ed.param[k][0], it.k = t.k;
      it.mean1 = c.saved + t.mean1_off, it.istd1 = c.saved + t.istd1_off;
      it.g1 = (const float *)ed.param[k][1], it.b1 = (const float *)ed.param[k][2], it.wpw = (const float *)ed.param[k][6];
      it.y = c.saved + t.y_off;
      it.partials = c.scratch + (mode == DS_FWD_STATS ? t.part1_off : t.part_off);
      C = ed.c_in, h = p.edges[e].in_h, w = p.edges[e].in_w, nblk = t.nblk;
      taps += t.k * t.k;
    }
  }
  if (a.n == 0) return 0;
  a.H = h, a.W = w, a.tiles_x = cdiv(w, ds_tile_w(C)), a.tile_rows = ds_rows(C, c.B, h, w), a.batch = c.B;
  void *st = c.S.stream(c.S.pick());
  dim3 grid(nblk, c.B);
  const double px = (double)c.B * h * w;
  if (mode == DS_FWD_STATS) {
    SENAS_TAG("ds_fwd_stats", 2.0 * px * taps * C, 4.0 * px * C);
    if (C == 32) {
      auto kern = ds_norm_kernel<32, DS_FWD_STATS>;
      SENAS_LAUNCH(kern, grid, dim3(128), 0, st, a);
    }
    else {
      auto kern = ds_norm_kernel<8, DS_FWD_STATS>;
      SENAS_LAUNCH(kern, grid, dim3(128), 0, st, a);
    }
  } else {
    SENAS_TAG("ds_fwd_y", 2.0 * px * (taps * C + 8.0 * C * a.n), 4.0 * px * (C + 8.0 * a.n));
    if (C == 32) {
      auto kern = ds_norm_kernel<32, DS_FWD_Y>;
      SENAS_LAUNCH(kern, grid, dim3(128), 0, st, a);
    }
    else {
      auto kern = ds_norm_kernel<8, DS_FWD_Y>;
      SENAS_LAUNCH(kern, grid, dim3(128), 0, st, a);
    }
  }
  return 0;
}

static int forward_edge(Call &c, int e, bool second_pass) {
  const senas_edge_desc_t &ed = c.d->edge[e];
  const EdgePlan &ep = c.p->edges[e];
  const Plan &p = *c.p;
  const int B = c.B, C = ed.c_in;
  int64_t x_ld;
  const float *x = state_ptr(c, ed.src, &x_ld);
  for (int k = 0; k < SENAS_MAX_CAND; ++k) {
    const TermPlan &t = ep.t[k];
    float *y = t.owns_y ? c.saved + t.y_off : nullptr;
    float *part = t.has_y ? c.scratch + t.part_off : nullptr;
    if (t.kind == SENAS_KIND_NONE || (second_pass && t.kind != SENAS_KIND_DEPSEP) ||
        (!second_pass && t.tc && (t.kind == SENAS_KIND_CONV || t.kind == SENAS_KIND_SE_CONV)))
      continue;
    if (t.kind == SENAS_KIND_DEPSEP && t.fused) continue;  // forward_ds_group
    const int ln = c.S.pick();  // candidates of a stage are independent: one lane each (with its tmp slice)
    void *st = c.S.stream(ln);
    if (second_pass) {
      PwArgs a;
      a.z = c.saved + t.z_off, a.z_ld = C, a.hw = p.hw, a.y = y;
      a.mean1 = c.saved + t.mean1_off, a.istd1 = c.saved + t.istd1_off;
      a.g1 = (const float *)ed.param[k][1], a.b1 = (const float *)ed.param[k][2];
      a.wpw = (const float *)ed.param[k][6], a.partials = part, a.z_bf = t.zb;
      dim3 grid(cdiv(p.hw, kPwPx), B);
      SENAS_TAG("pw_fwd", 2.0 * B * p.hw * C * 8, 4.0 * B * p.hw * (C + 8));
      if (C == 32) {
        auto kern = pw_fwd_kernel<32, false>;
        SENAS_LAUNCH(kern, grid, dim3(256), 0, st, a);
      } else {
        auto kern = pw_fwd_kernel<8, false>;
        SENAS_LAUNCH(kern, grid, dim3(256), 0, st, a);
      }
      continue;
    }
    switch (t.kind) {
      case SENAS_KIND_NONE:
        break;
      case SENAS_KIND_IDENTITY:
      case SENAS_KIND_AVG_POOL:
      case SENAS_KIND_UP_SAMPLE: {
        AdapterArgs a;
        a.x = x, a.x_ld = x_ld, a.x_h = ep.in_h, a.x_w = ep.in_w, a.y = y, a.o_h = p.out_h, a.o_w = p.out_w;
        a.w = (const float *)ed.param[k][0], a.partials = part;
        if ((C != 8) != (a.w != nullptr)) SENAS_FAIL("edge %d candidate %d: 1x1 weight does not match c_in", e, k);
        if (C == 32) {  // quad-layout 1x1 on the input grid (+ 8-channel upsample / pooling)
          const bool up = t.kind == SENAS_KIND_UP_SAMPLE, pool = t.kind == SENAS_KIND_AVG_POOL;
          const int in_px = ep.in_h * ep.in_w;
          PwArgs pa;
          memset(&pa, 0, sizeof(pa));
          pa.z = x, pa.z_ld = x_ld, pa.hw = in_px, pa.wpw = a.w;
          pa.y = (up || pool) ? c.tmp(ln) : y, pa.partials = (up || pool) ? nullptr : part;
          SENAS_TAG("adapter_fwd", 2.0 * B * in_px * C * 8, 4.0 * B * (in_px * C + p.hw * 8));
          auto kern = pw_fwd_kernel<32, true>;
          SENAS_LAUNCH(kern, dim3(cdiv(in_px, kPwPx), B), dim3(256), 0, st, pa);
          if (up) {
            Up8Args ua;
            ua.u = c.tmp(ln), ua.y = y, ua.h = ep.in_h, ua.w = ep.in_w, ua.partials = part;
            SENAS_TAG("adapter_fwd", 0, 0);
            SENAS_LAUNCH(up8_fwd_kernel, dim3(cdiv(p.hw, 128), B), dim3(128), 0, st, ua);
          }
          if (pool) {
            Pool8Args qa;
            qa.u = c.tmp(ln), qa.y = y, qa.h = ep.in_h, qa.w = ep.in_w, qa.oh = p.out_h, qa.ow = p.out_w, qa.partials = part;
            SENAS_TAG("adapter_fwd", 0, 0);
            SENAS_LAUNCH(pool8_fwd_kernel, dim3(cdiv(p.hw, 128), B), dim3(128), 0, st, qa);
          }
          break;
        }
        dim3 grid(cdiv(p.hw, 128), B);
#define SENAS_AD_FWD(CC, KK)                                   \
  {                                                            \
    auto kern = adapter_fwd_kernel<CC, KK>;                    \
    SENAS_TAG("adapter_fwd", 2.0 * B * p.hw * CC * 8, 4.0 * B * (ep.in_h * ep.in_w * CC + p.hw * 8)); \
    SENAS_LAUNCH(kern, grid, dim3(128), 0, st, a);       \
  }
        const int kk = t.kind == SENAS_KIND_IDENTITY ? AD_IDENTITY : (t.kind == SENAS_KIND_AVG_POOL ? AD_POOL : AD_UP);
        if (C == 32 && kk == AD_IDENTITY) SENAS_AD_FWD(32, AD_IDENTITY)
        else if (C == 32 && kk == AD_POOL) SENAS_AD_FWD(32, AD_POOL)
        else if (C == 32 && kk == AD_UP) SENAS_AD_FWD(32, AD_UP)
        else if (C == 8 && kk == AD_IDENTITY) SENAS_AD_FWD(8, AD_IDENTITY)
        else SENAS_FAIL("edge %d candidate %d: adapter kind %d with c_in %d unsupported", e, k, t.kind, C);
        break;
      }
      case SENAS_KIND_CONV:
      case SENAS_KIND_SE_CONV: {
        if (t.tc) break;  // done by its tcgen05 group
        Geo geo = make_geo(t.k, t.dil, ed.op_type, DIR_FWD);
        GatherArgs a;
        memset(&a, 0, sizeof(a));
        a.src = x, a.src_ld = x_ld, a.src_h = ep.in_h, a.src_w = ep.in_w;
        a.dst = y, a.dst_ld = 8, a.dst_h = p.out_h, a.dst_w = p.out_w;
        a.base_h = geo.base_is_out ? p.out_h : ep.in_h, a.base_w = geo.base_is_out ? p.out_w : ep.in_w;
        a.si = geo.si, a.so = geo.so;
        a.w = (const float *)ed.param[k][0], a.ws_t = 1;
        conv_weight_strides(ed.op_type, C, t.k * t.k, DIR_FWD, &a.ws_k, &a.ws_n);
        a.partials = part, a.taps = geo.taps;
        SENAS_TAG(ed.op_type == SENAS_OP_DOWN ? "conv_fwd.down" : (C == 8 ? "conv_fwd.n8" : (ed.op_type == SENAS_OP_UP ? "conv_fwd.up_small" : "conv_fwd.n32_small")),
                  2.0 * B * a.base_h * a.base_w * geo.taps.n * C * 8,
                  4.0 * B * (ep.in_h * ep.in_w * C + p.hw * 8));
        if (launch_gather_any(a, geo, C, 8, B, st, (c.d->reserved & 1) != 0)) return 1;
        break;
      }
      case SENAS_KIND_DEPSEP: {
        if (t.dwg) break;  // forward_dw_group
        Geo geo = make_geo(t.k, 1, ed.op_type, DIR_FWD);
        DwArgs a;
        a.x = x, a.x_ld = x_ld, a.x_h = ep.in_h, a.x_w = ep.in_w, a.z = c.saved + t.z_off;
        a.o_h = p.out_h, a.o_w = p.out_w;
        a.base_h = geo.base_is_out ? p.out_h : ep.in_h, a.base_w = geo.base_is_out ? p.out_w : ep.in_w;
        a.si = geo.si, a.so = geo.so, a.w = (const float *)ed.param[k][0], a.partials = c.scratch + t.part1_off;
        a.taps = geo.taps;
        dim3 grid(t.nblk1, B);
        SENAS_TAG("dw_fwd", 2.0 * B * a.base_h * a.base_w * geo.taps.n * C, 4.0 * B * (ep.in_h * ep.in_w * C + p.hw * C));
        if (false && geo.so == 1) {  // sliding-window variant: measured slower than the float4 kernel (more instructions)
#define SENAS_DWF(CC, KK, SS)                                                                                   \
  {                                                                                                             \
    auto kern = dw_sw_kernel<CC, KK, SS, false, true>;                                                          \
    SENAS_LAUNCH(kern, grid, dim3(256), 0, st, x, x_ld, ep.in_h, ep.in_w, a.z, (int64_t)CC, a.base_h,     \
                 a.base_w, a.w, 0, a.partials, kDwRows);                                                        \
  }
          if (C == 32 && t.k == 5 && geo.si == 1) SENAS_DWF(32, 5, 1)
          else if (C == 32 && t.k == 5 && geo.si == 2) SENAS_DWF(32, 5, 2)
          else if (C == 32 && t.k == 3 && geo.si == 1) SENAS_DWF(32, 3, 1)
          else if (C == 32 && t.k == 3 && geo.si == 2) SENAS_DWF(32, 3, 2)
          else if (C == 8 && t.k == 5 && geo.si == 1) SENAS_DWF(8, 5, 1)
          else if (C == 8 && t.k == 3 && geo.si == 1) SENAS_DWF(8, 3, 1)
          else SENAS_FAIL("dw fwd: unsupported geometry");
        } else if (C == 32) {
          auto kern = dw_fwd_kernel<32>;
          SENAS_LAUNCH(kern, grid, dim3(128), 0, st, a);
        } else {
          auto kern = dw_fwd_kernel<8>;
          SENAS_LAUNCH(kern, grid, dim3(128), 0, st, a);
        }
        break;
      }
    }
  }
  return 0;
}

static int check_common(senas_graph *g, int batch, const int32_t *ih, const int32_t *iw, Plan **p) {
  if (!g) SENAS_FAIL("null graph");
  return build_plan(g, batch, ih, iw, p);
}

static int check_cuda(const char *what) {
  cudaError_t e = cudaGetLastError();
#ifndef SENAS_EMU
  if (g_launch_err != cudaSuccess) {
    const cudaError_t le = g_launch_err;
    g_launch_err = cudaSuccess;
    SENAS_FAIL("%s: launch of a '%s' kernel failed: %s", what, g_launch_err_tag, cudaGetErrorString(le));
  }
#endif
  if (e != cudaSuccess) SENAS_FAIL("%s: CUDA error %s", what, cudaGetErrorString(e));
  return 0;
}

extern "C" int senas_graph_forward(senas_graph_t *g, const senas_fwd_args_t *a) {
  Plan *p;
  if (!a) SENAS_FAIL("null args");
  if (check_common(g, a->batch, a->in_h, a->in_w, &p)) return 1;
  const senas_graph_desc_t &d = g->d;
  if (!a->alpha || !a->out || !a->saved || !a->scratch) SENAS_FAIL("forward: null alpha/out/saved/scratch");
  for (int i = 0; i < d.n_inputs; ++i)
    if (!a->in[i] || (a->in_ld[i] & 3) || ((uintptr_t)a->in[i] & 15))
      SENAS_FAIL("forward: input %d must be non-null, 16-byte aligned, ld multiple of 4", i);
  if ((a->out_ld & 3) || ((uintptr_t)a->out & 15)) SENAS_FAIL("forward: out must be 16-byte aligned, ld multiple of 4");
  Call c;
  c.d = &d, c.p = p, c.B = a->batch, c.stream = a->stream;
  c.saved = (float *)a->saved, c.scratch = (float *)a->scratch;
  for (int i = 0; i < 2; ++i) c.in[i] = a->in[i], c.in_ld[i] = a->in_ld[i];
  c.out = a->out, c.out_ld = a->out_ld;
  memset(&c.bases, 0, sizeof(c.bases));
  c.bases.p[SP_SAVED] = c.saved, c.bases.p[SP_SCRATCH] = c.scratch;
  c.bases.p[SP_IN0] = (float *)a->in[0], c.bases.ld[SP_IN0] = a->in_ld[0];
  c.bases.p[SP_IN1] = (float *)a->in[1], c.bases.ld[SP_IN1] = a->in_ld[1];
  c.bases.p[SP_OUT] = a->out, c.bases.ld[SP_OUT] = a->out_ld;
  c.S.init(a->stream);
#ifndef SENAS_EMU
  for (int i = 0; i < d.n_inputs; ++i) {
    if (p->xb_off[i] < 0) continue;
    const int64_t npix = (int64_t)c.B * p->in_h[i] * p->in_w[i];
    SENAS_TAG("cast_bf16", 0, 6.0 * npix * 32);
    bool down = false;  // inputs of DOWN groups get the phase-major copy
    for (const TcGroup &g2 : p->tc_groups) down |= g2.src == i && g2.op == SENAS_OP_DOWN;
    SENAS_LAUNCH(cast_bf16_kernel, dim3((unsigned)((npix * 4 + 255) / 256)), dim3(256), 0, c.stream, a->in[i], a->in_ld[i],
                 reinterpret_cast<__nv_bfloat16 *>(c.saved + p->xb_off[i]), npix, down ? p->in_h[i] / 2 : 0,
                 down ? p->in_w[i] / 2 : 0);
  }
#endif
  c.S.fork();
#ifndef SENAS_EMU
  for (const TcGroup &g2 : p->tc_groups) {
    Geo geo = make_geo(g2.k, g2.dil, g2.op, DIR_FWD);
    TcConvArgs ta;
    memset(&ta, 0, sizeof(ta));
    const EdgePlan &ep0 = p->edges[g2.edge[0]];
    ta.nterms = g2.nterms;
    for (int i = 0; i < g2.nterms; ++i) {
      const TermPlan &t = p->edges[g2.edge[i]].t[g2.cand[i]];
      ta.w[i] = (const float *)d.edge[g2.edge[i]].param[g2.cand[i]][0];
      ta.y[i] = c.saved + t.y_off, ta.partials[i] = c.scratch + t.part_off;
    }
    ta.ws_t = 1;
    conv_weight_strides(g2.op, 32, g2.k * g2.k, DIR_FWD, &ta.ws_k, &ta.ws_n);
    const __nv_bfloat16 *xb = reinterpret_cast<const __nv_bfloat16 *>(c.saved + p->xb_off[g2.src]);
    void *gst = c.S.stream(c.S.pick());
    if (g2.op == SENAS_OP_DOWN) {  // one launch per non-empty input phase, accumulating into y; statistics on the last
      TapTable pt[4];
      down_phase_tables(geo.taps, pt);
      const int last = down_last_phase(pt);
      bool first = true;
      ta.H = p->out_h, ta.W = p->out_w, ta.Ho = p->out_h, ta.Wo = p->out_w, ta.so = 1, ta.img_mul = 4;
      for (int ph = 0; ph < 4; ++ph) {
        if (pt[ph].n == 0) continue;
        ta.taps = pt[ph], ta.img_add = ph, ta.acc_y = first ? 0 : 1, ta.no_stats = ph == last ? 0 : 1;
        ta.rows_per_cta = tc_rows(p->out_h, p->out_w, c.B, pt[ph].max_dy - pt[ph].min_dy);
        ta.row_chunks = cdiv(p->out_h, ta.rows_per_cta);
        SENAS_TAG("conv_tc_fwd", 2.0 * c.B * p->hw * pt[ph].n * 32 * 8 * g2.nterms,
                  2.0 * c.B * p->hw * 32 + 4.0 * c.B * p->hw * 8 * g2.nterms);
        const int rc = launch_conv_tc(xb, c.B, ta, gst);
        if (rc) SENAS_FAIL("tcgen05 conv launch failed (code %d)", rc);
        first = false;
      }
      continue;
    }
    ta.H = ep0.in_h, ta.W = ep0.in_w, ta.Ho = p->out_h, ta.Wo = p->out_w, ta.so = geo.so;
    ta.rows_per_cta = tc_rows(ep0.in_h, ep0.in_w, c.B, geo.taps.max_dy - geo.taps.min_dy);
    ta.row_chunks = cdiv(ep0.in_h, ta.rows_per_cta), ta.taps = geo.taps;
    SENAS_TAG("conv_tc_fwd", 2.0 * c.B * ep0.in_h * ep0.in_w * geo.taps.n * 32 * 8 * g2.nterms,
              2.0 * c.B * ep0.in_h * ep0.in_w * 32 + 4.0 * c.B * p->hw * 8 * g2.nterms);
    const int rc = launch_conv_tc(xb, c.B, ta, gst);
    if (rc) SENAS_FAIL("tcgen05 conv launch failed (code %d)", rc);
  }
#endif
  for (int s = 0; s < d.n_nodes; ++s) {
    if (s > 0) c.S.fork();
    for (int src = 0; src < d.n_inputs + d.n_nodes; ++src) {
      if (state_stage(d, src) != s) continue;
      if (forward_ds_group(c, src, DS_FWD_STATS)) return 1;
      if (!g_dw_fwd_per_edge) {
        if (forward_dw_group(c, src, -1)) return 1;
        continue;
      }
      for (int e = 0; e < d.n_edges; ++e)
        if (d.edge[e].src == src && forward_dw_group(c, src, e)) return 1;
    }
    for (int e = 0; e < d.n_edges; ++e)
      if (state_stage(d, d.edge[e].src) == s && forward_edge(c, e, false)) return 1;
    c.S.join();
    if (p->n_bnA[s]) {
      SENAS_TAG("bn_reduce", 0, 0);
      SENAS_LAUNCH(bn_reduce_kernel, dim3(p->n_bnA[s], c.B), dim3(256), 0, c.stream, (const BnDesc *)p->d_bnA[s], c.bases);
      SENAS_TAG("bn_finalize", 0, 0);
      SENAS_LAUNCH(bn_finalize_kernel, dim3(p->n_bnA[s]), dim3(64), 0, c.stream, (const BnDesc *)p->d_bnA[s], c.bases,
                   c.B, a->training);
    }
    if (p->n_bnB[s]) {
      c.S.fork();
      for (int src = 0; src < d.n_inputs + d.n_nodes; ++src)
        if (state_stage(d, src) == s && forward_ds_group(c, src, DS_FWD_Y)) return 1;
      for (int e = 0; e < d.n_edges; ++e)
        if (state_stage(d, d.edge[e].src) == s && forward_edge(c, e, true)) return 1;
      c.S.join();
      SENAS_TAG("bn_reduce", 0, 0);
      SENAS_LAUNCH(bn_reduce_kernel, dim3(p->n_bnB[s], c.B), dim3(256), 0, c.stream, (const BnDesc *)p->d_bnB[s], c.bases);
      SENAS_TAG("bn_finalize", 0, 0);
      SENAS_LAUNCH(bn_finalize_kernel, dim3(p->n_bnB[s]), dim3(64), 0, c.stream, (const BnDesc *)p->d_bnB[s], c.bases,
                   c.B, a->training);
    }
    const int thr = std::min(1024, ((c.B * 8 + 31) / 32) * 32);
    SENAS_TAG("node_coef", 0, 0);
    SENAS_LAUNCH(node_coef_kernel, dim3(1), dim3(thr), 0, c.stream, (const NodeDesc *)p->d_nodes, s, c.bases, a->alpha,
                 a->beta, c.B);
    SENAS_TAG("node_combine", 0, 4.0 * c.B * p->hw * 8 * (1 + 5 * (s + 2)));
    SENAS_LAUNCH(node_combine_kernel, dim3(cdiv(p->hw, 128), c.B), dim3(128), 0, c.stream, (const NodeDesc *)p->d_nodes,
                 s, c.bases, d.node_relu);
  }
  return check_cuda("forward");
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
struct BwdCall : Call {
  const senas_bwd_args_t *a;
  std::vector<int> dw_wait[2 + SENAS_MAX_NODES];  // lanes whose dz (NORM dep-sep candidates reading a state) must finish
  float *dstate[2 + SENAS_MAX_NODES];
  int64_t dstate_ld[2 + SENAS_MAX_NODES];
  bool touched[2 + SENAS_MAX_NODES];
};

static int backward_edge(BwdCall &c, int e) {
  const senas_edge_desc_t &ed = c.d->edge[e];
  const EdgePlan &ep = c.p->edges[e];
  const Plan &p = *c.p;
  const int B = c.B, C = ed.c_in, HW = p.hw;
  int64_t x_ld;
  const float *x = state_ptr(c, ed.src, &x_ld);
  float *dx = c.dstate[ed.src];
  const int64_t dx_ld = c.dstate_ld[ed.src];
  const float *gm = c.scratch + p.nodes[ed.dst].gm_off;
  float *gp = c.a->grad_params;
  const int dxl = c.S.dx_lane(ed.src);  // every accumulation into dx[src] runs on this lane, in issue order
  void *sdx = c.S.stream(dxl);
  for (int k = 0; k < SENAS_MAX_CAND; ++k) {
    const TermPlan &t = ep.t[k];
    if (!t.has_y) continue;
    // dep-sep / quad adapters start with a chain that feeds the data gradient; conv / old adapters only have a weight
    // gradient left on their lane
    const bool chain = t.kind == SENAS_KIND_DEPSEP || ((t.kind == SENAS_KIND_IDENTITY || t.kind == SENAS_KIND_UP_SAMPLE || t.kind == SENAS_KIND_AVG_POOL) && C == 32);
    const int ln = chain ? c.S.pick_chain() : c.S.pick_bg();
    void *st = c.S.stream(ln);
    float *tmp = c.tmp(ln);
    int64_t y_ld = 8;
    const float *y = t.owns_y ? c.saved + t.y_off : x;
    if (!t.owns_y) y_ld = x_ld;
    const float *cA = c.scratch + t.coef_off, *cB = cA + B * 8, *cC = cB + B * 8;
    const int T = t.k * t.k;
    switch (t.kind) {
      case SENAS_KIND_IDENTITY:
      case SENAS_KIND_AVG_POOL:
      case SENAS_KIND_UP_SAMPLE: {
        AdapterBwdArgs a;
        memset(&a, 0, sizeof(a));
        a.x = x, a.x_ld = x_ld, a.x_h = ep.in_h, a.x_w = ep.in_w, a.gm = gm, a.y = y, a.y_ld = y_ld;
        a.o_h = p.out_h, a.o_w = p.out_w, a.coefA = cA, a.coefB = cB, a.coefC = cC;
        a.w = (const float *)ed.param[k][0], a.dx = dx, a.dx_ld = dx_ld, a.accumulate = c.touched[ed.src];
        a.partials = tmp;
        const int kk = t.kind == SENAS_KIND_IDENTITY ? AD_IDENTITY : (t.kind == SENAS_KIND_AVG_POOL ? AD_POOL : AD_UP);
        const int in_px = ep.in_h * ep.in_w;
        if (C == 32) {  // one quad-layout sweep over the input grid: dx += W^T.dy and dW partials
          const bool up = kk != AD_IDENTITY, want_dw = ed.grad_off[k][0] >= 0 && !c.a->skip_wgrad;  // up: dy first goes through U^T / P^T
          float *du = tmp, *dwp = tmp + (up ? (int64_t)B * in_px * 8 : 0);
          if (kk == AD_UP) {
            SENAS_TAG("adapter_dx", 0, 4.0 * B * HW * 16);
            SENAS_LAUNCH(up8_bwd_kernel, dim3(cdiv(in_px, 128), B), dim3(128), 0, st, a, du);
          } else if (kk == AD_POOL) {
            SENAS_TAG("adapter_dx", 0, 4.0 * B * HW * 16);
            SENAS_LAUNCH(pool8_bwd_kernel, dim3(cdiv(in_px, 128), B), dim3(128), 0, st, a, du);
          }
          LinBwdArgs la;
          memset(&la, 0, sizeof(la));
          la.x = x, la.x_ld = x_ld, la.gm = up ? du : gm, la.y = y, la.w = a.w, la.hw = in_px;
          if (!up) la.coefA = cA, la.coefB = cB, la.coefC = cC;
          la.dx = dx, la.dx_ld = dx_ld, la.accumulate = c.touched[ed.src], la.partials = want_dw ? dwp : nullptr;
          const int nb = cdiv(in_px, kPwPx);
          c.S.dep(ln, dxl);  // (du ready; also orders the use of this lane's tmp slice)
          SENAS_TAG("adapter_dx", 4.0 * B * in_px * C * 8, 4.0 * B * (in_px * 3 * C + (up ? in_px : HW) * 16));
          auto kern = lin_bwd_q_kernel<32>;
          SENAS_LAUNCH(kern, dim3(nb, B), dim3(256), 0, sdx, la, kPwPx);
          if (dx) c.touched[ed.src] = true;
          c.S.dep(dxl, ln);
          if (want_dw) {
            SENAS_TAG("reduce", 0, 0);
            SENAS_LAUNCH(rows_reduce_kernel, dim3(cdiv(8 * C, kRowsReduceCols), 1), dim3(kRowsReduceThreads), 0, st, (const float *)dwp,
                         gp + ed.grad_off[k][0], nb * B, 8 * C);
          }
          break;
        }
        if (dx) {
          dim3 grid(cdiv(in_px, 128), B);
#define SENAS_AD_DX(CC, KK)                                    \
  {                                                            \
    auto kern = adapter_dx_kernel<CC, KK>;                     \
    SENAS_TAG("adapter_dx", 2.0 * B * in_px * CC * 8, 4.0 * B * (in_px * CC + HW * 16)); \
    SENAS_LAUNCH(kern, grid, dim3(128), 0, sdx, a);            \
  }
          if (C == 32 && kk == AD_IDENTITY) SENAS_AD_DX(32, AD_IDENTITY)
          else if (C == 32 && kk == AD_POOL) SENAS_AD_DX(32, AD_POOL)
          else if (C == 32 && kk == AD_UP) SENAS_AD_DX(32, AD_UP)
          else if (C == 8 && kk == AD_IDENTITY) SENAS_AD_DX(8, AD_IDENTITY)
          else SENAS_FAIL("adapter dx: unsupported kind/c_in");
          c.touched[ed.src] = true;
        }
        if (a.w != nullptr && ed.grad_off[k][0] >= 0 && !c.a->skip_wgrad) {
          const int gpx = kk == AD_POOL ? HW : in_px;
          dim3 grid(cdiv(gpx, 128 * kPxTilesPerBlock), B);
#define SENAS_AD_DW(CC, KK)                                    \
  {                                                            \
    auto kern = adapter_dw_kernel<CC, KK>;                     \
    SENAS_TAG("adapter_dw", 2.0 * B * gpx * CC * 8, 4.0 * B * (in_px * CC + HW * 16)); \
    SENAS_LAUNCH(kern, grid, dim3(128), 0, st, a);             \
  }
          if (kk == AD_IDENTITY) SENAS_AD_DW(32, AD_IDENTITY)
          else if (kk == AD_POOL) SENAS_AD_DW(32, AD_POOL)
          else SENAS_AD_DW(32, AD_UP)
          const int n = 8 * C;
          SENAS_TAG("reduce", 0, 0);
          SENAS_LAUNCH(rows_reduce_kernel, dim3(cdiv(n, kRowsReduceCols), 1), dim3(kRowsReduceThreads), 0, st, (const float *)tmp,
                       gp + ed.grad_off[k][0], (int)(grid.x * B), n);
        }
        break;
      }
      case SENAS_KIND_CONV:
      case SENAS_KIND_SE_CONV: {
        if (dx && !t.tc) {  // tcgen05 groups: data gradient at the end of backward
          Geo geo = make_geo(t.k, t.dil, ed.op_type, DIR_DGRAD);
          GatherArgs a;
          memset(&a, 0, sizeof(a));
          a.src = gm, a.src_ld = 8, a.src_h = p.out_h, a.src_w = p.out_w, a.src2 = y, a.src2_ld = y_ld;
          a.coefA = cA, a.coefB = cB, a.coefC = cC;
          a.dst = dx, a.dst_ld = dx_ld, a.dst_h = ep.in_h, a.dst_w = ep.in_w, a.accumulate = c.touched[ed.src];
          a.base_h = geo.base_is_out ? p.out_h : ep.in_h, a.base_w = geo.base_is_out ? p.out_w : ep.in_w;
          a.si = geo.si, a.so = geo.so, a.w = (const float *)ed.param[k][0], a.ws_t = 1;
          conv_weight_strides(ed.op_type, C, T, DIR_DGRAD, &a.ws_k, &a.ws_n);
          a.partials = nullptr, a.taps = geo.taps;
          SENAS_TAG(ed.op_type == SENAS_OP_DOWN ? "conv_dgrad.down" : (C == 8 ? "conv_dgrad.n8" : (ed.op_type == SENAS_OP_UP ? "conv_dgrad.up" : "conv_dgrad.n32_small")),
                    2.0 * B * a.base_h * a.base_w * geo.taps.n * C * 8,
                    4.0 * B * (ep.in_h * ep.in_w * C + HW * 16));
          if (launch_gather_any(a, geo, 8, C, B, sdx, (c.d->reserved & 1) != 0)) return 1;
          c.touched[ed.src] = true;
        }
        if (ed.grad_off[k][0] >= 0 && !(t.tc && c.a->grad_in[ed.src]) && !c.a->skip_wgrad) {
          Geo geo = make_geo(t.k, t.dil, ed.op_type, DIR_FWD);
          WgradArgs a;
          memset(&a, 0, sizeof(a));
          a.x = x, a.x_ld = x_ld, a.x_h = ep.in_h, a.x_w = ep.in_w, a.gm = gm, a.y = y, a.y_ld = y_ld;
          a.o_h = p.out_h, a.o_w = p.out_w, a.coefA = cA, a.coefB = cB, a.coefC = cC;
          a.base_h = geo.base_is_out ? p.out_h : ep.in_h, a.base_w = geo.base_is_out ? p.out_w : ep.in_w;
          a.si = geo.si, a.so = geo.so;
          a.batch = B, a.partials = tmp, a.taps = geo.taps;
          int nblk = 0;
#define SENAS_WGRAD2(KC, KK, SI_, SO_)                                                                             \
  {                                                                                                                \
    using TL = WgradTile<KC, KK, SI_, SO_>;                                                                        \
    a.tiles_x = cdiv(a.base_w, TL::TW), a.tiles_y = cdiv(a.base_h, TL::TH);                                        \
    nblk = std::min(a.tiles_x * a.tiles_y * B, kPersistBlocks);                                                    \
    const int XR = (TL::TH - 1) * SI_ + geo.taps.max_dy - geo.taps.min_dy + 1;                                     \
    const int XC = (TL::TW - 1) * SI_ + geo.taps.max_dx - geo.taps.min_dx + 1;                                     \
    const size_t smem = (size_t)((XR * XC * KC + 3) / 4) * 16 + (size_t)2 * TL::TH * SO_ * TL::TW * SO_ * 16;     \
    auto kern = conv_wgrad2_kernel<KC, KK, SI_, SO_>;                                                              \
    allow_smem(kern, smem);                                                                                        \
    SENAS_TAG(SI_ == 2 ? "conv_wgrad.down" : (KC == 8 ? "conv_wgrad.n8" : (SO_ == 2 ? "conv_wgrad.up" : "conv_wgrad.n32_small")), \
              2.0 * B * a.base_h * a.base_w * T * KC * 8, 4.0 * B * (ep.in_h * ep.in_w * KC + HW * 16)); \
    SENAS_LAUNCH(kern, dim3(nblk), dim3(TL::THREADS), smem, st, a);                                          \
  }
          const int si_ = geo.si, so_ = geo.so;
          if (C == 8 && t.k == 5 && si_ == 1 && so_ == 1 && geo.taps.n == kWmTaps && (c.d->reserved & 1) && g_wgrad_mma) {
            // bf16 mode: the 8 -> 8 node edges' weight gradient on the tensor cores (mma.sync TF32, pixels = GEMM-K)
            a.tiles_x = cdiv(a.base_w, kWmTW), a.tiles_y = cdiv(a.base_h, kWmTH);
            nblk = std::min(a.tiles_x * a.tiles_y * B, kPersistBlocks);
            const int XR = kWmTH + geo.taps.max_dy - geo.taps.min_dy, XC = kWmTW + geo.taps.max_dx - geo.taps.min_dx;
            const size_t smem = (size_t)(XR * XC * 8 + kWmTH * kWmTW * 8 + 4 * kWmTaps * 64) * sizeof(float);
            auto kern = conv_wgrad_mma8_kernel;
            allow_smem(kern, smem);
            SENAS_TAG("conv_wgrad.n8", 2.0 * B * a.base_h * a.base_w * T * 8 * 8, 4.0 * B * (ep.in_h * ep.in_w * 8 + HW * 16));
            SENAS_LAUNCH(kern, dim3(nblk), dim3(kWmThreads), smem, st, a);
          } else
          if (C == 32 && t.k == 5 && si_ == 1 && so_ == 1) SENAS_WGRAD2(32, 5, 1, 1)
          else if (C == 32 && t.k == 5 && si_ == 2 && so_ == 1) SENAS_WGRAD2(32, 5, 2, 1)
          else if (C == 32 && t.k == 5 && si_ == 1 && so_ == 2) SENAS_WGRAD2(32, 5, 1, 2)
          else if (C == 32 && t.k == 3 && si_ == 2 && so_ == 1) SENAS_WGRAD2(32, 3, 2, 1)
          else if (C == 32 && t.k == 3 && si_ == 1 && so_ == 2) SENAS_WGRAD2(32, 3, 1, 2)
          else if (C == 32 && t.k == 3 && si_ == 1 && so_ == 1) SENAS_WGRAD2(32, 3, 1, 1)
          else if (C == 8 && t.k == 5 && si_ == 1 && so_ == 1) SENAS_WGRAD2(8, 5, 1, 1)
          else if (C == 8 && t.k == 3 && si_ == 1 && so_ == 1) SENAS_WGRAD2(8, 3, 1, 1)
          else SENAS_FAIL("conv wgrad: unsupported c_in %d k %d geometry", C, t.k);
          int ws_ci, ws_co;
          conv_weight_strides(ed.op_type, C, T, DIR_FWD, &ws_ci, &ws_co);
          const int n = T * C * 8;
          SENAS_TAG("reduce", 0, 0);
          SENAS_LAUNCH(wgrad_reduce_kernel, dim3(cdiv(n, 32)), dim3(256), 0, st, (const float *)tmp, nblk, T, C,
                       gp + ed.grad_off[k][0], 1, ws_ci, ws_co, geo.taps);
        }
        break;
      }
      case SENAS_KIND_DEPSEP: {
        if (t.fused) {  // the dep-sep candidates of the edge as ONE chain: stats sweep -> fold -> dz sweep (z recomputed)
          bool first = true;
          for (int k2 = 0; k2 < k; ++k2) first &= !(ep.t[k2].kind == SENAS_KIND_DEPSEP && ep.t[k2].fused);
          if (!first) break;
          DsArgs da;
          memset(&da, 0, sizeof(da));
          da.x = x, da.x_ld = x_ld, da.H = ep.in_h, da.W = ep.in_w, da.batch = B;
          da.tiles_x = cdiv(ep.in_w, ds_tile_w(C)), da.tile_rows = ds_rows(C, B, ep.in_h, ep.in_w);
          const int nb = t.nblk;
          const int64_t per = (int64_t)B * nb * 10 * C + 16 * C;  // [partials | sums1 (12C) | coef1 (3C)] per candidate
          int ks[SENAS_MAX_CAND], nk = 0;
          double taps = 0;
          for (int k2 = k; k2 < SENAS_MAX_CAND; ++k2) {
            const TermPlan &t2 = ep.t[k2];
            if (t2.kind != SENAS_KIND_DEPSEP || !t2.fused) continue;
            if (ed.grad_off[k2][1] < 0 || ed.grad_off[k2][2] < 0 || ed.grad_off[k2][6] < 0 || ed.grad_off[k2][0] < 0)
              SENAS_FAIL("dep-sep candidate needs gradient slots 0,1,2,6");
            if ((nk + 1) * per > p.tmp_floats) break;  // (cannot happen: tmp is sized for 2 candidates, see build_plan)
            DsItem &it = da.it[da.n++];
            float *base = tmp + nk * per;
            it.w = (const float *)ed.param[k2][0], it.k = t2.k, it.training = c.a->training, it.dz_bf = t2.zb;
            it.mean1 = c.saved + t2.mean1_off, it.istd1 = c.saved + t2.istd1_off;
            it.g1 = (const float *)ed.param[k2][1], it.b1 = (const float *)ed.param[k2][2];
            it.wpw = (const float *)ed.param[k2][6], it.y = c.saved + t2.y_off, it.gm = gm;
            it.coef = c.scratch + t2.coef_off, it.dz = c.saved + t2.z_off;
            it.partials = base, it.bn1_coef = base + (int64_t)B * nb * 10 * C + 12 * C;
            ks[nk++] = k2;
            taps += t2.k * t2.k;
          }
          dim3 grid(nb, B);
          const double px = (double)B * HW;
          SENAS_TAG("ds_bwd_stats", 2.0 * px * (taps * C + 24.0 * C * nk), 4.0 * px * (C + 16.0 * nk));
          if (C == 32) {
            auto kern = ds_norm_kernel<32, DS_BWD_STATS>;
            SENAS_LAUNCH(kern, grid, dim3(128), 0, st, da);
          }
          else {
            auto kern = ds_norm_kernel<8, DS_BWD_STATS>;
            SENAS_LAUNCH(kern, grid, dim3(128), 0, st, da);
          }
          for (int q = 0; q < nk; ++q) {
            const int k2 = ks[q];
            float *base = tmp + q * per, *sums = base + (int64_t)B * nb * 10 * C;
            SENAS_TAG("reduce", 0, 0);  // fold + BN1-backward finalize in one launch
            SENAS_LAUNCH(pw_reduce_fin_kernel, dim3(cdiv(10 * C, kRowsReduceCols), 1), dim3(kRowsReduceThreads), 0, st,
                         (const float *)base, nb * B, C, (float)B * (float)HW, da.it[q].g1, da.it[q].istd1, sums + 12 * C,
                         gp + ed.grad_off[k2][1], gp + ed.grad_off[k2][2], gp + ed.grad_off[k2][6]);
          }
          SENAS_TAG("ds_bwd_dz", 2.0 * px * (taps * C + 16.0 * C * nk), 4.0 * px * (C + 16.0 * nk + (double)C * nk));
          if (C == 32) {
            auto kern = ds_norm_kernel<32, DS_BWD_DZ>;
            SENAS_LAUNCH(kern, grid, dim3(128), 0, st, da);
          }
          else {
            auto kern = ds_norm_kernel<8, DS_BWD_DZ>;
            SENAS_LAUNCH(kern, grid, dim3(128), 0, st, da);
          }
          c.dw_wait[ed.src].push_back(ln);  // data / weight gradient of the depthwise halves: backward_dw_group
          break;
        }
        PwBwdArgs a;
        memset(&a, 0, sizeof(a));
        float *sums1 = nullptr, *coef1 = nullptr;
        a.gm = gm, a.y = y, a.z = c.saved + t.z_off, a.hw = HW, a.batch = B, a.coefA = cA, a.coefB = cB, a.coefC = cC;
        a.mean1 = c.saved + t.mean1_off, a.istd1 = c.saved + t.istd1_off;
        a.g1 = (const float *)ed.param[k][1], a.b1 = (const float *)ed.param[k][2];
        a.wpw = (const float *)ed.param[k][6], a.partials = tmp, a.bn1_coef = coef1, a.z_bf = t.zb;
        const int px_pb = 512, nblk_cc = cdiv(HW, px_pb);
        dim3 grid_cc(nblk_cc, B);
        if (ed.grad_off[k][1] < 0 || ed.grad_off[k][2] < 0 || ed.grad_off[k][6] < 0 || ed.grad_off[k][0] < 0)
          SENAS_FAIL("dep-sep candidate needs gradient slots 0,1,2,6");
        sums1 = tmp + (int64_t)B * nblk_cc * 10 * C, coef1 = sums1 + 12 * C;
        a.bn1_coef = coef1;
        SENAS_TAG("pw_bwd_stats", 4.0 * B * HW * C * 8, 4.0 * B * HW * (C + 16));
        // (round 2: a version with du = dy.W and dW += dy^T.a as mma.sync TF32 -- 16 pixels per warp step, z read in the D and
        // the B fragment layout -- was correct (emulator + GPU suite) and SLOWER: 11.9 vs 10.6 ms per step; it trades FMAs for
        // twice as many narrow load instructions and the sweep is bound by those.  Removed.)
        if (C == 32) {  // (2 / 3 / 4 pixel steps in flight measured the same under the 128-register cap: 2)
          auto kern = pw_bwd_q_kernel<32, 1>;
          SENAS_LAUNCH(kern, grid_cc, dim3(256), 0, st, a, px_pb, c.a->training);
        } else {
          auto kern = pw_bwd_q_kernel<8, 1>;
          SENAS_LAUNCH(kern, grid_cc, dim3(256), 0, st, a, px_pb, c.a->training);
        }
        SENAS_TAG("reduce", 0, 0);  // fold + BN1-backward finalize in one launch
        SENAS_LAUNCH(pw_reduce_fin_kernel, dim3(cdiv(10 * C, kRowsReduceCols), 1), dim3(kRowsReduceThreads), 0, st,
                     (const float *)tmp, (int)(nblk_cc * B), C, (float)B * (float)HW, a.g1, a.istd1, coef1,
                     gp + ed.grad_off[k][1], gp + ed.grad_off[k][2], gp + ed.grad_off[k][6]);
        SENAS_TAG("pw_bwd_dz", 2.0 * B * HW * C * 8, 4.0 * B * HW * (2 * C + 16));
        if (C == 32) {
          auto kern = pw_bwd_q_kernel<32, 2>;
          SENAS_LAUNCH(kern, grid_cc, dim3(256), 0, st, a, px_pb, c.a->training);
        } else {
          auto kern = pw_bwd_q_kernel<8, 2>;
          SENAS_LAUNCH(kern, grid_cc, dim3(256), 0, st, a, px_pb, c.a->training);
        }
        if (t.dwg) {  // data / weight gradient of the depthwise half: backward_dw_group
          c.dw_wait[ed.src].push_back(ln);
          break;
        }
        DwBwdArgs w;
        memset(&w, 0, sizeof(w));
        w.dz = c.saved + t.z_off, w.z_h = p.out_h, w.z_w = p.out_w, w.x_h = ep.in_h, w.x_w = ep.in_w;
        w.w = (const float *)ed.param[k][0];
        if (dx) {
          Geo geo = make_geo(t.k, 1, ed.op_type, DIR_DGRAD);
          w.dx = dx, w.dx_ld = dx_ld, w.accumulate = c.touched[ed.src];
          w.base_h = geo.base_is_out ? p.out_h : ep.in_h, w.base_w = geo.base_is_out ? p.out_w : ep.in_w;
          w.si = geo.si, w.so = geo.so, w.taps = geo.taps;
          dim3 g2(cdiv(w.base_h * w.base_w, 128 / (C / 4)), B);
          c.S.dep(ln, dxl);  // dz of this candidate is ready
          SENAS_TAG("dw_dx", 2.0 * B * w.base_h * w.base_w * T * C, 4.0 * B * (HW * C + 2 * ep.in_h * ep.in_w * C));
          if (false && ed.op_type == SENAS_OP_NORM) {  // sliding-window variant: measured slower, kept for reference
            dim3 g2s(cdiv(ep.in_h, kDwRows), B);
#define SENAS_DWB(CC, KK)                                                                                        \
  {                                                                                                              \
    auto kern = dw_sw_kernel<CC, KK, 1, true, false>;                                                            \
    SENAS_LAUNCH(kern, g2s, dim3(256), 0, sdx, (const float *)w.dz, (int64_t)CC, p.out_h, p.out_w, dx, dx_ld, \
                 ep.in_h, ep.in_w, w.w, (int)c.touched[ed.src], (float *)nullptr, kDwRows);                      \
  }
            if (C == 32 && t.k == 5) SENAS_DWB(32, 5)
            else if (C == 32 && t.k == 3) SENAS_DWB(32, 3)
            else if (C == 8 && t.k == 5) SENAS_DWB(8, 5)
            else if (C == 8 && t.k == 3) SENAS_DWB(8, 3)
            else SENAS_FAIL("dw dx: unsupported geometry");
          } else if (C == 32) {
            auto kern = dw_dx_kernel<32>;
            SENAS_LAUNCH(kern, g2, dim3(128), 0, sdx, w);
          } else {
            auto kern = dw_dx_kernel<8>;
            SENAS_LAUNCH(kern, g2, dim3(128), 0, sdx, w);
          }
          c.touched[ed.src] = true;
        }
        if (!c.a->skip_wgrad) {
          Geo geo = make_geo(t.k, 1, ed.op_type, DIR_FWD);
          w.x = x, w.x_ld = x_ld, w.partials = tmp, w.batch = B, w.chunk = kDwChunk;
          w.base_h = geo.base_is_out ? p.out_h : ep.in_h, w.base_w = geo.base_is_out ? p.out_w : ep.in_w;
          w.si = geo.si, w.so = geo.so, w.taps = geo.taps;
          dim3 g3(cdiv(w.base_h * w.base_w, w.chunk), B);
          SENAS_TAG("dw_wgrad", 2.0 * B * w.base_h * w.base_w * T * C, 4.0 * B * (HW * C + ep.in_h * ep.in_w * C));
#define SENAS_DWSW(CC, KK, SS)                                 \
  {                                                            \
    auto kern = dw_wgrad_sw_kernel<CC, KK, SS>;                \
    SENAS_LAUNCH(kern, g3, dim3(256), 0, st, w);         \
  }
          {  // sliding register window kernels: `chunk` = base rows per block
            w.chunk = 4;
            g3 = dim3(cdiv(w.base_h, w.chunk), B);
          }
#define SENAS_DWUP(CC, KK)                                     \
  {                                                            \
    auto kern = dw_wgrad_up_kernel<CC, KK>;                    \
    SENAS_LAUNCH(kern, g3, dim3(256), 0, st, w);         \
  }
#define SENAS_DWW(CC, TT)                                      \
  {                                                            \
    auto kern = dw_wgrad_kernel<CC, TT>;                       \
    SENAS_LAUNCH(kern, g3, dim3(256), 0, st, w);         \
  }
          if (geo.so == 1 && C == 32 && t.k == 5 && geo.si == 1) SENAS_DWSW(32, 5, 1)
          else if (geo.so == 1 && C == 32 && t.k == 5 && geo.si == 2) SENAS_DWSW(32, 5, 2)
          else if (geo.so == 1 && C == 32 && t.k == 3 && geo.si == 1) SENAS_DWSW(32, 3, 1)
          else if (geo.so == 1 && C == 32 && t.k == 3 && geo.si == 2) SENAS_DWSW(32, 3, 2)
          else if (geo.so == 1 && C == 8 && t.k == 5 && geo.si == 1) SENAS_DWSW(8, 5, 1)
          else if (geo.so == 1 && C == 8 && t.k == 3 && geo.si == 1) SENAS_DWSW(8, 3, 1)
          else if (geo.so == 2 && C == 32 && t.k == 5) SENAS_DWUP(32, 5)
          else if (geo.so == 2 && C == 32 && t.k == 3) SENAS_DWUP(32, 3)
          else if (C == 32 && T == 25) SENAS_DWW(32, 25)
          else if (C == 32 && T == 9) SENAS_DWW(32, 9)
          else if (C == 8 && T == 25) SENAS_DWW(8, 25)
          else if (C == 8 && T == 9) SENAS_DWW(8, 9)
          else SENAS_FAIL("dw wgrad: unsupported c_in %d k %d", C, t.k);
          const int n = C * T;
          SENAS_TAG("reduce", 0, 0);
          SENAS_LAUNCH(rows_reduce_kernel, dim3(cdiv(n, kRowsReduceCols), 1), dim3(kRowsReduceThreads), 0, st, (const float *)tmp,
                       gp + ed.grad_off[k][0], (int)(g3.x * B), n);
        }
        break;
      }
      default:
        break;
    }
  }
  return 0;
}

// data gradient + weight gradient of the depthwise halves (k3 + k5) of one NORM / UP edge, as one launch each
static int backward_dw_group(BwdCall &c, int src, int only_edge) {
  const senas_graph_desc_t &d = *c.d;
  const Plan &p = *c.p;
  DwMultiArgs a;
  memset(&a, 0, sizeof(a));
  int C = 0, h = 0, w = 0, nblk = 0;
  bool up = false, down = false;
  int64_t x_ld = 0;
  const float *x = state_ptr(c, src, &x_ld);
  int64_t goff[kDwMaxItems];
  for (int e = 0; e < d.n_edges; ++e) {
    const senas_edge_desc_t &ed = d.edge[e];
    if (ed.src != src || (only_edge >= 0 && e != only_edge)) continue;
    up = ed.op_type == SENAS_OP_UP, down = ed.op_type == SENAS_OP_DOWN;
    for (int k = 0; k < SENAS_MAX_CAND; ++k) {
      const TermPlan &t = p.edges[e].t[k];
      if (t.kind != SENAS_KIND_DEPSEP || !t.dwg) continue;
      if (a.n == kDwMaxItems) SENAS_FAIL("more than %d dep-sep candidates read state %d", kDwMaxItems, src);
      goff[a.n] = ed.grad_off[k][0];
      DwItem &it = a.it[a.n++];
      it.in = c.saved + t.z_off, it.in_ld = ed.c_in, it.in_bf = t.zb;  // dz (in place over z)
      it.w = (const float *)ed.param[k][0], it.k = t.k, it.flip = (up || down) ? 0 : 1;
      C = ed.c_in, h = down ? p.out_h : p.edges[e].in_h, w = down ? p.out_w : p.edges[e].in_w, nblk = t.nblk1;
    }
  }
  if (a.n == 0) return 0;
  const bool col1 = up || down;  // tiles on the low-resolution grid, one column per thread
  bool lane = !up && !down && dw_lane_ok(C, w, DWL_DX), lane_w = !up && !down && dw_lane_ok(C, w, DWL_WGRAD);
  for (int m = 0; m < a.n; ++m) lane = lane && !a.it[m].in_bf;  // (forward / data-gradient lane kernels: fp32 dz only)
  a.H = h, a.W = w, a.tiles_x = dw_tiles_x(C, w, col1, lane), a.tile_rows = dw_rows(C, c.B, h, w, col1, lane);
  double taps = 0;
  for (int m = 0; m < a.n; ++m) taps += a.it[m].k * a.it[m].k;
  nblk = dw_nblk(C, c.B, h, w, col1, lane);  // (t.nblk1 is the statistics grid: the fused dep-sep kernels tile differently)
  dim3 grid(nblk, c.B);
  if (up && C != 32) SENAS_FAIL("UP depthwise: c_in %d unsupported", C);
  float *dx = c.dstate[src];
  const int dxl = c.S.dx_lane(src), ln = c.S.pick_bg();
  for (int l : c.dw_wait[src]) c.S.dep(l, dxl), c.S.dep(l, ln);
  c.dw_wait[src].clear();
  if (dx) {
    DwMultiArgs g = a;
    for (int m = 0; m < g.n; ++m)
      g.it[m].out = dx, g.it[m].out_ld = c.dstate_ld[src], g.it[m].accumulate = (m > 0 || c.touched[src]) ? 1 : 0;
    SENAS_TAG("dw_dx", 2.0 * c.B * h * w * taps * C, 4.0 * c.B * h * w * C * (1 + g.n * (up ? 4 : 1)));
    if (up) {
      auto kern = dw_up_multi_kernel<32, false, false>;
      SENAS_LAUNCH(kern, grid, dim3(128), 0, c.S.stream(dxl), g);
    } else if (down) {  // dx[2i + p] += sum dz[i + d] w: the scatter onto the high-resolution grid
      auto kern = dw_up_multi_kernel<32, true, false>;
      SENAS_LAUNCH(kern, grid, dim3(128), 0, c.S.stream(dxl), g);
    } else if (lane && C == 32) {
      auto kern = dwl_multi_kernel<32, false>;
      SENAS_LAUNCH(kern, grid, dim3(128), 0, c.S.stream(dxl), g);
    } else if (lane) {
      auto kern = dwl_multi_kernel<8, false>;
      SENAS_LAUNCH(kern, grid, dim3(128), 0, c.S.stream(dxl), g);
    } else if (C == 32) {
      auto kern = dw_multi_kernel<32, false>;
      SENAS_LAUNCH(kern, grid, dim3(128), 0, c.S.stream(dxl), g);
    } else {
      auto kern = dw_multi_kernel<8, false>;
      SENAS_LAUNCH(kern, grid, dim3(128), 0, c.S.stream(dxl), g);
    }
    c.touched[src] = true;
  }
  if (!c.a->skip_wgrad) {
    DwMultiArgs g = a;
    float *tmp = c.tmp(ln);
    void *st = c.S.stream(ln);
    g.tiles_x = dw_tiles_x(C, w, true, lane_w), g.tile_rows = dw_rows(C, c.B, h, w, true, lane_w);
    nblk = dw_nblk(C, c.B, h, w, true, lane_w);
    grid = dim3(nblk, c.B);
    const int64_t per = (int64_t)c.B * nblk * C * 25;
    for (int m = 0; m < g.n; ++m) {
      g.it[m].flip = 0, g.it[m].partials = tmp + m * per;
      if (down)  // low-resolution operand = dz (already in .in), high-resolution operand = x
        g.it[m].in2 = x, g.it[m].in2_ld = (int32_t)x_ld, g.it[m].in2_bf = 0;
      else
        g.it[m].in2 = g.it[m].in, g.it[m].in2_bf = g.it[m].in_bf, g.it[m].in = x, g.it[m].in_ld = x_ld, g.it[m].in_bf = 0;
    }
    SENAS_TAG("dw_wgrad", 2.0 * c.B * h * w * taps * C, 4.0 * c.B * h * w * C * (1 + g.n * (up ? 4 : 1)));
    if (up || down) {
      auto kern = dw_up_wgrad_multi_kernel<32>;
      SENAS_LAUNCH(kern, grid, dim3(128), 0, st, g);
    } else if (lane_w && C == 32) {
      auto kern = dwl_wgrad_multi_kernel<32>;
      SENAS_LAUNCH(kern, grid, dim3(128), 0, st, g);
    } else if (lane_w) {
      auto kern = dwl_wgrad_multi_kernel<8>;
      SENAS_LAUNCH(kern, grid, dim3(128), 0, st, g);
    } else if (C == 32) {
      auto kern = dw_wgrad_multi_kernel<32>;
      SENAS_LAUNCH(kern, grid, dim3(128), 0, st, g);
    } else {
      auto kern = dw_wgrad_multi_kernel<8>;
      SENAS_LAUNCH(kern, grid, dim3(128), 0, st, g);
    }
    for (int m = 0; m < g.n; ++m) {
      const int nn = C * g.it[m].k * g.it[m].k;
      SENAS_TAG("reduce", 0, 0);
      SENAS_LAUNCH(rows_reduce_kernel, dim3(cdiv(nn, kRowsReduceCols), 1), dim3(kRowsReduceThreads), 0, st, (const float *)(tmp + m * per),
                   c.a->grad_params + goff[m], (int)(nblk * c.B), nn);
    }
  }
  return 0;
}

extern "C" int senas_graph_backward(senas_graph_t *g, const senas_bwd_args_t *a) {
  Plan *p;
  if (!a) SENAS_FAIL("null args");
  if (check_common(g, a->batch, a->in_h, a->in_w, &p)) return 1;
  const senas_graph_desc_t &d = g->d;
  if (!a->alpha || !a->out || !a->grad_out || !a->saved || !a->scratch || !a->grad_alpha || !a->grad_params)
    SENAS_FAIL("backward: null pointer argument");
  if ((a->grad_out_ld & 3) || ((uintptr_t)a->grad_out & 15)) SENAS_FAIL("backward: grad_out alignment");
  BwdCall c;
  c.a = a, c.d = &d, c.p = p, c.B = a->batch, c.stream = a->stream;
  c.saved = (float *)a->saved, c.scratch = (float *)a->scratch;
  for (int i = 0; i < 2; ++i) c.in[i] = a->in[i], c.in_ld[i] = a->in_ld[i];
  c.out = (float *)a->out, c.out_ld = a->out_ld;
  memset(&c.bases, 0, sizeof(c.bases));
  c.bases.p[SP_SAVED] = c.saved, c.bases.p[SP_SCRATCH] = c.scratch;
  c.bases.p[SP_IN0] = (float *)a->in[0], c.bases.ld[SP_IN0] = a->in_ld[0];
  c.bases.p[SP_IN1] = (float *)a->in[1], c.bases.ld[SP_IN1] = a->in_ld[1];
  c.bases.p[SP_OUT] = (float *)a->out, c.bases.ld[SP_OUT] = a->out_ld;
  c.bases.p[SP_GOUT] = (float *)a->grad_out, c.bases.ld[SP_GOUT] = a->grad_out_ld;
  for (int i = 0; i < d.n_inputs; ++i) {
    c.dstate[i] = a->grad_in[i], c.dstate_ld[i] = a->grad_in_ld[i], c.touched[i] = false;
    if (a->grad_in[i] && ((a->grad_in_ld[i] & 3) || ((uintptr_t)a->grad_in[i] & 15))) SENAS_FAIL("backward: grad_in alignment");
  }
  for (int i = 0; i < d.n_nodes; ++i) {
    const NodePlan &np = p->nodes[i];
    c.dstate[d.n_inputs + i] = np.has_consumer ? c.scratch + np.dnode_off : nullptr;
    c.dstate_ld[d.n_inputs + i] = 8, c.touched[d.n_inputs + i] = false;
  }
  c.S.init(a->stream);
  cudaMemsetAsync(a->grad_params, 0, sizeof(float) * d.grad_floats, (cudaStream_t)c.stream);
  const int64_t node_bytes = sizeof(float) * (int64_t)c.B * p->hw * 8;
  for (int i = d.n_nodes - 1; i >= 0; --i) {
    const NodePlan &np = p->nodes[i];
    c.S.dep(c.S.dx_lane(d.n_inputs + i), -1);  // the node's gradient from its consumers is complete
    if (np.has_consumer && !c.touched[d.n_inputs + i])
      cudaMemsetAsync(c.scratch + np.dnode_off, 0, node_bytes, (cudaStream_t)c.stream);
    SENAS_TAG("node_bstats", 0, 4.0 * c.B * p->hw * 8 * (3 + 5 * (i + 2)));
    SENAS_LAUNCH(node_bstats_kernel, dim3(np.nblk, c.B), dim3(128), 0, c.stream, (const NodeDesc *)p->d_nodes, i, c.bases,
                 d.node_relu);
    {
      const int V = (1 + np.nterms) * 8;
      SENAS_TAG("reduce", 0, 0);
      SENAS_LAUNCH(rows_reduce_kernel, dim3(cdiv(V, kRowsReduceCols), c.B), dim3(kRowsReduceThreads), 0, c.stream,
                   (const float *)(c.scratch + np.bpart_off), c.scratch + np.bsum_off, np.nblk, V);
    }
    SENAS_TAG("node_bfin", 0, 0);
    int n_in = 0;
    for (int e = 0; e < d.n_edges; ++e) n_in += d.edge[e].dst == i;
    SENAS_LAUNCH(node_bfin_kernel, dim3(n_in, SENAS_MAX_CAND), dim3(128), 0, c.stream, (const NodeDesc *)p->d_nodes, i, c.bases, a->alpha,
                 a->beta, a->grad_alpha, a->grad_beta, a->grad_params, c.B, a->training);
    c.S.fork();  // the candidate chains of this node's edges run on the lanes, concurrently with the next node's sweep
    // the depthwise halves (k3 + k5) of an edge go out together as soon as the edge's dz exist: waiting for all edges
    // of a state (as the forward does to share the x reads) would leave the whole group in the tail of the call
    for (int e = 0; e < d.n_edges; ++e) {
      if (d.edge[e].dst != i) continue;
      if (backward_edge(c, e)) return 1;
      if (g_dw_per_edge && backward_dw_group(c, d.edge[e].src, e)) return 1;
    }
    if (!g_dw_per_edge) {  // per state: once all edges that read it have produced their dz
      if (i > 0) {
        if (backward_dw_group(c, d.n_inputs + i - 1, -1)) return 1;
      } else {
        for (int src = 0; src < d.n_inputs; ++src)
          if (backward_dw_group(c, src, -1)) return 1;
      }
    }
  }
#ifndef SENAS_EMU
  // grouped data gradient + weight gradient of the tcgen05 groups: dx[src] += sum over the group's edges and taps
  // (GEMM-K = 8 x edges).  NORM: one conv over the packed dy.  UP (ConvTranspose2d as 4 output phases on the input grid,
  // y_ph[i] = sum_{t in ph} x[i + d_t] W_t): dy is packed phase-major, and each phase is an ordinary stride-1 problem on
  // the input grid -- dx[i] += sum_{t in ph} dy_ph[i - d_t] W_t^T,  dW_t = sum_i x[i + d_t]^T dy_ph(t)[i].
  for (size_t gi = 0; gi < p->tc_groups.size(); ++gi) {
    const TcGroup &g2 = p->tc_groups[gi];
    if (!a->grad_in[g2.src]) continue;
    const bool up = g2.op == SENAS_OP_UP, down = g2.op == SENAS_OP_DOWN;
    const EdgePlan &ep0 = p->edges[g2.edge[0]];
    // grid of the GEMM-M pixels: the input grid (NORM: = output grid; UP: the 4 output phases live on it), or, for DOWN,
    // the output grid (the 4 INPUT phases live on it)
    const int gh = down ? p->out_h : ep0.in_h, gw = down ? p->out_w : ep0.in_w;
    const int64_t npix = (int64_t)c.B * gh * gw;
    __nv_bfloat16 *dyb = reinterpret_cast<__nv_bfloat16 *>(c.scratch + p->dyb_off[gi]);
    const int lp = c.S.pick_chain(), ln = c.S.pick_bg(), dxl = c.S.dx_lane(g2.src);
    void *st = c.S.stream(ln);
    PackDyArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.nterms = g2.nterms, pa.hw = gh * gw, pa.batch = c.B, pa.dst = dyb;
    if (up) pa.up_h = ep0.in_h, pa.up_w = ep0.in_w;
    TcConvArgs ta;
    memset(&ta, 0, sizeof(ta));
    ta.nterms = g2.nterms, ta.mode = 1;
    for (int i = 0; i < g2.nterms; ++i) {
      const int e = g2.edge[i], k = g2.cand[i];
      const TermPlan &t = p->edges[e].t[k];
      pa.gm[i] = c.scratch + p->nodes[d.edge[e].dst].gm_off, pa.y[i] = c.saved + t.y_off, pa.y_ld[i] = 8;
      pa.coef[i] = c.scratch + t.coef_off;
      ta.w[i] = (const float *)d.edge[e].param[k][0];
    }
    const int nph = up ? 4 : 1;
    SENAS_TAG("pack_dy", 0, npix * nph * (64.0 * g2.nterms + 64.0));
    SENAS_LAUNCH(pack_dy_kernel, dim3((unsigned)((npix * nph * 4 + 255) / 256)), dim3(256), 0, c.S.stream(lp), pa);
    c.S.dep(lp, dxl), c.S.dep(lp, ln);
    const Geo gf = make_geo(g2.k, g2.dil, g2.op, DIR_FWD);
    ta.ws_t = 1;
    conv_weight_strides(g2.op, 32, g2.k * g2.k, DIR_DGRAD, &ta.ws_k, &ta.ws_n);
    ta.H = gh, ta.W = gw, ta.Ho = ep0.in_h, ta.Wo = ep0.in_w, ta.so = down ? 2 : 1;
    ta.out32 = a->grad_in[g2.src], ta.out_ld = a->grad_in_ld[g2.src];
    for (int ph = 0; ph < nph; ++ph) {
      if (down) {  // dx[2i + p] = sum_{t in p} dy[i + d_t] W_t^T: the 4-phase table in ONE launch (as the UP forward)
        ta.taps = make_geo(g2.k, g2.dil, g2.op, DIR_DGRAD).taps, ta.img_mul = 1, ta.img_add = 0;
      } else if (up) {  // mirrored taps of this output phase, as a single-phase table on the input grid
        const int t0 = gf.taps.pstart[ph], t1 = gf.taps.pstart[ph + 1];
        if (t1 == t0) continue;
        TapTable tt;
        memset(&tt, 0, sizeof(tt));
        tt.n = t1 - t0, tt.nphase = 1;
        tt.min_dy = tt.min_dx = 1000, tt.max_dy = tt.max_dx = -1000;
        for (int t = t0; t < t1; ++t) {
          const int q = t - t0, dy = -gf.taps.dy[t], dx = -gf.taps.dx[t];
          tt.dy[q] = (int8_t)dy, tt.dx[q] = (int8_t)dx, tt.widx[q] = gf.taps.widx[t], tt.phase[q] = 0;
          tt.min_dy = std::min(tt.min_dy, dy), tt.max_dy = std::max(tt.max_dy, dy);
          tt.min_dx = std::min(tt.min_dx, dx), tt.max_dx = std::max(tt.max_dx, dx);
        }
        tt.pstart[0] = 0;
        for (int q = 1; q <= 4; ++q) tt.pstart[q] = tt.n;
        ta.taps = tt, ta.img_mul = 4, ta.img_add = ph;
      } else {
        ta.taps = make_geo(g2.k, g2.dil, g2.op, DIR_DGRAD).taps, ta.img_mul = 1, ta.img_add = 0;
      }
      ta.rows_per_cta = tc_rows(gh, gw, c.B, ta.taps.max_dy - ta.taps.min_dy);
      ta.row_chunks = cdiv(gh, ta.rows_per_cta);
      ta.accumulate = c.touched[g2.src];
      SENAS_TAG("conv_tc_dgrad", 2.0 * npix * ta.taps.n * 32 * 8 * g2.nterms, 2.0 * npix * 32 + 8.0 * npix * 32);
      const int rc = launch_conv_tc(dyb, c.B, ta, c.S.stream(dxl));
      if (rc) SENAS_FAIL("tcgen05 dgrad launch failed (code %d)", rc);
      c.touched[g2.src] = true;
    }
    // weight gradients of the group from the same packed dy (pixels = GEMM-K)
    if (!a->skip_wgrad) {
      float *dst[kTcMaxTerms] = {nullptr, nullptr, nullptr, nullptr};
      for (int i = 0; i < g2.nterms; ++i) dst[i] = a->grad_params + d.edge[g2.edge[i]].grad_off[g2.cand[i]][0];
      int ws_ci, ws_co;
      conv_weight_strides(g2.op, 32, g2.k * g2.k, DIR_FWD, &ws_ci, &ws_co);
      const __nv_bfloat16 *xb = reinterpret_cast<const __nv_bfloat16 *>(c.saved + p->xb_off[g2.src]);
      TapTable pt[4];
      if (down) down_phase_tables(gf.taps, pt);
      for (int ph = 0; ph < (down ? 4 : nph); ++ph) {
        if (down) {  // dW_t = sum_o x_ph(t)[o + d_t]^T dy[o]: x rows from phase image 4n + ph
          if (pt[ph].n == 0) continue;
          const int rcw = launch_conv_tc_wgrad(xb, dyb, c.B, gh, gw, pt[ph], 0, pt[ph].n, 1, 0, c.tmp(ln), dst, g2.nterms,
                                               ws_ci, ws_co, st, 4, ph);
          if (rcw) SENAS_FAIL("tcgen05 wgrad launch failed (code %d)", rcw);
          continue;
        }
        const int t0 = up ? gf.taps.pstart[ph] : 0, t1 = up ? gf.taps.pstart[ph + 1] : gf.taps.n;
        if (t1 == t0) continue;
        const int rcw = launch_conv_tc_wgrad(xb, dyb, c.B, ep0.in_h, ep0.in_w, gf.taps, t0, t1, up ? 4 : 1, up ? ph : 0,
                                             c.tmp(ln), dst, g2.nterms, ws_ci, ws_co, st);
        if (rcw) SENAS_FAIL("tcgen05 wgrad launch failed (code %d)", rcw);
      }
    }
  }
#endif
  c.S.join_backward();
  for (int i = 0; i < d.n_inputs; ++i)
    if (a->grad_in[i] && !c.touched[i]) {
      if (a->grad_in_ld[i] != d.edge[0].c_in) SENAS_FAIL("backward: cannot zero a strided grad_in");
      cudaMemsetAsync(a->grad_in[i], 0, sizeof(float) * (int64_t)c.B * p->in_h[i] * p->in_w[i] * a->grad_in_ld[i],
                      (cudaStream_t)c.stream);
    }
  return check_cuda("backward");
}

// ------------------------------------------------------------------------------------------------
// C ABI: lifecycle
// ------------------------------------------------------------------------------------------------
extern "C" const char *senas_version(void) { return "senas_b200 0.2 (sm_100a)"; }
extern "C" const char *senas_last_error(void) { return g_err.c_str(); }
extern "C" int64_t senas_launch_count(void) { return g_launch_count; }
extern "C" int senas_set_slot(int slot) {
#ifndef SENAS_EMU
  g_slot = slot < 0 ? 0 : slot;
#else
  (void)slot;
#endif
  return 0;
}
// NHWC AvgPool2d(3, 2, 1, count_include_pad=False) of the down cells' preprocess0 (row f1): y [B][ceil(H/2)][ceil(W/2)][C]
extern "C" int senas_avgpool_forward(const float *x, int64_t x_ld, float *y, int32_t B, int32_t H, int32_t W, int32_t C,
                                     void *stream) {
  if (!x || !y || B < 1 || H < 1 || W < 1 || C < 4 || (C & 3) || (x_ld & 3)) SENAS_FAIL("avgpool forward: bad arguments");
  const int64_t total = (int64_t)B * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4);
  SENAS_TAG("stock_avgpool", 0, 4.0 * B * H * W * C * 1.25);
  SENAS_LAUNCH(avgpool_fwd_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, x, x_ld, y, B, H, W, C);
  return check_cuda("avgpool forward");
}
extern "C" int senas_avgpool_backward(const float *gy, float *gx, int32_t B, int32_t H, int32_t W, int32_t C, void *stream) {
  if (!gy || !gx || B < 1 || H < 1 || W < 1 || C < 4 || (C & 3)) SENAS_FAIL("avgpool backward: bad arguments");
  const int64_t total = (int64_t)B * H * W * (C / 4);
  SENAS_TAG("stock_avgpool", 0, 4.0 * B * H * W * C * 1.25);
  SENAS_LAUNCH(avgpool_bwd_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, gy, gx, B, H, W, C);
  return check_cuda("avgpool backward");
}
// ------------------------------------------------------------------------------------------------
// C ABI: SURVEY rows f4 / f3 (optim.cuh) -- fused clip + SGD, Adam over flat buffers; gamma mix into the concat buffer
// ------------------------------------------------------------------------------------------------
extern "C" int senas_sgd_clip_step(float *param, float *grad, float *momentum, int64_t n, const float *lr_dev,
                                   float mom, float wd, float max_norm, float *scratch, float *norm_out, void *stream) {
  if (!param || !grad || !momentum || !lr_dev || n < 1) SENAS_FAIL("sgd step: bad arguments");
  if (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)momentum) & 15) SENAS_FAIL("sgd step: buffers must be 16-byte aligned");
  if (max_norm > 0.f && !scratch) SENAS_FAIL("sgd step: clipping needs a scratch buffer of %d floats", kOptBlocks);
  const int blocks = (int)std::min<int64_t>(kOptBlocks, ((n >> 2) + 255) / 256 + 1);
  if (max_norm > 0.f) {
    SENAS_TAG("opt_sqnorm", 2.0 * n, 4.0 * n);
    SENAS_LAUNCH(opt_sqnorm_kernel, dim3(blocks), dim3(256), 0, stream, (const float *)grad, n, scratch);
  }
  SENAS_TAG("opt_sgd", 6.0 * n, 24.0 * n);
  SENAS_LAUNCH(opt_sgd_kernel, dim3(blocks), dim3(256), 0, stream, param, grad, momentum, n, lr_dev, mom, wd, max_norm,
               (const float *)scratch, blocks, norm_out);
  return check_cuda("sgd step");
}
extern "C" int senas_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float *step, int64_t n,
                               const float *lr_dev, float beta1, float beta2, float eps, float wd, void *stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step || !lr_dev || n < 1) SENAS_FAIL("adam step: bad arguments");
  SENAS_TAG("opt_adam", 12.0 * n, 28.0 * n);
  SENAS_LAUNCH(opt_adam_kernel, dim3(1), dim3(256), 0, stream, param, grad, exp_avg, exp_avg_sq, step, n, lr_dev, beta1, beta2,
               eps, wd);
  return check_cuda("adam step");
}
extern "C" int senas_dice_ce_forward(const float *logits, const int64_t *target, int32_t B, int32_t C, int64_t HW, int64_t sn,
                                     int64_t sc, int64_t sp, float ce_scale, float smooth, float *loss, float *coef,
                                     float *scratch, void *stream) {
  if (!logits || !target || !loss || !coef || !scratch || B < 1 || HW < 1 || C < 1 || C > kLossMaxC)
    SENAS_FAIL("dice_ce forward: bad arguments (1 <= classes <= %d)", kLossMaxC);
  LossArgs a;
  memset(&a, 0, sizeof(a));
  a.logits = logits, a.target = target, a.sn = sn, a.sc = sc, a.sp = sp, a.HW = HW, a.B = B, a.C = C, a.partials = scratch;
  const int blocks = (int)std::min<int64_t>(kLossBlocks, ((int64_t)B * HW + 255) / 256);
  SENAS_TAG("loss_fwd", 0, (double)B * HW * (4.0 * C + 8.0));
  SENAS_LAUNCH(loss_fwd_kernel, dim3(blocks), dim3(256), 0, stream, a);
  SENAS_TAG("reduce", 0, 0);
  SENAS_LAUNCH(loss_final_kernel, dim3(1), dim3(32), 0, stream, (const float *)scratch, blocks, C, (float)((double)B * HW), ce_scale,
               smooth, loss, coef);
  return check_cuda("dice_ce forward");
}
extern "C" int senas_dice_ce_backward(const float *logits, const int64_t *target, int32_t B, int32_t C, int64_t HW, int64_t sn,
                                      int64_t sc, int64_t sp, const float *coef, const float *grad_loss, float *grad_logits,
                                      void *stream) {
  if (!logits || !target || !coef || !grad_loss || !grad_logits || B < 1 || HW < 1 || C < 1 || C > kLossMaxC)
    SENAS_FAIL("dice_ce backward: bad arguments");
  LossArgs a;
  memset(&a, 0, sizeof(a));
  a.logits = logits, a.target = target, a.sn = sn, a.sc = sc, a.sp = sp, a.HW = HW, a.B = B, a.C = C;
  a.coef = coef, a.gout = grad_loss, a.dlogits = grad_logits;
  const int blocks = (int)std::min<int64_t>(4 * kLossBlocks, ((int64_t)B * HW + 255) / 256);
  SENAS_TAG("loss_bwd", 0, (double)B * HW * (8.0 * C + 8.0));
  SENAS_LAUNCH(loss_bwd_kernel, dim3(blocks), dim3(256), 0, stream, a);
  return check_cuda("dice_ce backward");
}
extern "C" int senas_mix_forward(const float *a, int64_t a_ld, const float *b, int64_t b_ld, const float *w, float *out,
                                 int64_t out_ld, int32_t c0, int32_t C, int64_t npix, void *stream) {
  if (!a || !out || (b && !w) || C < 4 || (C & 3) || (c0 & 3) || (a_ld & 3) || (b_ld & 3) || (out_ld & 3) || npix < 1)
    SENAS_FAIL("mix forward: bad arguments");
  MixArgs q;
  q.a = a, q.b = b, q.w = w, q.out = out, q.npix = npix, q.a_ld = a_ld, q.b_ld = b_ld, q.out_ld = out_ld, q.C = C, q.c0 = c0;
  const int64_t total = npix * (C / 4);
  SENAS_TAG("mix_fwd", 0, 4.0 * npix * C * (b ? 3 : 2));
  SENAS_LAUNCH(mix_fwd_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, q);
  return check_cuda("mix forward");
}
extern "C" int senas_mix_backward(const float *a, int64_t a_ld, const float *b, int64_t b_ld, const float *w, const float *g,
                                  int64_t g_ld, int32_t c0, int32_t C, int64_t npix, float *da, float *db, float *dw,
                                  float *scratch, void *stream) {
  if (!a || !g || !w || !dw || !scratch || C < 4 || (C & 3) || (c0 & 3) || (a_ld & 3) || (b_ld & 3) || (g_ld & 3) || npix < 1)
    SENAS_FAIL("mix backward: bad arguments");
  MixBwdArgs q;
  q.a = a, q.b = b, q.w = w, q.g = g, q.da = da, q.db = db, q.partials = scratch;
  q.npix = npix, q.a_ld = a_ld, q.b_ld = b_ld, q.g_ld = g_ld, q.C = C, q.c0 = c0;
  const int64_t total = npix * (C / 4);
  const int blocks = (int)std::min<int64_t>(kOptBlocks * 2, (total + 255) / 256);
  SENAS_TAG("mix_bwd", 0, 4.0 * npix * C * (b ? 5 : 3));
  SENAS_LAUNCH(mix_bwd_kernel, dim3(blocks), dim3(256), 0, stream, q);
  SENAS_TAG("reduce", 0, 0);
  SENAS_LAUNCH(mix_bwd_final_kernel, dim3(1), dim3(32), 0, stream, (const float *)scratch, blocks, dw);
  return check_cuda("mix backward");
}
// ------------------------------------------------------------------------------------------------
// C ABI: SURVEY row f1 (convbn.cuh) -- ShrinkBlock / RectifyBlock: [ReLU ->] Conv2d 3x3 (c_in -> 32) -> BatchNorm2d
// ------------------------------------------------------------------------------------------------
extern "C" int senas_convbn_workspace(int32_t B, int32_t H, int32_t W, int32_t c_in, int64_t *saved_bytes,
                                      int64_t *scratch_bytes) {
#ifdef SENAS_EMU
  (void)B, (void)H, (void)W, (void)c_in, (void)saved_bytes, (void)scratch_bytes;
  SENAS_FAIL("convbn: tcgen05 path, not in the emulator build");
#else
  CbnGeo g;
  if (cbn_geo(B, H, W, c_in, &g)) SENAS_FAIL("convbn: unsupported geometry B=%d H=%d W=%d c_in=%d (W %% 64 == 0, c_in in {24,32,64,96,128})", B, H, W, c_in);
  if (saved_bytes) *saved_bytes = g.saved_floats * 4;
  if (scratch_bytes) *scratch_bytes = g.scratch_floats * 4;
  return 0;
#endif
}
extern "C" int senas_convbn_forward(const senas_convbn_args_t *a) {
#ifdef SENAS_EMU
  (void)a;
  SENAS_FAIL("convbn: tcgen05 path, not in the emulator build");
#else
  if (!a || !a->x || !a->weight || !a->gamma || !a->beta || !a->out || !a->saved || !a->scratch) SENAS_FAIL("convbn forward: null argument");
  CbnGeo g;
  if (cbn_geo(a->batch, a->h, a->w, a->c_in, &g)) SENAS_FAIL("convbn forward: unsupported geometry");
  if ((a->x_ld & 3) || ((uintptr_t)a->x & 15) || ((uintptr_t)a->out & 15) || ((uintptr_t)a->saved & 15) || ((uintptr_t)a->scratch & 15))
    SENAS_FAIL("convbn forward: alignment");
  if (!a->training && (!a->running_mean || !a->running_var)) SENAS_FAIL("convbn forward: eval mode needs running statistics");
  float *saved = (float *)a->saved, *scratch = (float *)a->scratch;
  float *y = saved + g.y_off, *stats = saved + g.stats_off;
  __nv_bfloat16 *xb = reinterpret_cast<__nv_bfloat16 *>(saved + g.xb_off);
  void *st = a->stream;
  SENAS_TAG("cbn_cast", 0, g.npix * (4.0 * a->c_in + 64.0 * g.nslices));
  SENAS_LAUNCH(cbn_cast_kernel, dim3((unsigned)((g.npix * g.nslices * 4 + 255) / 256)), dim3(256), 0, st, a->x, a->x_ld, a->c_in,
               a->relu_in, xb, g.npix, g.nslices);
  const Geo geo = make_geo(3, 1, SENAS_OP_NORM, DIR_FWD);
  const int T = 9, s_ci = T, s_co = a->c_in * T;
  for (int sl = 0; sl < g.nslices; ++sl) {
    TcConvArgs ta;
    memset(&ta, 0, sizeof(ta));
    ta.nterms = 4, ta.mode = 0;
    for (int t = 0; t < 4; ++t) {
      ta.w[t] = a->weight + (int64_t)t * 8 * s_co + (int64_t)sl * 32 * s_ci;
      ta.y[t] = y + 8 * t, ta.partials[t] = scratch + g.part_off + (int64_t)t * a->batch * g.ctas * 16;
    }
    ta.y_ld = 32, ta.ws_t = 1, ta.ws_k = s_ci, ta.ws_n = s_co;
    ta.cin_valid = std::min(32, a->c_in - 32 * sl);
    ta.H = a->h, ta.W = a->w, ta.Ho = a->h, ta.Wo = a->w, ta.so = 1;
    ta.rows_per_cta = g.rows, ta.row_chunks = g.chunks, ta.taps = geo.taps;
    ta.acc_y = sl > 0, ta.no_stats = sl + 1 < g.nslices;
    SENAS_TAG("convbn_fwd", 2.0 * g.npix * T * ta.cin_valid * 32, g.npix * (64.0 + 128.0 * (sl > 0 ? 2 : 1)));
    const int rc = launch_conv_tc(xb + (int64_t)sl * g.npix * 32, a->batch, ta, st);
    if (rc) SENAS_FAIL("convbn forward: tcgen05 launch failed (code %d)", rc);
  }
  const float M = (float)g.npix;
  if (a->training) {
    SENAS_TAG("reduce", 0, 0);
    SENAS_LAUNCH(rows_reduce_kernel, dim3(2, 4), dim3(kRowsReduceThreads), 0, st, (const float *)(scratch + g.part_off),
                 scratch + g.sums_off, a->batch * g.ctas, 16);
  }
  SENAS_TAG("bn_finalize", 0, 0);
  SENAS_LAUNCH(cbn_finalize_kernel, dim3(1), dim3(32), 0, st, (const float *)(scratch + g.sums_off), M, a->gamma, a->beta,
               a->running_mean, a->running_var, a->num_batches_tracked, a->momentum, a->eps, a->training, stats);
  SENAS_TAG("convbn_apply", 0, g.npix * 256.0);
  SENAS_LAUNCH(cbn_apply_kernel, dim3((unsigned)((g.npix * 8 + 255) / 256)), dim3(256), 0, st, (const float *)y, (const float *)stats,
               a->out, g.npix);
  return check_cuda("convbn forward");
#endif
}
extern "C" int senas_convbn_backward(const senas_convbn_args_t *a) {
#ifdef SENAS_EMU
  (void)a;
  SENAS_FAIL("convbn: tcgen05 path, not in the emulator build");
#else
  if (!a || !a->x || !a->weight || !a->gamma || !a->grad_out || !a->saved || !a->scratch || !a->grad_gamma || !a->grad_beta)
    SENAS_FAIL("convbn backward: null argument");
  CbnGeo g;
  if (cbn_geo(a->batch, a->h, a->w, a->c_in, &g)) SENAS_FAIL("convbn backward: unsupported geometry");
  if ((a->grad_out_ld & 3) || ((uintptr_t)a->grad_out & 15) || (a->grad_x && ((uintptr_t)a->grad_x & 15))) SENAS_FAIL("convbn backward: alignment");
  float *saved = (float *)a->saved, *scratch = (float *)a->scratch;
  const float *y = saved + g.y_off, *stats = saved + g.stats_off;
  const __nv_bfloat16 *xb = reinterpret_cast<const __nv_bfloat16 *>(saved + g.xb_off);
  __nv_bfloat16 *dyb = reinterpret_cast<__nv_bfloat16 *>(scratch + g.dy_off);
  float *sums = scratch + g.sums_off, *coef = scratch + g.coef_off, *bpart = scratch + g.bpart_off;
  void *st = a->stream;
  SENAS_TAG("convbn_bstats", 0, g.npix * 256.0);
  SENAS_LAUNCH(cbn_bwd_stats_kernel, dim3(g.bwd_blocks), dim3(256), 0, st, a->grad_out, a->grad_out_ld, y, stats, g.npix, g.bwd_px,
               bpart);
  SENAS_TAG("reduce", 0, 0);
  SENAS_LAUNCH(rows_reduce_kernel, dim3(8, 1), dim3(kRowsReduceThreads), 0, st, (const float *)bpart, sums, g.bwd_blocks, 64);
  SENAS_TAG("bn_finalize", 0, 0);
  SENAS_LAUNCH(cbn_bwd_finalize_kernel, dim3(1), dim3(32), 0, st, (const float *)sums, (float)g.npix, stats, a->training,
               a->grad_gamma, a->grad_beta, coef);
  SENAS_TAG("pack_dy", 0, g.npix * (256.0 + 64.0));
  SENAS_LAUNCH(cbn_pack_dy_kernel, dim3((unsigned)((g.npix * 4 + 255) / 256)), dim3(256), 0, st, a->grad_out, a->grad_out_ld, y,
               (const float *)coef, dyb, g.npix);
  const int T = 9, s_ci = T, s_co = a->c_in * T;
  const Geo gf = make_geo(3, 1, SENAS_OP_NORM, DIR_FWD), gd = make_geo(3, 1, SENAS_OP_NORM, DIR_DGRAD);
  for (int sl = 0; sl < g.nslices; ++sl) {
    const int cv = std::min(32, a->c_in - 32 * sl);
    if (a->grad_x) {
      TcConvArgs ta;
      memset(&ta, 0, sizeof(ta));
      ta.nterms = 4, ta.mode = 1;
      for (int t = 0; t < 4; ++t) ta.w[t] = a->weight + (int64_t)t * 8 * s_co + (int64_t)sl * 32 * s_ci;
      ta.ws_t = 1, ta.ws_k = s_co, ta.ws_n = s_ci, ta.cin_valid = cv;
      ta.H = a->h, ta.W = a->w, ta.Ho = a->h, ta.Wo = a->w, ta.so = 1;
      ta.out32 = a->grad_x + 32 * sl, ta.out_ld = a->c_in, ta.accumulate = 0;
      ta.mask = a->relu_in ? a->x + 32 * sl : nullptr;
      if (a->relu_in && a->x_ld != a->c_in) SENAS_FAIL("convbn backward: the ReLU mask needs x dense (x_ld == c_in)");
      ta.rows_per_cta = g.rows, ta.row_chunks = g.chunks, ta.taps = gd.taps;
      SENAS_TAG("convbn_dgrad", 2.0 * g.npix * T * cv * 32, g.npix * (64.0 + 4.0 * cv));
      const int rc = launch_conv_tc(dyb, a->batch, ta, st);
      if (rc) SENAS_FAIL("convbn backward: tcgen05 dgrad launch failed (code %d)", rc);
    }
    if (!a->grad_weight) continue;  // (architecture step: no weight gradient wanted)
    float *dst[kTcMaxTerms];
    for (int t = 0; t < 4; ++t) dst[t] = a->grad_weight + (int64_t)t * 8 * s_co + (int64_t)sl * 32 * s_ci;
    const int rcw = launch_conv_tc_wgrad(xb + (int64_t)sl * g.npix * 32, dyb, a->batch, a->h, a->w, gf.taps, 0, T, 1, 0,
                                         scratch + g.wpart_off, dst, 4, s_ci, s_co, st, 1, 0, cv);
    if (rcw) SENAS_FAIL("convbn backward: tcgen05 wgrad launch failed (code %d)", rcw);
  }
  return check_cuda("convbn backward");
#endif
}
extern "C" int senas_mix_dx(const float *g, int64_t g_ld, const float *w0, int32_t off0, const float *w1, int32_t off1,
                            const float *w2, int32_t off2, float *out, int32_t C, int64_t npix, void *stream) {
  // term j is present iff off_j >= 0; a present term with a NULL weight has coefficient 1
  if (!g || !out || C < 4 || (C & 3) || (g_ld & 3) || npix < 1) SENAS_FAIL("mix dx: bad arguments");
  MixDxArgs q;
  q.g = g, q.g_ld = g_ld, q.out = out, q.C = C, q.npix = npix;
  const float *w[3] = {w0, w1, w2};
  const int32_t off[3] = {off0, off1, off2};
  int n = 0;
  for (int j = 0; j < 3; ++j) {
    q.w[j] = w[j], q.off[j] = off[j] < 0 ? 0 : off[j], q.on[j] = off[j] >= 0;
    if (off[j] >= 0 && (off[j] & 3)) SENAS_FAIL("mix dx: channel offsets must be multiples of 4");
    n += q.on[j];
  }
  const int64_t total = npix * (C / 4);
  SENAS_TAG("mix_bwd", 0, 4.0 * npix * C * (n + 1));
  SENAS_LAUNCH(mix_dx_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, q);
  return check_cuda("mix dx");
}
// ------------------------------------------------------------------------------------------------
// C ABI: gradient exchange (NCCL over NVLink / NVSwitch), one process per GPU
// ------------------------------------------------------------------------------------------------
extern "C" int senas_comm_unique_id(void *id128) {
#ifdef SENAS_EMU
  (void)id128;
  SENAS_FAIL("no NCCL in the emulator build");
#else
  auto &a = senas_comm::api();
  if (!a.ok) SENAS_FAIL("comm: %s", a.why.c_str());
  if (!id128) SENAS_FAIL("comm: null id buffer");
  const int rc = a.GetUniqueId(reinterpret_cast<senas_comm::UniqueId *>(id128));
  if (rc) SENAS_FAIL("ncclGetUniqueId: %s", a.GetErrorString ? a.GetErrorString(rc) : "error");
  return 0;
#endif
}
extern "C" int senas_comm_init(const void *id128, int rank, int world, void **comm) {
#ifdef SENAS_EMU
  (void)id128, (void)rank, (void)world, (void)comm;
  SENAS_FAIL("no NCCL in the emulator build");
#else
  auto &a = senas_comm::api();
  if (!a.ok) SENAS_FAIL("comm: %s", a.why.c_str());
  if (!id128 || !comm || world < 1 || rank < 0 || rank >= world) SENAS_FAIL("comm_init: bad arguments");
  senas_comm::UniqueId id;
  memcpy(&id, id128, sizeof(id));
  senas_comm::Comm c = nullptr;
  const int rc = a.CommInitRank(&c, world, id, rank);  // binds to the CURRENT device of the calling thread
  if (rc) SENAS_FAIL("ncclCommInitRank: %s", a.GetErrorString ? a.GetErrorString(rc) : "error");
  *comm = c;
  return 0;
#endif
}
// in-place fp32 sum over all ranks, enqueued on `stream` (capturable into a CUDA graph)
extern "C" int senas_comm_allreduce(void *comm, float *buf, int64_t count, void *stream) {
#ifdef SENAS_EMU
  (void)comm, (void)buf, (void)count, (void)stream;
  SENAS_FAIL("no NCCL in the emulator build");
#else
  auto &a = senas_comm::api();
  if (!a.ok) SENAS_FAIL("comm: %s", a.why.c_str());
  if (!comm || !buf || count < 0) SENAS_FAIL("comm_allreduce: bad arguments");
  const int rc = a.AllReduce(buf, buf, (size_t)count, senas_comm::kFloat32, senas_comm::kSum, (senas_comm::Comm)comm,
                             (cudaStream_t)stream);
  if (rc) SENAS_FAIL("ncclAllReduce: %s", a.GetErrorString ? a.GetErrorString(rc) : "error");
  return 0;
#endif
}
extern "C" int senas_comm_destroy(void *comm) {
#ifdef SENAS_EMU
  (void)comm;
  return 0;
#else
  auto &a = senas_comm::api();
  if (comm && a.ok) a.CommDestroy((senas_comm::Comm)comm);
  return 0;
#endif
}

extern "C" int senas_set_gather_mma(int on) {  // bf16 mode: mma.sync (1, default) or CUDA-core FMA (0) for the non-tcgen05 convs
  g_gather_mma = on != 0;
  return 0;
}
extern "C" int senas_set_ds_fused(int on) {  // affects graphs planned afterwards
  g_ds_fused = on != 0;
  return 0;
}
extern "C" int senas_set_z_bfloat(int on) {  // affects graphs planned afterwards (bf16 mode only)
  g_z_bf16 = on != 0;
  return 0;
}
extern "C" int senas_set_defer(int on) {
#ifndef SENAS_EMU
  g_defer = on != 0;
#else
  (void)on;
#endif
  return 0;
}
extern "C" int senas_flush(void *stream) {
#ifndef SENAS_EMU
  for (LaneSet *ls : g_lane_sets) {
    if (!ls->pending) continue;
    for (int i = 0; i < kLanes; ++i) {
      cudaEvent_t e = ls->ev[ls->next_ev];
      ls->next_ev = (ls->next_ev + 1) % kEventPool;
      cudaEventRecord(e, ls->s[i]);
      cudaStreamWaitEvent((cudaStream_t)stream, e, 0);
    }
    ls->pending = 0;
  }
#else
  (void)stream;
#endif
  return 0;
}
extern "C" int senas_set_lanes(int n) {
  g_lanes = n < 0 ? -1 : std::min(n, kLanes);
  return 0;
}

// per-kernel-family timing: senas_profile(1) starts recording CUDA events around every launch,
// senas_profile(0) stops; senas_profile_dump() synchronises the recorded events and returns one line
// per family: "name launches total_ms algorithmic_flops algorithmic_bytes".
extern "C" int senas_profile(int on) {
#ifndef SENAS_EMU
  if (on) {
    for (auto &r : g_prof) cudaEventDestroy(r.e0), cudaEventDestroy(r.e1);
    g_prof.clear();
  }
  g_prof_on = on != 0;
#else
  (void)on;
#endif
  return 0;
}
extern "C" int64_t senas_profile_dump(char *buf, int64_t cap) {
  std::string out;
#ifndef SENAS_EMU
  struct Acc {
    int64_t n = 0;
    double ms = 0, flops = 0, bytes = 0;
  };
  std::map<std::string, Acc> acc;
  for (auto &r : g_prof) {
    cudaEventSynchronize(r.e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    Acc &a = acc[r.name];
    a.n++, a.ms += ms, a.flops += r.flops, a.bytes += r.bytes;
  }
  char line[256];
  for (auto &kv : acc) {
    snprintf(line, sizeof(line), "%s %lld %.6f %.6e %.6e\n", kv.first.c_str(), (long long)kv.second.n, kv.second.ms,
             kv.second.flops, kv.second.bytes);
    out += line;
  }
#endif
  if (buf && cap > 0) {
    const int64_t n = std::min<int64_t>((int64_t)out.size(), cap - 1);
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (int64_t)out.size();
}

extern "C" int senas_device_check(int device) {
#ifdef SENAS_EMU
  (void)device;
  return 0;
#else
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) SENAS_FAIL("cannot query CUDA device %d", device);
  if (prop.major != 10) SENAS_FAIL("device %d is sm_%d%d; libsenas_b200 is built for sm_100a only", device, prop.major, prop.minor);
  return 0;
#endif
}

extern "C" int senas_graph_create(const senas_graph_desc_t *desc, senas_graph_t **out) {
  if (!desc || !out) SENAS_FAIL("null argument");
  const senas_graph_desc_t &d = *desc;
  if (d.n_inputs < 1 || d.n_inputs > 2) SENAS_FAIL("n_inputs must be 1 or 2");
  if (d.n_nodes < 1 || d.n_nodes > SENAS_MAX_NODES) SENAS_FAIL("n_nodes out of range");
  if (d.n_edges < 1 || d.n_edges > SENAS_MAX_EDGES) SENAS_FAIL("n_edges out of range");
  if (d.c_out != 8) SENAS_FAIL("only c_out == 8 (Cell.k = 4 with 32 channels) is supported, got %d", d.c_out);
  for (int e = 0; e < d.n_edges; ++e) {
    const senas_edge_desc_t &ed = d.edge[e];
    if (ed.c_in != 8 && ed.c_in != 32) SENAS_FAIL("edge %d: c_in %d unsupported (8 or 32)", e, ed.c_in);
    if (ed.op_type < SENAS_OP_UP || ed.op_type > SENAS_OP_NORM) SENAS_FAIL("edge %d: bad op_type", e);
    if (ed.src < 0 || ed.src >= d.n_inputs + d.n_nodes || ed.dst < 0 || ed.dst >= d.n_nodes) SENAS_FAIL("edge %d: bad src/dst", e);
    if (ed.src >= d.n_inputs && ed.src - d.n_inputs >= ed.dst) SENAS_FAIL("edge %d: node edges must go forward", e);
    if (ed.src >= d.n_inputs && (ed.op_type != SENAS_OP_NORM || ed.c_in != 8)) SENAS_FAIL("edge %d: node edges are NORM 8->8", e);
    for (int k = 0; k < SENAS_MAX_CAND; ++k) {
      const int kind = ed.kind[k];
      if ((kind == SENAS_KIND_CONV || kind == SENAS_KIND_SE_CONV || kind == SENAS_KIND_DEPSEP)) {
        if (ed.ksize[k] != 3 && ed.ksize[k] != 5) SENAS_FAIL("edge %d candidate %d: kernel size %d", e, k, ed.ksize[k]);
        if (!ed.param[k][0]) SENAS_FAIL("edge %d candidate %d: missing conv weight", e, k);
      }
      if (kind == SENAS_KIND_DEPSEP && !ed.param[k][6]) SENAS_FAIL("edge %d candidate %d: missing pointwise weight", e, k);
      if (kind == SENAS_KIND_AVG_POOL && ed.op_type != SENAS_OP_DOWN) SENAS_FAIL("avg_pool is a DOWN candidate");
      if (kind == SENAS_KIND_UP_SAMPLE && ed.op_type != SENAS_OP_UP) SENAS_FAIL("up_sample is an UP candidate");
    }
  }
  senas_graph *g = new senas_graph();
  g->d = d;
  *out = g;
  return 0;
}

extern "C" void senas_graph_destroy(senas_graph_t *g) {
  if (!g) return;
  for (auto &kv : g->plans) {
    Plan *p = kv.second;
    for (auto q : p->d_bnA) if (q) dev_free(q);
    for (auto q : p->d_bnB) if (q) dev_free(q);
    if (p->d_nodes) dev_free(p->d_nodes);
    delete p;
  }
  delete g;
}

extern "C" int senas_graph_plan(senas_graph_t *g, int32_t batch, const int32_t in_h[2], const int32_t in_w[2],
                                senas_plan_info_t *info) {
  Plan *p;
  if (!info) SENAS_FAIL("null info");
  if (check_common(g, batch, in_h, in_w, &p)) return 1;
  info->out_h = p->out_h, info->out_w = p->out_w;
  info->saved_bytes = p->saved_floats * 4 + 64, info->scratch_bytes = p->scratch_floats * 4 + 64;
  return 0;
}
