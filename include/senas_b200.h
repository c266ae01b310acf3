/*
 * senas_b200.h -- C ABI of libsenas_b200.so: the SENAS supernet-search hot path on B200 (sm_100a).
 *
 * The reference (RayburnChen/senas) has no FFI: its hot path sits behind two Python methods,
 *     search.cell.MixedOp.forward(x, alpha_normal, alpha_up_dn)            search/cell.py:32-43
 *     search.cell.Cell.forward(in0, in1, weights_norm, weights_chg, betas) search/cell.py:92-110
 * (node loop :95-108 and concat :110; preprocess0/1 and post_process are outside this ABI).
 * The entry points below are what a ctypes binding of those two methods calls; see INTEGRATION.md
 * for the reference-side stub.  One "edge graph" object covers both granularities:
 *     a MixedOp  = 1 input state, 1 node, 1 edge, no beta, no ReLU;
 *     a Cell     = 2 input states, meta_node_num nodes, 2+3+..+(n+1) edges, beta-weighted node
 *                  sums, ReLU per node, nodes written side by side into one NHWC concat buffer.
 *
 * Conventions
 *   - plain pointers and sizes only; every tensor is allocated by the caller (PyTorch's caching
 *     allocator) and the library keeps no pointer past a call, except the parameter pointers
 *     recorded in the graph descriptor (they must stay valid and fixed for the graph's lifetime;
 *     optimizers update parameters in place, so they are).
 *   - activations are NHWC ("channels_last"), fp32; `*_ld` is the distance in elements between
 *     consecutive pixels (so a channel slice of a wider buffer can be passed without a copy).
 *   - every function returns 0 on success, non-zero on failure with a message available from
 *     senas_last_error() (thread local).  There is no CPU fallback and no multi-backend dispatch:
 *     a non-sm_100 device or an unsupported shape is an error.
 *   - all work is enqueued on the caller's stream; nothing synchronises the device.
 */
#ifndef SENAS_B200_H_
#define SENAS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SENAS_MAX_CAND 6   /* candidates per MixedOp = alpha columns (utils/operations.py:23-48) */
#define SENAS_SLOTS 12     /* parameter/buffer slots per candidate, see table below */
#define SENAS_MAX_EDGES 16
#define SENAS_MAX_NODES 3 /* (a 4th node would have 5 incoming edges = 30 candidate terms; the search configs use 3) */
#define SENAS_FLAG_TC_BF16 1

/* OpType ids: the reference's OpType.value['id'] (utils/operations.py:51-54) */
enum { SENAS_OP_UP = 1, SENAS_OP_DOWN = 2, SENAS_OP_NORM = 3 };

/* candidate kinds (OPS registry, utils/operations.py:8-21) */
enum {
  SENAS_KIND_NONE = 0,      /* ZeroOp  -> (1x1 conv) -> BN          :9,155-164  */
  SENAS_KIND_IDENTITY = 1,  /* Identity-> (1x1 conv) -> BN          :10         */
  SENAS_KIND_AVG_POOL = 2,  /* AvgPool2d(3,s,1,count_include_pad=False) -> 1x1 -> BN  :62 */
  SENAS_KIND_UP_SAMPLE = 3, /* Upsample(x2, bilinear, align_corners=False) -> 1x1 -> BN :13 */
  SENAS_KIND_CONV = 4,      /* ConvBn   (dil_3_conv_5, dil_2_conv_5) :89-95     */
  SENAS_KIND_SE_CONV = 5,   /* ConvBnSe (se_conv_3)                  :98-104    */
  SENAS_KIND_DEPSEP = 6     /* DepSepConv (dep_sep_conv_3/5)         :107-115   */
};

/*
 * Parameter slots (device pointers to fp32 unless noted; NULL when absent).
 *   BN group at slot s: s+0 weight(gamma) s+1 bias(beta) s+2 running_mean s+3 running_var
 *                       s+4 num_batches_tracked (int64)
 *   NONE/IDENTITY/AVG_POOL/UP_SAMPLE : 0 conv.weight [8,c_in,1,1] (NULL when c_in == 8), BN @1
 *   CONV                             : 0 conv weight ([8,c_in,k,k], or [c_in,8,k,k] for UP), BN @1
 *   SE_CONV                          : as CONV, plus 6 excitation.0.weight [1,8], 7 excitation.2.weight [8,1]
 *   DEPSEP                           : 0 depthwise weight [c_in,1,k,k], BN(c_in) @1,
 *                                      6 pointwise weight [8,c_in,1,1], BN(8) @7
 */
typedef struct {
  int32_t src;                 /* state read by the edge: 0..n_inputs-1 inputs, n_inputs+i = node i */
  int32_t dst;                 /* node that accumulates it */
  int32_t op_type;             /* SENAS_OP_* */
  int32_t c_in;                /* 8 or 32 */
  int32_t kind[SENAS_MAX_CAND];
  int32_t ksize[SENAS_MAX_CAND];    /* 3 or 5 for CONV/SE_CONV/DEPSEP, else 0 */
  int32_t dilation[SENAS_MAX_CAND]; /* 1, 2 or 3 */
  void *param[SENAS_MAX_CAND][SENAS_SLOTS];
  int64_t grad_off[SENAS_MAX_CAND][SENAS_SLOTS]; /* offset (floats) in the flat gradient buffer, -1 = none */
} senas_edge_desc_t;

typedef struct {
  int32_t n_inputs;            /* 1 or 2 */
  int32_t n_nodes;             /* 1..SENAS_MAX_NODES */
  int32_t n_edges;
  int32_t c_out;               /* channels per node; only 8 is supported (Cell.k = 4, c = 32) */
  int32_t node_relu;           /* 1: node = relu(sum), cell.py:107; 0: plain sum (MixedOp) */
  int32_t reserved;            /* flags: bit 0 = SENAS_FLAG_TC_BF16 (tcgen05 convs with bf16 operands, fp32 accumulation) */
  int64_t grad_floats;         /* length of the flat parameter-gradient buffer */
  senas_edge_desc_t edge[SENAS_MAX_EDGES];
} senas_graph_desc_t;

typedef struct senas_graph senas_graph_t;

typedef struct {
  int32_t out_h, out_w;        /* spatial size of the nodes */
  int64_t saved_bytes;         /* forward -> backward buffer (pre-BN candidate outputs, statistics) */
  int64_t scratch_bytes;       /* scratch, may be shared by all graphs of one stream */
} senas_plan_info_t;

typedef struct {
  int32_t batch;
  int32_t training;            /* 1: batch statistics + running-stat update; 0: running statistics */
  int32_t in_h[2], in_w[2];
  const float *in[2];          /* input states, NHWC */
  int64_t in_ld[2];
  const float *alpha;          /* [n_edges][6] softmaxed weights, row already selected by OpType (cell.py:33-36) */
  const float *beta;           /* [n_edges] per-edge weights (cell.py:104) or NULL for 1 */
  float *out;                  /* [B, out_h, out_w, n_nodes*c_out] = cat(states[-n:], 1) (cell.py:110) */
  int64_t out_ld;
  void *saved;
  void *scratch;
  void *stream;                /* cudaStream_t */
} senas_fwd_args_t;

typedef struct {
  int32_t batch;
  int32_t training;
  int32_t in_h[2], in_w[2];
  const float *in[2];
  int64_t in_ld[2];
  const float *alpha;
  const float *beta;
  const float *out;            /* forward result (ReLU mask) */
  int64_t out_ld;
  const float *grad_out;       /* dL/d out, same geometry */
  int64_t grad_out_ld;
  void *saved;                 /* the buffer forward filled; backward may overwrite it */
  void *scratch;
  float *grad_in[2];           /* written (not accumulated); NULL to skip an input */
  int64_t grad_in_ld[2];
  float *grad_alpha;           /* [n_edges][6] */
  float *grad_beta;            /* [n_edges] or NULL */
  float *grad_params;          /* flat buffer of desc.grad_floats floats, fully written */
  void *stream;
  int32_t skip_wgrad;          /* != 0: the caller wants data / alpha / beta gradients only (the architecture step of
                                * search_arc.py:268-271 discards every weight gradient: model_optimizer.zero_grad() follows):
                                * no convolution / depthwise / 1x1 weight-gradient kernel is launched and grad_params is
                                * left UNDEFINED */
  int32_t reserved_;
} senas_bwd_args_t;

const char *senas_version(void);
const char *senas_last_error(void);
/* 0 iff `device` is a compute-capability 10.x GPU this library was built for */
int senas_device_check(int device);

int senas_graph_create(const senas_graph_desc_t *desc, senas_graph_t **out);
void senas_graph_destroy(senas_graph_t *g);
/* geometry + workspace sizes for one batch/input size (cached inside the graph) */
int senas_graph_plan(senas_graph_t *g, int32_t batch, const int32_t in_h[2], const int32_t in_w[2],
                     senas_plan_info_t *info);
int senas_graph_forward(senas_graph_t *g, const senas_fwd_args_t *a);
int senas_graph_backward(senas_graph_t *g, const senas_bwd_args_t *a);

/* Row f1 of the scope table ("next"): AvgPool2d(3, stride 2, padding 1, count_include_pad=False) of the down cells'
 * preprocess0 (utils/operations.py:141-152, search/cell.py:57-60) on NHWC fp32.  x: [B][H][W][x_ld >= C], y and gy:
 * [B][ceil(H/2)][ceil(W/2)][C] dense, gx: [B][H][W][C] dense (written, not accumulated). */
int senas_avgpool_forward(const float *x, int64_t x_ld, float *y, int32_t B, int32_t H, int32_t W, int32_t C, void *stream);
int senas_avgpool_backward(const float *gy, float *gx, int32_t B, int32_t H, int32_t W, int32_t C, void *stream);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
int64_t senas_launch_count(void);
/* number of side streams ("lanes") over which independent candidate chains of a call are spread (fork/join with
 * events on the caller's stream; a captured step becomes a DAG).  0 = strictly serial on the caller's stream,
 * negative = default (environment variable SENAS_LANES, else 16).  Results are bit-identical for every setting. */
int senas_set_lanes(int n);
/* lane set used by the following calls of this process (default 0).  A host that runs independent graphs concurrently
 * on different streams gives each stream its own slot (and its own scratch buffer) so that their lanes do not
 * serialise against each other; one host thread drives the library, so the selection is process-global. */
int senas_set_slot(int slot);
/* Deferred join of the weight-gradient lanes.  With senas_set_defer(1), senas_graph_backward returns (in stream order)
 * as soon as the data gradients, grad_alpha and grad_beta are complete; grad_params and every buffer the call read
 * (inputs, saved, scratch of the slot) must then stay untouched and alive until the next call in the same slot or until
 * senas_flush(stream), which makes `stream` wait for all such pending work of the process.  Default 0: backward
 * returns with everything ordered on the caller's stream. */
int senas_set_defer(int on);
int senas_flush(void *stream);
/* Experimental: dep-sep candidates of NORM edges through the recompute kernels (ds_norm_kernel: the depthwise output is
 * recomputed from the input in every sweep and never stored).  Off by default (measured slower than the spill path on
 * B200); applies to graphs planned after the call.  Environment variable SENAS_DS_FUSED sets the default. */
int senas_set_ds_fused(int on);
/* bf16 mode: store the depthwise output z of the dep-sep chains (and the gradient dz written over it) as bf16: half the
 * traffic of the chain's six sweeps; statistics, ReLU mask and all consumers read the same rounded values.  Affects
 * graphs planned afterwards. */
int senas_set_z_bfloat(int on);

/* Data-parallel gradient exchange, one process per GPU (replaces the reference's in-process replica path,
 * search/senas_search.py:262-279 and utils/utils.py:233-237, broken as shipped).  NCCL over NVLink 5 / NVSwitch, bound
 * with dlopen at the first call (no link-time dependency).  senas_comm_unique_id fills 128 bytes on one rank; the host
 * distributes them (any channel) and every rank calls senas_comm_init with its CUDA device current.  The all-reduce is an
 * in-place fp32 SUM enqueued on `stream`; it may be captured into a CUDA graph together with the kernels around it. */
/* SURVEY 8f rows f4 / f3 (the blocks between the cells), flat fp32 device buffers, 16-byte aligned.
 * senas_sgd_clip_step: torch.nn.utils.clip_grad_norm_(max_norm, 2) + torch.optim.SGD(momentum, weight_decay).step() of
 *   experiments/search_arc.py:282-293 as two launches: grad is scaled in place by min(1, max_norm / (||grad|| + 1e-6)),
 *   momentum = mom * momentum + (grad + wd * param), param -= lr * momentum.  lr_dev points to ONE device float (a
 *   scheduler rewrites it; nothing is baked into a captured graph).  scratch: >= 296 floats (norm partials, fixed
 *   summation order); max_norm <= 0 disables clipping; norm_out (optional) receives the unclipped norm.
 * senas_adam_step: torch.optim.Adam(betas, eps, weight_decay).step() on the architecture parameters; `step` is a device
 *   float that the call increments (Adam's bias-correction counter).
 * senas_mix_forward / backward: out[:, c0:c0+C] = w[0] * a + w[1] * b over npix NHWC pixels (b may be NULL: w[0] * a; then w may be NULL too: a plain copy) --
 *   the gamma-weighted skip mix of search/senas_search.py:96-107 written straight into the concat buffer that feeds the
 *   cell's ShrinkBlock; backward returns da = w[0] * g, db = w[1] * g (dense [npix][C], either may be NULL) and
 *   dw[0..1] = <g, a>, <g, b>; scratch: >= 1184 floats. */
int senas_sgd_clip_step(float *param, float *grad, float *momentum, int64_t n, const float *lr_dev, float mom, float wd,
                        float max_norm, float *scratch, float *norm_out, void *stream);
int senas_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float *step, int64_t n,
                    const float *lr_dev, float beta1, float beta2, float eps, float wd, void *stream);
/* SegmentationLosses('dice_ce') of the search configuration (utils/loss/loss.py:45-70,124-159): mean cross entropy *
 * ce_scale + soft dice over the classes >= 1 (smooth 1e-5), forward (2 launches) and backward (1 launch).  logits / grad_logits:
 * element offset n * sn + c * sc + pixel * sp (NCHW: sn = C HW, sc = HW, sp = 1; channels_last: sn = C HW, sc = 1, sp = C);
 * target [B][HW] int64; loss: one float; coef: 3 C + 1 floats kept for backward; scratch: >= 296 * 25 floats;
 * grad_loss: device scalar (the upstream gradient).  classes <= 8. */
int senas_dice_ce_forward(const float *logits, const int64_t *target, int32_t B, int32_t C, int64_t HW, int64_t sn, int64_t sc,
                          int64_t sp, float ce_scale, float smooth, float *loss, float *coef, float *scratch, void *stream);
int senas_dice_ce_backward(const float *logits, const int64_t *target, int32_t B, int32_t C, int64_t HW, int64_t sn, int64_t sc,
                           int64_t sp, const float *coef, const float *grad_loss, float *grad_logits, void *stream);
int senas_mix_forward(const float *a, int64_t a_ld, const float *b, int64_t b_ld, const float *w, float *out, int64_t out_ld,
                      int32_t c0, int32_t C, int64_t npix, void *stream);
int senas_mix_backward(const float *a, int64_t a_ld, const float *b, int64_t b_ld, const float *w, const float *g, int64_t g_ld,
                       int32_t c0, int32_t C, int64_t npix, float *da, float *db, float *dw, float *scratch, void *stream);
/* the Cell's pre / post blocks -- ShrinkBlock (utils/operations.py:206-218: ReLU -> Conv2d 3x3 c_in ->
 * 32, padding 1, no bias -> BatchNorm2d) and RectifyBlock (:221-232: the same without the ReLU, c_in = 24) as one op on the
 * tcgen05 kernels (bf16 operands, fp32 accumulation; the 2e-2 mode).  All tensors NHWC fp32; x may be a strided view
 * (x_ld = its pixel stride, e.g. a slice of / the whole concat buffer).  weight: PyTorch Conv2d layout [32][c_in][3][3].
 * w must be a multiple of 64; c_in in {24, 32, 64, 96, 128}.  saved / scratch: senas_convbn_workspace bytes; `saved` lives
 * from forward to backward (pre-BN output y, batch statistics, the bf16 copy of x), scratch is free after each call.
 * Training mode updates running_mean / running_var / num_batches_tracked in place like nn.BatchNorm2d. */
typedef struct {
  int32_t batch, h, w, c_in, relu_in, training;
  const float *x; int64_t x_ld;
  const float *weight, *gamma, *beta;
  float *running_mean, *running_var; int64_t *num_batches_tracked;
  float momentum, eps;
  float *out;                                   /* forward: [batch][h][w][32] */
  void *saved, *scratch;
  const float *grad_out; int64_t grad_out_ld;   /* backward inputs */
  float *grad_x;                                /* backward: [batch][h][w][c_in] dense, or NULL */
  float *grad_weight, *grad_gamma, *grad_beta;  /* backward: [32][c_in][3][3] (NULL: not wanted), [32], [32] */
  void *stream;
} senas_convbn_args_t;
int senas_convbn_workspace(int32_t batch, int32_t h, int32_t w, int32_t c_in, int64_t *saved_bytes, int64_t *scratch_bytes);
int senas_convbn_forward(const senas_convbn_args_t *a);
int senas_convbn_backward(const senas_convbn_args_t *a);
/* gradient of one skip tensor: out[pix][c] = sum_j coef_j * g[pix][off_j + c] over the (up to 3) channel slices of the
 * concat gradient it reached; off_j < 0: term absent; w_j == NULL: coefficient 1, else *w_j (device float). */
int senas_mix_dx(const float *g, int64_t g_ld, const float *w0, int32_t off0, const float *w1, int32_t off1, const float *w2,
                 int32_t off2, float *out, int32_t C, int64_t npix, void *stream);
int senas_comm_unique_id(void *id128);
int senas_comm_init(const void *id128, int rank, int world, void **comm);
int senas_comm_allreduce(void *comm, float *buf, int64_t count, void *stream);
int senas_comm_destroy(void *comm);
/* bf16 mode: the convolutions that are not on the tcgen05 path (8 -> 8 node edges, maps not a multiple of 64 wide) run as
 * mma.sync m16n8k8 TF32 (1; environment variable SENAS_GATHER_MMA) or as exact fp32 FMA (0, default: see DESIGN.md). */
int senas_set_gather_mma(int on);
/* per-kernel-family device timing (CUDA events on the launch stream): senas_profile(1) starts a
 * recording, senas_profile(0) stops it, senas_profile_dump() waits for the recorded events and writes
 * "family launches total_ms algorithmic_flops algorithmic_bytes" lines (returns the text length). */
int senas_profile(int on);
int64_t senas_profile_dump(char *buf, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* SENAS_B200_H_ */
