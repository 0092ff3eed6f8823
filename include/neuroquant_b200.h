/*
 * neuroquant_b200.h -- C ABI of libnq_sm100.so, the B200 (sm_100a) kernels behind the NeuroQuant
 * post-training-quantisation hot path.
 *
 * The reference (Eric-qi/NeuroQuant) has no FFI layer: its boundary for this path is the Python API
 * of quantization/ (quantizer, quant_layer, quant_block, quant_model, calib_model).  Each entry point below names the reference statement(s) it replaces
 * (paths relative to the reference root).  The library is stateless: the caller (PyTorch host code in
 * neuroquant_b200/, or any other host) owns every buffer and passes raw DEVICE pointers, explicit
 * sizes and the CUDA stream (cudaStream_t as void*).  Every function returns 0 on success or a negative
 * nq_status; nothing throws and nothing synchronises.  All tensors are fp32 unless stated.
 *
 * Layouts
 *   "ref"    weight  (C_out, C_in, KH, KW) contiguous, exactly the reference's nn.Conv2d layout
 *            (with --hadamard the quantised tensor is (C_out, C_pow2, KH, KW)).
 *   "packed" weight  Wk[(kh*KW+kw)*cin_p + ci][n']   (GEMM-K rows, GEMM-N contiguous), n' the packed
 *            output channel: n' = (i*rw + j)*cg + c  for reference channel  c*rh*rw + i*rw + j,
 *            so that the up-shuffle (nn.PixelShuffle / the stem fold) becomes a contiguous store.
 *            cin_p, cg are channel counts rounded up to a multiple of 4; pad entries are zero.
 *   "packedT" Wt[((KH-1-kh)*KW+(KW-1-kw))*nout_p + n'][ci]  the flipped transpose used by dgrad.
 *   activations      NHWC  x[n][h][w][c_p]  (channels-last, padded channels hold zeros)
 *   frames           NCHW  (B, 3, H, W) exactly as the reference's dataloader produces them.
 */
#ifndef NEUROQUANT_B200_H_
#define NEUROQUANT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum nq_status {
  NQ_OK = 0,
  NQ_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, bits outside 2..8 (quantizer.py:96) */
  NQ_ERR_BAD_SHAPE = -2,    /* padding / divisibility contract of a layout violated */
  NQ_ERR_UNSUPPORTED = -3,  /* e.g. rotation length > 256 or not a power of two */
  NQ_ERR_WORKSPACE = -4,    /* caller-provided workspace too small */
  NQ_ERR_CUDA = -5          /* a CUDA runtime call / launch failed (see nq_last_cuda_error) */
} nq_status;

const char* nq_status_string(int status);
/* cudaError_t of the most recent NQ_ERR_CUDA on the calling thread (0 if none). */
int nq_last_cuda_error(void);
/* ABI version; bumped on any signature change. */
int nq_abi_version(void);
/* Device properties the host needs for grid sizing: SM count of the current device. */
int nq_sm_count(void);

/* ------------------------------------------------------------------------------------------------
 * Quantisers (quantization/quantizer.py)
 * ------------------------------------------------------------------------------------------------ */

/* UniformAffineQuantizer.init_quantization_scale, 'max', asymmetric (quantizer.py:127-168), replacing
 * the per-channel Python loop with .item() syncs (:139-140).  x: rows x row_len (one row per output
 * channel; rows==1 for the per-tensor bias case, :144-152).  Writes delta[rows], zero_point[rows]. */
int nq_uaq_init_max(const float* x, int64_t rows, int64_t row_len, int n_bits,
                    float* delta, float* zero_point, void* stream);

/* The other asymmetric initialisers of init_quantization_scale: method 1 'mse' (ten shrinking ranges scored by the L_3.5
 * norm, quantizer.py:170-187), 2 'l1' (:204-220), 3 'gaussian' (mu +- 6 var, :189-202).  Same layout as nq_uaq_init_max. */
int nq_uaq_init_search(const float* x, int64_t rows, int64_t row_len, int n_bits, int method,
                       float* delta, float* zero_point, void* stream);

typedef enum nq_round_mode {
  NQ_ROUND_NEAREST = 0, /* UAQ forward, round-half-even + STE   (quantizer.py:117-119, :53-57) */
  NQ_ROUND_SOFT = 1,    /* AdaRound floor + h(alpha)           (quantizer.py:288-291, :302-303) */
  NQ_ROUND_HARD = 2     /* AdaRound floor + [alpha >= 0]        (quantizer.py:292-293) */
} nq_round_mode;

/* Fake-quantise x (rows x row_len, row r uses delta[r*d_stride], zero_point[r*d_stride]; d_stride 0
 * broadcasts one scale).  codes (may be NULL) receives clamp(x_int + zp, 0, 2^bits-1) -- the tensor the
 * reference caches in AdaRoundQuantizer.x_quant (quantizer.py:297); deq (may be NULL) receives
 * (codes - zp) * delta (quantizer.py:119, :298).  alpha is required for the two AdaRound modes.
 * reg_sum (may be NULL): *reg_sum += sum(1 - |2 h(alpha) - 1|^reg_b) (calib_model.py:44-45, before the
 * `weight` factor); the caller zeroes it. */
int nq_fakequant_fwd(const float* x, const float* alpha, const float* delta, const float* zero_point,
                     int64_t rows, int64_t row_len, int d_stride, int n_bits, int mode,
                     float* codes, float* deq, float* reg_sum, float reg_b, void* stream);

/* Backward of the above w.r.t. the learnable of each phase (closed forms verified against autograd of
 * quantizer.py, see tests):
 *   NQ_ROUND_NEAREST: d_delta[r] = sum_row g * ((codes - zp) - [in range] * x / delta)     (phase 1)
 *   NQ_ROUND_SOFT   : d_alpha    = g * delta * [in range] * h'(alpha)  + reg_w * d reg / d alpha (phase 2)
 * g is the gradient w.r.t. the de-quantised tensor.  reg_w = 0 disables the regulariser term
 * (calib_model.py:77-78 warm-up, and always for biases, :44).  grad_scale multiplies g first (1/G for the
 * mean over G data-parallel ranks). */
int nq_fakequant_bwd(const float* g, const float* x, const float* alpha, const float* delta,
                     const float* zero_point, int64_t rows, int64_t row_len, int d_stride, int n_bits,
                     int mode, float grad_scale, float reg_w, float reg_b,
                     float* d_alpha, float* d_delta, void* stream);

/* AdaRoundQuantizer.init_alpha (quantizer.py:305-313). */
int nq_adaround_init_alpha(const float* x, const float* delta, int64_t rows, int64_t row_len,
                           int d_stride, float* alpha, void* stream);

/* torch.optim.Adam single step (defaults: amsgrad off, no weight decay; calib_model.py:134,195) over one
 * flat tensor; step is 1-based.  Arithmetic order follows torch's single-tensor implementation. */
int nq_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                 double lr, double beta1, double beta2, double eps, int step, void* stream);

/* CUDA-graph variants of the two calls whose scalars change every iteration: hyper_dev is a device array
 * {reg_w, reg_b, lr / (1 - beta1^t), sqrt(1 - beta2^t)} that the host refreshes (one 16-byte copy) before each
 * replay of the captured iteration.  use_reg = 0 ignores the regulariser entries (bias quantisers). */
int nq_fakequant_bwd_soft_dev(const float* g, const float* x, const float* alpha, const float* delta,
                              const float* zero_point, int64_t rows, int64_t row_len, int d_stride, int n_bits,
                              float grad_scale, int use_reg, const float* hyper_dev, float* d_alpha, void* stream);
int nq_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                     double beta1, double beta2, double eps, const float* hyper_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-tensor launches.  A decoder has 2 quantisers per stage (weight, bias) of 12 .. 1.6 M elements; one
 * launch per tensor is launch-latency bound.  These entry points take a HOST array of per-tensor tasks (device
 * pointers inside) and process up to NQ_MULTI_MAX tensors per kernel launch; element arithmetic is the
 * single-tensor kernels', value for value.
 * ------------------------------------------------------------------------------------------------ */
#define NQ_MULTI_MAX 16

typedef struct nq_fq_task {    /* one nq_fakequant_fwd call */
  const float* x;
  const float* alpha;          /* may be NULL for NQ_ROUND_NEAREST */
  const float* delta;
  const float* zero_point;
  float* codes;                /* may be NULL */
  float* deq;                  /* may be NULL */
  int64_t rows, row_len;
  int32_t channel_wise;        /* 1: delta / zero_point per row, 0: one value */
  int32_t n_bits, mode;        /* nq_round_mode */
  int32_t want_reg;            /* 1: this tensor's soft targets add sum(1 - |2h - 1|^reg_b) to *reg_sum */
} nq_fq_task;
int nq_fakequant_fwd_multi(const nq_fq_task* tasks, int n_tasks, float* reg_sum, float reg_b, void* stream);

typedef struct nq_ada_task {   /* nq_fakequant_bwd_soft_dev + nq_adam_step_dev of one AdaRound quantiser */
  const float* g;              /* dLoss / d(de-quantised tensor) */
  const float* x;              /* the tensor being quantised */
  float* alpha;                /* rounding variables V: the Adam parameter, updated in place */
  const float* delta;
  const float* zero_point;
  float* exp_avg;              /* Adam state */
  float* exp_avg_sq;
  int64_t rows, row_len;
  int32_t channel_wise, n_bits;
  int32_t use_reg;             /* 1: add the rounding regulariser's gradient (weights), 0: not (biases) */
  int32_t reserved;
} nq_ada_task;
/* alpha <- Adam(alpha, d(loss + reg)/d alpha) for every task in one pass (calib_model.py:213-218: loss.backward();
 * optimizer.step()); hyper_dev as in nq_fakequant_bwd_soft_dev / nq_adam_step_dev. */
int nq_adaround_step_multi(const nq_ada_task* tasks, int n_tasks, float grad_scale, double beta1, double beta2, double eps,
                           const float* hyper_dev, void* stream);


/* ------------------------------------------------------------------------------------------------
 * Walsh-Hadamard rotation (quantization/quant_layer.py:16-22; third-party hadamard_transform)
 * ------------------------------------------------------------------------------------------------ */

/* Orthonormal Sylvester WHT of length n (power of two, <= 256) along a strided axis:
 * element k of vector v lives at base + (v / inner) * outer_stride + (v % inner) + k * inner, which
 * covers both "rows of a matrix" (inner = 1, outer_stride = n) and "the C_in axis of a ref-layout
 * weight" (inner = KH*KW, outer_stride = n*KH*KW).  src == dst is allowed. */
int nq_fwht(const float* src, float* dst, int64_t n_vectors, int n, int64_t inner,
            int64_t outer_stride, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Weight packing between the reference layout and the GEMM layouts
 * ------------------------------------------------------------------------------------------------ */
typedef struct nq_conv_desc {
  int32_t n, h, w;   /* input grid: batch, height, width */
  int32_t cin;       /* real input channels */
  int32_t cin_p;     /* padded input channels (multiple of 4) */
  int32_t ksize;     /* square, odd; stride 1, "same" zero padding (quant_layer.py:33-34) */
  int32_t cout;      /* real output channels of the conv = c_grp * rh * rw */
  int32_t rh, rw;    /* up-shuffle factors applied after the conv (1,1 = none) */
  int32_t c_grp;     /* real channels after the shuffle */
  int32_t cg;        /* padded channels after the shuffle (multiple of 4) */
  int32_t act;       /* 0 none, 1 exact-erf GELU (nn.GELU, _layers.py:105), 2 the same GELU with the `z` buffer
                        holding GELU'(pre-activation) instead of the pre-activation: what autograd's backward of
                        nn.GELU needs, evaluated once in the forward epilogue next to GELU itself */
} nq_conv_desc;
/* derived: nout_p = rh*rw*cg (GEMM N), kdim = ksize*ksize*cin_p (GEMM K) */

/* ref (cout, cin_src, k, k) -> packed Wk [kdim][nout_p] and/or packedT Wt [k*k*nout_p][cin_p]
 * (either may be NULL); bias_ref (cout) -> bias_packed [nout_p] (either NULL to skip).  cin_src >= cin is
 * the channel count of the source tensor (C_pow2 after the inverse rotation; extra channels dropped,
 * quant_layer.py:71 `[:, :self.C]`). */
int nq_pack_weight(const nq_conv_desc* d, const float* w_ref, int cin_src, const float* bias_ref,
                   float* wk, float* wt, float* bias_packed, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Decoder stages (quantization/quant_layer.py:80 F.conv2d; quant_block.py:31-35; models/HNeRV.py:49-71)
 * ------------------------------------------------------------------------------------------------ */

/* conv(k, same) + bias + up-shuffle(rh, rw) + activation, NHWC.
 *   x  (n, h, w, cin_p)   wk packed [kdim][nout_p]   bias_packed [nout_p]
 *   y  (n, h*rh, w*rw, cg) activated output;  z (same shape, may be NULL) pre-activation, kept for
 *   the backward pass.  Exact fp32 FFMA path (NQ_PREC_FP32). */
int nq_conv_fwd(const nq_conv_desc* d, const float* x, const float* wk, const float* bias_packed,
                float* z, float* y, void* stream);

/* Data gradient of the stage, fused with the previous stage's activation derivative and un-shuffle:
 *   dz      (n, h, w, nout_p)      gradient w.r.t. this stage's conv output (packed channel order)
 *   wt      packedT weights
 *   z_prev  (n, h, w, cin_p) or NULL   pre-activation of the previous stage (same grid as x); with prev_act 2
 *           the derivative GELU'(pre-activation) that the forward kernel saved under act 2
 *   dz_prev (n, h/prev_rh, w/prev_rw, prev_rh*prev_rw*cin_p): gradient w.r.t. the previous stage's conv
 *           output, i.e. (dx * act'(z_prev)) un-shuffled.  prev_act as nq_conv_desc.act. */
int nq_conv_dgrad(const nq_conv_desc* d, const float* dz, const float* wt, const float* z_prev,
                  int prev_rh, int prev_rw, int prev_act, float* dz_prev, void* stream);

/* Weight + bias gradient.  Writes dwk [(kdim + 4)][nout_p]: rows < kdim are dWk in packed layout, row
 * kdim is the bias gradient (sum of dz over pixels), rows kdim+1.. are zero.  The pixel axis is split
 * over `splits` CTAs per tile; partials go to `workspace` (>= splits*(kdim+4)*nout_p floats when
 * splits > 1) and are summed in a fixed order, so the result is run-to-run deterministic. */
int nq_conv_wgrad(const nq_conv_desc* d, const float* x, const float* dz, float* dwk,
                  float* workspace, int64_t workspace_floats, int splits, void* stream);

/* packed gradient -> ref layout: dw_ref (cout, cin_dst, k, k) with channels >= cin zero-filled (the
 * zero pad of quant_layer.py:47 before the rotation), db_ref (cout).  Either output may be NULL. */
int nq_unpack_wgrad(const nq_conv_desc* d, const float* dwk, int cin_dst, float* dw_ref, float* db_ref,
                    void* stream);

/* Head: 3x3 conv to 3 channels + OutImg + reconstruction loss (models/HNeRV.py:63-64, _layers.py:10-16,
 * quantizer.py:66-73), one pass.
 *   x (n, h, w, cin_p); w_head packed [9*cin_p][4]; bias_head [4]
 *   out_bias: 0 tanh (0.5*tanh+0.5), 1 sigmoid
 *   img   (n, 3, h, w) NCHW output frame (may be NULL)
 *   target (n, 3, h, w) or NULL (decode only).  With a target: *loss_sum += sum_c |img - tgt|^p over all
 *   pixels (caller zeroes; divide by the GLOBAL pixel count for lp_loss), and dz_head (n, h, w, 4) (may be
 *   NULL) receives d lp_loss / d(conv output) with the mean taken over `mean_pixels` pixels. */
int nq_head_fwd_loss(const nq_conv_desc* d, const float* x, const float* w_head, const float* bias_head,
                     int out_bias, const float* target, float p, float mean_pixels,
                     float* img, float* loss_sum, float* dz_head, void* stream);

/* Same with the split-bf16 tensors of the tensor-core engine: x_split (n, h, w, cin_p) planes in,
 * dz_head_split (n, h, w, 8) planes out (3 real channels + 5 zeros = one 16-byte chunk per pixel and plane). */
int nq_head_fwd_loss_split(const nq_conv_desc* d, const void* x_split, const float* w_head, const float* bias_head,
                           int out_bias, const float* target, float p, float mean_pixels,
                           float* img, float* loss_sum, void* dz_head_split, void* stream);
/* Same contract on the tensor cores, "tap-expanded": the 9 taps become GEMM columns (N = 27, K = C, no halo in the
 * activation operand) and the convolution is finished by nine shifted adds per output -- 9 MMAs per 128 input pixels
 * instead of 81 (nq_tc_head_fwd_loss) or 1080 FMAs per pixel (nq_head_fwd_loss_split).  cin_p <= 64. */
int nq_head_fwd_loss_tapexp(const nq_conv_desc* d, const void* x_split, const float* w_head, const float* bias_head,
                            int out_bias, const float* target, float p, float mean_pixels, float* img,
                            float* loss_sum, void* dz_head_split, void* stream);
/* Same, with the target frames as the data set stores them: uint8 (n, 3, h, w); the kernel evaluates value / 255 in fp32
 * (IEEE division: bit-identical to the reference's `read_image(...) / 255.0`, videosets/datasets.py:8-54, consumed at
 * calib_model.py:150).  A quarter of the bytes over PCIe and out of HBM per iteration. */
int nq_head_fwd_loss_tapexp_u8(const nq_conv_desc* d, const void* x_split, const float* w_head, const float* bias_head,
                               int out_bias, const uint8_t* target_u8, float p, float mean_pixels,
                               float* img, float* loss_sum, void* dz_head_split, void* stream);

/* Head weight/bias gradient: dwk_head [(9*cin_p + 4)][4] as nq_conv_wgrad; workspace >= blocks*(9*cin_p+4)*4
 * floats with blocks = nq_head_wgrad_blocks(d). */
int nq_head_wgrad_blocks(const nq_conv_desc* d);
int nq_head_wgrad(const nq_conv_desc* d, const float* x, const float* dz_head, float* dwk_head,
                  float* workspace, int64_t workspace_floats, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core (tcgen05 / TMEM) convolution path: same contract as nq_conv_fwd / nq_conv_dgrad, bf16
 * operands split hi + lo so that products carry 16 mantissa bits, fp32 accumulation in TMEM.
 * Channel contract: cin_p a multiple of 8; GEMM dims are padded to 16 internally (zero channels / unwritten columns).
 * ------------------------------------------------------------------------------------------------ */
typedef struct nq_tc_plan {
  int32_t dir;                 /* 0 forward, 1 data gradient */
  int32_t C, N;                /* GEMM-K channels per tap, GEMM-N columns */
  int32_t NT, KC, SBC;         /* N tile, channels per activation unit, channels per weight stage */
  int32_t a_planes, b_planes;  /* bf16 planes of the activation / weight operand (1 = hi, 2 = hi + lo) */
  int32_t PW, PH, CGS;         /* halo tile width, height (pixels); channel-group stride (bytes) */
  int32_t a_plane_bytes, a_buf_bytes, b_stage_bytes, n_bstages, smem_bytes;
  int32_t tiles_x, tiles_y, tiles_n, total_tiles;
  int32_t cluster;             /* CTAs per cluster sharing each weight stage by TMA multicast (caller may set 1, 2, 4) */
  int32_t mt;                  /* 16x8 pixel tiles per CTA step sharing each weight stage (1, or 2 when NT <= 128) */
  int32_t bcat;                /* 1: weight planes packed side by side per k-group, 2 MMAs per K step instead of 3 */
  int32_t n_abuf, n_acc;       /* ring depths: activation buffers in shared memory, accumulator slots in TMEM */
  int32_t acc_stride;          /* TMEM columns per accumulator slot */
  int32_t n_epi;               /* epilogue warps (8 or 12 of the 16 worker warps; the rest load activations) */
  int32_t resident;            /* 1: all weight stages of a tile fit the ring; loaded once per CTA */
  int32_t ksplit;              /* data gradient: CTAs sharing the K range of one tile (1 = off; the caller may set 1) */
  int32_t cg2;                 /* 1: CTA-pair MMAs (tcgen05 cta_group::2, M = 256): two pixel tiles per MMA, each CTA stages half of
                                  every weight stage; needs cluster = 2 (forced at launch), excludes resident; with bcat the leader
                                  stages the hi plane, the peer the lo plane, both also their half of the hi plane's columns */
  int32_t gst;                 /* weight stages per ring slot (one bulk copy and one barrier round trip for all of them); the ring has
                                  n_bstages slots of gst * b_stage_bytes bytes */
  int32_t reserved;
  int64_t wpk_bytes;           /* size of the packed weight buffer the caller allocates.  The order AND the size of the packed
                                  operand belong to this plan, and the plan depends on every field of the descriptor, d->n included
                                  (CTA pairs are chosen from 32 pixel tiles on; cg2 with bcat packs 6 instead of 4 bytes per weight):
                                  pack with the plan the convolution is launched with, re-pack when the batch size changes */
  int64_t workspace_floats;    /* fp32 partial sums nq_tc_conv_dgrad needs when ksplit > 1 */
} nq_tc_plan;

/* Tiling / staging plan of one stage and direction; pure host arithmetic. */
int nq_tc_plan_conv(const nq_conv_desc* d, int dir, int a_planes, int b_planes, nq_tc_plan* plan);

/* ref-layout weights (cout, cin_src, k, k) -> bf16 stage order of `plan` (wpk: plan->wpk_bytes bytes).
 * zero_point (may be NULL) is subtracted per output channel first: passing the quantiser's codes
 * (quantizer.py:297) and zero point packs the INTEGER weights code - zp, exact in one bf16 plane for
 * n_bits <= 8; the per-channel step size then goes to nq_tc_pack_epilogue / the conv epilogue. */
int nq_tc_pack_weight(const nq_conv_desc* d, const nq_tc_plan* plan, const float* w_ref, int cin_src,
                      const float* zero_point, int zp_stride, void* wpk, void* stream);

/* Per-column epilogue vectors in packed channel order: scale_packed[n'] = delta[co * d_stride] (1 when
 * delta is NULL or n' is a pad column), bias_packed[n'] = bias_ref[co] (0 for pads). */
int nq_tc_pack_epilogue(const nq_conv_desc* d, const float* delta, int d_stride, const float* bias_ref,
                        float* scale_packed, float* bias_packed, void* stream);

/* Several nq_tc_pack_weight (wpk != NULL) and / or nq_tc_pack_epilogue (scale_packed or bias_packed != NULL) calls
 * of the same or different stages in one launch: the weights of a calibration iteration are re-packed for every
 * stage and direction after each optimiser step, 19 launches of 3-8 us otherwise. */
typedef struct nq_tc_pack_task {
  const nq_conv_desc* d;
  const nq_tc_plan* plan;      /* may be NULL when wpk is NULL */
  const float* w_ref; const float* zero_point; void* wpk;
  const float* delta; const float* bias_ref; float* scale_packed; float* bias_packed;
  int32_t cin_src, zp_stride, d_stride, reserved;
} nq_tc_pack_task;
int nq_tc_pack_multi(const nq_tc_pack_task* tasks, int n_tasks, void* stream);

/* "split-bf16" storage: a tensor-core stage reads and writes its activations / gradients as TWO bf16
 * planes, hi = bf16(v) and lo = bf16(v - hi), plane 0 then plane 1, each NHWC with the padded channel
 * count of the fp32 layout (same bytes as fp32, 16 mantissa bits).  The producing epilogue converts once;
 * consumers copy 16-byte chunks asynchronously.  Pre-activations (z) stay fp32. */

/* y = act(shuffle(conv(x) * scale + bias)).  x_split (n, h, w, cin_p) and y_split (n, h*rh, w*rw, cg) are
 * split-bf16; z (fp32, same shape as y) receives the pre-activation.  Either output may be NULL (not both);
 * scale_packed / bias_packed may be NULL. */
int nq_tc_conv_fwd(const nq_conv_desc* d, const nq_tc_plan* plan, const void* x_split, const void* wpk,
                   const float* scale_packed, const float* bias_packed, float* z, void* y_split, void* stream);

/* Head on the tensor cores: nq_head_fwd_loss_split's contract (3x3 conv to 3 channels + OutImg + loss sum + dL/dz)
 * with the convolution as an N = 16 (3 real columns) tcgen05 GEMM; wpk / scale / bias packed as for any stage
 * with a dir = 0 plan of the head's descriptor. */
int nq_tc_head_fwd_loss(const nq_conv_desc* d, const nq_tc_plan* plan, const void* x_split, const void* wpk,
                        const float* scale_packed, const float* bias_packed, int out_bias, const float* target,
                        float p, float mean_pixels, float* img, float* loss_sum, void* dz_head_split, void* stream);

/* dz_prev = unshuffle(conv_transpose(dz) * act'(z_prev)).  dz_split (n, h, w, nout_p rounded up to 8) and
 * dz_prev_split (n, h/prev_rh, w/prev_rw, prev_rh*prev_rw*cin_p) are split-bf16, z_prev fp32; workspace
 * (plan->workspace_floats floats, may be NULL when plan->ksplit == 1) holds the split-K partial sums; wpk_t is
 * packed with a dir = 1 plan from the DE-QUANTISED weights. */
int nq_tc_conv_dgrad(const nq_conv_desc* d, const nq_tc_plan* plan, const void* dz_split, const void* wpk_t,
                     const float* z_prev, int prev_rh, int prev_rw, int prev_act, void* dz_prev_split, float* workspace, int64_t workspace_floats, void* stream);

/* Edges of the split-bf16 domain.  NCHW fp32 (n, c, h, w) <-> split NHWC (n, h, w, c_p); flat fp32 <-> split. */
int nq_nchw_to_split(const float* src, void* dst_split, int n, int c, int h, int w, int c_p, void* stream);
int nq_split_to_nchw(const void* src_split, float* dst, int n, int c, int h, int w, int c_p, void* stream);
int nq_f32_to_split(const float* src, void* dst_split, int64_t numel, void* stream);
int nq_split_to_f32(const void* src_split, float* dst, int64_t numel, void* stream);
/* Frame ingest (videosets/datasets.py:8-54: `read_image(path) / 255.0`): uint8 -> fp32 value / 255, IEEE division, so the
 * result is bit-identical to the reference's host-side conversion.  16-byte aligned pointers. */
int nq_u8_to_f32(const uint8_t* src, float* dst, int64_t numel, void* stream);

/* Weight + bias gradient on the tensor cores; output contract identical to nq_conv_wgrad (dwk
 * [(kdim + 4)][nout_p], bias gradient in row kdim).  cin_p % 8 == 0, rh*rw*cg % 16 == 0. */
typedef struct nq_tc_wgrad_plan {
  int32_t C, N, a_planes, b_planes;
  int32_t ncg, G, MB, NC, nsplits, TR;
  int32_t nkh, khg, AR;        /* kernel rows per CTA, kernel-row groups, staged input rows */
  int32_t msplit, ncg_c;       /* input channel-group slices (GEMM-M split across CTAs), groups per slice */
  int32_t bcat;                /* 1: both dZ planes as one MMA operand (2 MMAs per pixel step instead of 3) */
  int32_t CGS_A, CGS_B, a_plane_bytes, b_plane_bytes, buf_bytes, nbuf, smem_bytes;
  int32_t tiles_x, tiles_y, tiles_total, psplits, tiles_per_split;
  int64_t workspace_floats;    /* partial-gradient workspace the caller provides */
} nq_tc_wgrad_plan;
int nq_tc_plan_wgrad(const nq_conv_desc* d, int a_planes, int b_planes, nq_tc_wgrad_plan* plan);
/* x_split (n, h, w, cin_p) and dz_split (n, h, w, nout_p rounded up to 8) are split-bf16. */
int nq_tc_conv_wgrad(const nq_conv_desc* d, const nq_tc_wgrad_plan* plan, const void* x_split, const void* dz_split,
                     float* dwk, float* workspace, int64_t workspace_floats, void* stream);

/* One launch that finishes the weight gradients of several stages: the fixed-order sum over each stage's
 * pixel-split partials (what nq_tc_conv_wgrad does itself when dwk != NULL) fused with nq_unpack_wgrad.  Call
 * nq_tc_conv_wgrad with dwk = NULL for those stages.  d is the descriptor the wgrad plan was made for; channels
 * cin .. cin_dst of dw_ref (the zero pad of a rotated layer, quant_layer.py:66) are written as zeros. */
typedef struct nq_wgrad_finish_task {
  const nq_conv_desc* d;
  const float* workspace;   /* (plan->psplits, k*k*cin_p + 4, plan->N) */
  float* dw_ref;            /* (cout, cin_dst, k, k) or NULL */
  float* db_ref;            /* (cout) or NULL */
  int32_t psplits, n_cols;  /* plan->psplits, plan->N */
  int32_t cin_dst, reserved;
} nq_wgrad_finish_task;
int nq_tc_wgrad_finish_multi(const nq_wgrad_finish_task* tasks, int n_tasks, void* stream);

/* Head weight gradient, tap-expanded (the backward counterpart of nq_head_fwd_loss_tapexp): input pixels as GEMM-K, the
 * 27 (tap, channel) pairs as columns of a dZ operand assembled with the nine shifts, one extra row of ones for the bias
 * gradient.  Writes nq_head_wgrad_tapexp_splits(d) partial sums of (9 * cin_p + 4) x 16 floats into `workspace`, in the
 * layout nq_tc_wgrad_finish_multi reads (psplits = that count, n_cols = 16, d with cg = 16).  cin_p <= 64. */
int nq_head_wgrad_tapexp_splits(const nq_conv_desc* d);
int nq_head_wgrad_tapexp(const nq_conv_desc* d, const void* x_split, const void* dz_split, float* workspace,
                         int64_t workspace_floats, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Omega = dw^T H dw (methods/bit_assign.py:57-118,171-203) by second-order FORWARD propagation:
 * every stage carries (y, y', y'') = value and first/second directional derivative along the weight
 * perturbation; the convolutions are the forward kernels above applied to y, y', y'' with the weights
 * w or the perturbation v; these two kernels are the elementwise chain rule and the MSE head.
 * ------------------------------------------------------------------------------------------------ */
/* z' = zd1 + zd2, z'' = zdd1 + 2*zdd2 (any of the four may be NULL = 0); act as nq_conv_desc.act.
 * y = f(z), yd = f'(z) z', ydd = f''(z) z'^2 + f'(z) z''. */
int nq_jet_act(const float* z, const float* zd1, const float* zd2, const float* zdd1, const float* zdd2,
               int64_t numel, int act, void* y, void* yd, void* ydd, int split_out, void* stream);
/* split_out = 1: y, yd, ydd are written as split-bf16 (2 * numel bf16 each) for the tensor-core convolutions. */
/* Head pre-activations (n, h, w, 4) NHWC, target (n, 3, h, w): *omega_acc += d^2/d eps^2 of
 * nn.MSELoss(OutImg(z), target) (mean over n*3*h*w).  omega_acc is a device double the caller zeroes. */
int nq_jet_head(const float* z, const float* zd1, const float* zd2, const float* zdd1, const float* zdd2,
                const float* target, int n, int h, int w, int out_bias, double* omega_acc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Layout edges and reductions
 * ------------------------------------------------------------------------------------------------ */
/* Packed weight codes of the quantised artefact (readme.md:125-127: the hand-off to entropy coding; the reference keeps
 * the codes as fp32, quantizer.py:297 / quant_model.py:74-80).  Integer codes 0 .. 2^n_bits - 1 <-> a dense little-endian
 * bit stream: element i occupies bits [i * n_bits, (i + 1) * n_bits); the stream is padded to whole groups of eight
 * elements, nq_packed_bytes(numel, n_bits) = ceil(numel / 8) * n_bits bytes.  nq_pack_codes sets *not_integer_flag (device
 * int, may be NULL) when an input is not an integer in range (e.g. soft-rounded codes) and stores 0 for it. */
int64_t nq_packed_bytes(int64_t numel, int n_bits);
int nq_pack_codes(const float* codes, int64_t numel, int n_bits, void* packed, int* not_integer_flag, void* stream);
int nq_unpack_codes(const void* packed, int64_t numel, int n_bits, float* codes, void* stream);

/* NCHW (n, c, h, w) <-> NHWC (n, h, w, c_p); pad channels are written as zero / ignored. */
int nq_nchw_to_nhwc(const float* src, float* dst, int n, int c, int h, int w, int c_p, void* stream);
int nq_nhwc_to_nchw(const float* src, float* dst, int n, int c, int h, int w, int c_p, void* stream);

/* Activation backward + un-shuffle for a stage whose consumer is outside this library (autograd of
 * nn.GELU + nn.PixelShuffle, quant_block.py:33-34): dy, z (n, h*rh, w*rw, cg) -> dz (n, h, w, rh*rw*cg). */
int nq_act_bwd_unshuffle(const float* dy, const float* z, int n, int h, int w, int rh, int rw, int cg,
                         int act, float* dz, void* stream);

/* Block-wise reconstruction (calib_block.py:62-63,168-170): lp_loss of one stage's output y against its
 * full-precision output tgt, fused with the backward through activation + up-shuffle.  y_split (n, h*rh, w*rw, cg)
 * split-bf16 as nq_tc_conv_fwd writes it, tgt the same grid in fp32 NHWC (pad channels zero), gprime the saved
 * GELU'(pre-activation) (nq_conv_desc.act = 2; NULL for a linear stage).  frame_idx (device, n entries, may be NULL):
 * batch entry b compares against frame frame_idx[b] of a target cache (N, h*rh, w*rw, cg).  *loss_sum += sum |y - tgt|^p;
 * dz_split (n, h, w, rh*rw*cg) = unshuffle(grad_scale * p |d|^(p-1) sign(d) * gprime), ready for nq_tc_conv_wgrad. */
int nq_block_loss_bwd(const void* y_split, const float* tgt, const int32_t* frame_idx, const float* gprime, int n, int h, int w,
                      int rh, int rw, int cg, float p, float grad_scale, float* loss_sum, void* dz_split, void* stream);

/* Mini-batch assembly of block_reconstruction (calib_block.py:160-164: cached_inps[0][idx], cached_inps[1][idx] and
 * torch.where(torch.rand_like(cur_inp) < input_prob, cur_inp, cur_sym)) on the HBM-resident caches.  inp_split / sym_split:
 * (2, n_cache, frame_elems) split-bf16 caches (sym_split and rnd NULL: plain gather, input_prob = 1); frame_idx: n device
 * ints; rnd: (n, frame_elems) uniform draws in the batch's own layout; out_split (2, n, frame_elems). */
int nq_qdrop_gather(const void* inp_split, const void* sym_split, const int32_t* frame_idx, const float* rnd, float input_prob,
                    int n, int n_cache, int64_t frame_elems, void* out_split, void* stream);

/* Fisher-weighted block losses (LossFunction.__call__, calib_block.py:66-72) on the layouts of nq_block_loss_bwd.
 * fisher: the cached |dL/d(block output)| + 1 (save_grad_data, data_utils.py:91-119), fp32 NHWC (N, h*rh, w*rw, cg) like
 * the target cache and addressed through the same frame_idx.
 *   mode 1 'fisher_diag': *loss_sum += sum d^2 F^2 (caller divides by n*H*W);  dy = grad_scale * 2 d F^2
 *   mode 2 'fisher_full': frame_dot[b] = sum_{c,h,w} |d| F (n floats, overwritten); dy = grad_scale * 2 frame_dot[b] F sign(d)
 *                         (the loss is sum_b frame_dot[b]^2 / (n*C*H*W*100), formed by the caller; loss_sum is not touched)
 * dz_split as nq_block_loss_bwd: dy * gprime, un-shuffled, split-bf16. */
int nq_block_loss_bwd_fisher(const void* y_split, const float* tgt, const float* fisher, const int32_t* frame_idx,
                             const float* gprime, int n, int h, int w, int rh, int rw, int cg, int mode, float grad_scale,
                             float* loss_sum, float* frame_dot, void* dz_split, void* stream);

/* lp_loss (quantizer.py:66-73) standalone: *loss_sum += sum |pred - tgt|^p; grad (may be NULL) receives
 * grad_scale * p * |d|^(p-1) * sign(d). */
int nq_lp_loss(const float* pred, const float* tgt, int64_t numel, float p, float grad_scale,
               float* loss_sum, float* grad, void* stream);

/* Multi-tensor reductions for bit_assign (bit_assign.py:198-200, :211): out[t] = sum_i a_t[i]*b_t[i]
 * (mode 0, Omega) or sum_i a_t[i]^2 * b_t[i]^2 (mode 1, diagonal Fisher).  a_ptrs/b_ptrs/sizes are HOST
 * arrays of n_tensors entries (device pointers inside); out is a device array, overwritten. */
int nq_multi_dot(const float* const* a_ptrs, const float* const* b_ptrs, const int64_t* sizes,
                 int n_tensors, int mode, float* out, void* stream);

/* bit_assign as a search (the reference scores a hand-written candidate list, methods/bit_assign.py:343-372, one
 * Hessian-vector product each).  Omega is quadratic in the perturbation, so with the Gram table
 * gram[(l, b), (m, b')] = v_l(b)^T H_lm v_m(b') (row-major (L * nb)^2 doubles, symmetric; option index fastest within a
 * layer) the score of configuration (c_0 .. c_{L-1}) is sum_l gram[(l,c_l),(l,c_l)] + 2 sum_{l<m} gram[(l,c_l),(m,c_m)].
 * Scores all nb^L configurations (index = sum_l c_l * nb^l) and returns the admissible one -- sum_l bits_weight[l * nb + c_l]
 * <= budget -- of smallest score (ties: smallest index); best_index = -1 when none is admissible.  `scores` (optional,
 * nb^L doubles) receives every admissible configuration's score, +inf elsewhere.  n_layers <= 8, n_options <= 16. */
int nq_omega_search_workspace(int n_layers, int n_options, int64_t* n_configs, int64_t* workspace_bytes);
int nq_omega_search(const double* gram, const double* bits_weight, int n_layers, int n_options, double budget,
                    double* scores, void* workspace, int64_t workspace_bytes, double* best_score, int64_t* best_index,
                    void* stream);

/* PSNR per frame (utils.py:148-151): psnr[i] = -10 log10(mean((a_i - b_i)^2) + 1e-9), frames of
 * `frame_numel` elements. */
int nq_psnr(const float* a, const float* b, int n_frames, int64_t frame_numel, float* psnr, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NEUROQUANT_B200_H_ */
