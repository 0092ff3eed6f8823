// Rotation (FWHT), weight packing, layout edges, loss / dot / PSNR reductions, library plumbing.
#include <cuda_bf16.h>
#include "nq_common.cuh"

namespace nq {

thread_local int g_last_cuda_error = 0;

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

static inline int grid_for(int64_t numel, int block = 256) {
  int64_t b = (numel + block - 1) / block;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

// ---------------------------------------------------------------------------------------------
// Warp-shuffle Walsh-Hadamard: one warp per vector, element k = lane + 32*r held in register r of
// lane `lane` (n <= 256 -> R <= 8).  Butterflies over bits 0-4 are lane exchanges, over bits 5-7 are
// register exchanges.  Same stage order (h = 1, 2, 4, ...) and operand order (a+b, a-b) as the
// Sylvester butterfly the reference's hadamard_transform package performs.
// src and dst MAY ALIAS (the engine rotates in place): neither is __restrict__, and a warp reads its whole vector
// into registers before it stores any of it, so in-place operation is well defined.
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256) fwht_kernel(const float* src, float* dst,
                                                   int64_t n_vectors, int n, int64_t inner,
                                                   int64_t outer_stride, float norm) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t v = warp; v < n_vectors; v += n_warps) {
    const int64_t base = (v / inner) * outer_stride + (v % inner);
    float r[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int k = lane + 32 * i;
      r[i] = (k < n) ? src[base + (int64_t)k * inner] : 0.f;
    }
#pragma unroll
    for (int h = 1; h < 32; h <<= 1) {
      if (h < n) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const float o = __shfl_xor_sync(0xffffffffu, r[i], h);
          r[i] = (lane & h) ? (o - r[i]) : (r[i] + o);
        }
      }
    }
#pragma unroll
    for (int hr = 1; hr < R; hr <<= 1) {
      if (32 * hr < n) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
          if ((i & hr) == 0) {
            const float a = r[i], b = r[i | hr];
            r[i] = a + b;
            r[i | hr] = a - b;
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int k = lane + 32 * i;
      if (k < n) dst[base + (int64_t)k * inner] = __fdiv_rn(r[i], norm);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// ref <-> packed weight layouts
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int packed_channel(int co, int rh, int rw, int cg) {
  const int rr = rh * rw;
  const int c = co / rr, rem = co - c * rr;
  return rem * cg + c;
}

__global__ void __launch_bounds__(256) pack_weight_kernel(nq_conv_desc d, const float* __restrict__ w_ref,
                                                          int cin_src, const float* __restrict__ bias_ref,
                                                          float* __restrict__ wk, float* __restrict__ wt,
                                                          float* __restrict__ bias_packed) {
  const int kk = d.ksize * d.ksize;
  const int nout_p = d.rh * d.rw * d.cg;
  const int64_t total = (int64_t)d.cout * d.cin * kk;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int tap = (int)(e % kk);
    const int ci = (int)((e / kk) % d.cin);
    const int co = (int)(e / ((int64_t)kk * d.cin));
    const float v = w_ref[((int64_t)co * cin_src + ci) * kk + tap];
    const int np = packed_channel(co, d.rh, d.rw, d.cg);
    if (wk) wk[((int64_t)tap * d.cin_p + ci) * nout_p + np] = v;
    if (wt) wt[((int64_t)(kk - 1 - tap) * nout_p + np) * d.cin_p + ci] = v;
  }
  if (bias_ref && bias_packed) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < d.cout;
         e += (int64_t)gridDim.x * blockDim.x)
      bias_packed[packed_channel((int)e, d.rh, d.rw, d.cg)] = bias_ref[e];
  }
}

__global__ void __launch_bounds__(256) unpack_wgrad_kernel(nq_conv_desc d, const float* __restrict__ dwk,
                                                           int cin_dst, float* __restrict__ dw_ref,
                                                           float* __restrict__ db_ref) {
  const int kk = d.ksize * d.ksize;
  const int nout_p = d.rh * d.rw * d.cg;
  const int64_t kdim = (int64_t)kk * d.cin_p;
  if (dw_ref) {
    const int64_t total = (int64_t)d.cout * cin_dst * kk;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
      const int tap = (int)(e % kk);
      const int ci = (int)((e / kk) % cin_dst);
      const int co = (int)(e / ((int64_t)kk * cin_dst));
      float v = 0.f;
      if (ci < d.cin) v = dwk[((int64_t)tap * d.cin_p + ci) * nout_p + packed_channel(co, d.rh, d.rw, d.cg)];
      dw_ref[e] = v;
    }
  }
  if (db_ref) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < d.cout;
         e += (int64_t)gridDim.x * blockDim.x)
      db_ref[e] = dwk[kdim * nout_p + packed_channel((int)e, d.rh, d.rw, d.cg)];
  }
}

// ---------------------------------------------------------------------------------------------
// NCHW <-> NHWC edges (embedding in, debug/feature taps out).  Small tensors; one thread per
// destination element, destination-coalesced.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                           int n, int c, int h, int w, int c_p) {
  const int64_t total = (int64_t)n * h * w * c_p;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(e % c_p);
    const int64_t pix = e / c_p;
    const int x = (int)(pix % w), y = (int)((pix / w) % h), b = (int)(pix / ((int64_t)w * h));
    dst[e] = ch < c ? src[(((int64_t)b * c + ch) * h + y) * w + x] : 0.f;
  }
}
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                           int n, int c, int h, int w, int c_p) {
  const int64_t total = (int64_t)n * c * h * w;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(e % w), y = (int)((e / w) % h);
    const int ch = (int)((e / ((int64_t)w * h)) % c), b = (int)(e / ((int64_t)w * h * c));
    dst[e] = src[(((int64_t)b * h + y) * w + x) * c_p + ch];
  }
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float abs_pow(float a, float p) { return p == 2.0f ? a * a : powf(a, p); }

__global__ void __launch_bounds__(256) lp_loss_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                                      int64_t numel, float p, float grad_scale,
                                                      float* __restrict__ loss_sum, float* __restrict__ grad) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < numel;
       e += (int64_t)gridDim.x * blockDim.x) {
    const float dlt = pred[e] - tgt[e];
    const float a = fabsf(dlt);
    acc += abs_pow(a, p);
    if (grad) {
      const float sgn = dlt > 0.f ? 1.f : (dlt < 0.f ? -1.f : 0.f);
      grad[e] = grad_scale * (p == 2.0f ? 2.0f * dlt : p * powf(a, p - 1.0f) * sgn);
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss_sum, acc);
}

struct DotTable {
  const float* a[16];
  const float* b[16];
  int64_t n[16];
};

// one CTA column per tensor (blockIdx.y), fixed-order two-level reduction -> deterministic
__global__ void __launch_bounds__(256) multi_dot_partial_kernel(DotTable t, int mode, float* __restrict__ partial) {
  __shared__ float red[32];
  const float* a = t.a[blockIdx.y];
  const float* b = t.b[blockIdx.y];
  const int64_t n = t.n[blockIdx.y];
  float acc = 0.f;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const float x = a[e], y = b[e];
    acc += mode == 0 ? x * y : (x * x) * (y * y);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = acc;
}
__global__ void __launch_bounds__(32) multi_dot_final_kernel(const float* __restrict__ partial, int per_tensor,
                                                             float* __restrict__ out) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < per_tensor; i += 32) acc += partial[blockIdx.x * per_tensor + i];
  acc = warp_sum(acc);
  if (threadIdx.x == 0) out[blockIdx.x] = acc;
}

// one CTA-group per frame; blockIdx.y = frame.  Two-pass (partials then log) kept in one kernel via
// atomics on a per-frame accumulator would be non-deterministic; frames are small in number, so use
// grid.x CTAs per frame + a last-block finaliser.
__global__ void __launch_bounds__(256) psnr_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                           int64_t frame_numel, float* __restrict__ partial) {
  __shared__ float red[32];
  const float* pa = a + (int64_t)blockIdx.y * frame_numel;
  const float* pb = b + (int64_t)blockIdx.y * frame_numel;
  float acc = 0.f;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < frame_numel;
       e += (int64_t)gridDim.x * blockDim.x) {
    const float d = pa[e] - pb[e];
    acc += d * d;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = acc;
}
__global__ void __launch_bounds__(32) psnr_final_kernel(const float* __restrict__ partial, int per_frame,
                                                        float inv_numel, float* __restrict__ psnr) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < per_frame; i += 32) acc += partial[blockIdx.x * per_frame + i];
  acc = warp_sum(acc);
  if (threadIdx.x == 0) psnr[blockIdx.x] = -10.0f * log10f(acc * inv_numel + 1e-9f);
}

}  // namespace nq

using namespace nq;

extern "C" const char* nq_status_string(int status) {
  switch (status) {
    case NQ_OK: return "ok";
    case NQ_ERR_BAD_ARG: return "bad argument (null pointer, non-positive size, or bit-width outside 2..8)";
    case NQ_ERR_BAD_SHAPE: return "shape violates the layout contract (padding / divisibility)";
    case NQ_ERR_UNSUPPORTED: return "unsupported configuration";
    case NQ_ERR_WORKSPACE: return "workspace too small";
    case NQ_ERR_CUDA: return "CUDA error (see nq_last_cuda_error)";
    default: return "unknown status";
  }
}
extern "C" int nq_last_cuda_error(void) { return g_last_cuda_error; }
extern "C" int nq_abi_version(void) { return 1; }
extern "C" int nq_sm_count(void) { return sm_count(); }

extern "C" int nq_fwht(const float* src, float* dst, int64_t n_vectors, int n, int64_t inner, int64_t outer_stride,
                       void* stream) {
  if (!src || !dst || n_vectors <= 0 || inner <= 0) return NQ_ERR_BAD_ARG;
  if (n < 1 || n > 256 || (n & (n - 1)) != 0) return NQ_ERR_UNSUPPORTED;
  const float norm = (float)sqrt((double)n);
  int64_t blocks = (n_vectors + 7) / 8;  // 8 warps per CTA
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = as_stream(stream);
  if (n <= 32) fwht_kernel<1><<<(unsigned)blocks, 256, 0, s>>>(src, dst, n_vectors, n, inner, outer_stride, norm);
  else if (n == 64) fwht_kernel<2><<<(unsigned)blocks, 256, 0, s>>>(src, dst, n_vectors, n, inner, outer_stride, norm);
  else if (n == 128) fwht_kernel<4><<<(unsigned)blocks, 256, 0, s>>>(src, dst, n_vectors, n, inner, outer_stride, norm);
  else fwht_kernel<8><<<(unsigned)blocks, 256, 0, s>>>(src, dst, n_vectors, n, inner, outer_stride, norm);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

static int check_desc(const nq_conv_desc* d) {
  if (!d) return NQ_ERR_BAD_ARG;
  if (d->n <= 0 || d->h <= 0 || d->w <= 0 || d->cin <= 0 || d->cout <= 0 || d->rh <= 0 || d->rw <= 0) return NQ_ERR_BAD_ARG;
  if (d->ksize < 1 || (d->ksize & 1) == 0) return NQ_ERR_BAD_SHAPE;
  if (d->cin_p < d->cin || (d->cin_p & 3) || d->cg < d->c_grp || (d->cg & 3)) return NQ_ERR_BAD_SHAPE;
  if (d->cout != d->c_grp * d->rh * d->rw) return NQ_ERR_BAD_SHAPE;
  if (d->act < 0 || d->act > 2) return NQ_ERR_BAD_ARG;
  return NQ_OK;
}
namespace nq { int check_conv_desc(const nq_conv_desc* d) { return check_desc(d); } }

extern "C" int nq_pack_weight(const nq_conv_desc* d, const float* w_ref, int cin_src, const float* bias_ref,
                              float* wk, float* wt, float* bias_packed, void* stream) {
  int st = check_desc(d);
  if (st) return st;
  if (!w_ref || cin_src < d->cin) return NQ_ERR_BAD_ARG;
  const int64_t total = (int64_t)d->cout * d->cin * d->ksize * d->ksize;
  pack_weight_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(*d, w_ref, cin_src, bias_ref, wk, wt, bias_packed);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

extern "C" int nq_unpack_wgrad(const nq_conv_desc* d, const float* dwk, int cin_dst, float* dw_ref, float* db_ref,
                               void* stream) {
  int st = check_desc(d);
  if (st) return st;
  if (!dwk || cin_dst < d->cin) return NQ_ERR_BAD_ARG;
  const int64_t total = (int64_t)d->cout * cin_dst * d->ksize * d->ksize;
  unpack_wgrad_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(*d, dwk, cin_dst, dw_ref, db_ref);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

// ---------------------------------------------------------------------------------------------
// Block-wise reconstruction (calib_block.py:62-63,168-170): lp_loss of ONE stage's output against its full-precision
// output, with the backward through the activation and the up-shuffle fused -- the block's output gradient goes
// straight into the split-bf16 dZ the weight-gradient kernel consumes.  One pass, 16 bytes per element.
// ---------------------------------------------------------------------------------------------
namespace nq {
// one thread = 8 consecutive channels of one output pixel: 16-byte accesses on every operand, 32-bit index math
__global__ void __launch_bounds__(256) block_loss_bwd_kernel(const uint4* __restrict__ y_hi, const uint4* __restrict__ y_lo,
                                                             const float4* __restrict__ tgt, const int* __restrict__ frame_idx,
                                                             const float4* __restrict__ gprime,
                                                             int n, int h, int w, int rh, int rw, int cg, float p,
                                                             float grad_scale, float* __restrict__ loss_sum,
                                                             uint4* __restrict__ dz_hi, uint4* __restrict__ dz_lo) {
  __shared__ float red[32];
  const int c8n = cg >> 3;
  const int W2 = w * rw, H2 = h * rh;
  const int total = n * H2 * W2 * c8n;       // 8-channel groups (checked < 2^31 on the host)
  const int frame8 = H2 * W2 * c8n;
  float loss = 0.f;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int c8 = e % c8n;
    const int pix = e / c8n;
    const int x = pix % W2, t = pix / W2;
    const int yy = t % H2, b = t / H2;
    const int te = frame_idx != nullptr ? e + (frame_idx[b] - b) * frame8 : e;  // target cache of many frames
    const uint4 yh = y_hi[e], yl = y_lo[e];
    const float4 t0 = tgt[2 * (size_t)te], t1 = tgt[2 * (size_t)te + 1];
    float4 g0 = make_float4(1.f, 1.f, 1.f, 1.f), g1 = g0;
    if (gprime != nullptr) { g0 = gprime[2 * (size_t)e]; g1 = gprime[2 * (size_t)e + 1]; }
    const uint32_t hw[4] = {yh.x, yh.y, yh.z, yh.w}, lw[4] = {yl.x, yl.y, yl.z, yl.w};
    const float tv[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    uint32_t oh[4], ol[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gr[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint32_t hb = q ? (hw[k] & 0xFFFF0000u) : (hw[k] << 16), lb = q ? (lw[k] & 0xFFFF0000u) : (lw[k] << 16);
        const float d = (__uint_as_float(hb) + __uint_as_float(lb)) - tv[2 * k + q];
        float g;
        if (p == 2.0f) {
          loss += d * d;
          g = 2.0f * d;
        } else {
          const float a = fabsf(d);
          loss += powf(a, p);
          g = p * powf(a, p - 1.0f) * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
        }
        gr[q] = g * grad_scale * gv[2 * k + q];
      }
      const __nv_bfloat16 h0 = __float2bfloat16_rn(gr[0]), h1 = __float2bfloat16_rn(gr[1]);
      oh[k] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      ol[k] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(gr[0] - __bfloat162float(h0))) |
              ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(gr[1] - __bfloat162float(h1))) << 16);
    }
    const int qh = yy / rh, si = yy - qh * rh, qw = x / rw, sj = x - qw * rw;
    const size_t o = ((size_t)((b * h + qh) * w + qw) * (rh * rw) + (si * rw + sj)) * c8n + c8;
    dz_hi[o] = make_uint4(oh[0], oh[1], oh[2], oh[3]);
    dz_lo[o] = make_uint4(ol[0], ol[1], ol[2], ol[3]);
  }
  loss = block_sum(loss, red);
  if (threadIdx.x == 0 && loss_sum != nullptr) atomicAdd(loss_sum, loss);
}
}  // namespace nq

// ---------------------------------------------------------------------------------------------
// Mini-batch assembly of block-wise reconstruction (calib_block.py:160-164): gather the batch's frames from the
// HBM-resident split-bf16 input cache and, with QDrop, take each element from the quantised-predecessor cache or the
// full-precision one according to a uniform draw -- one pass instead of four frame copies + compare + where.
// ---------------------------------------------------------------------------------------------
namespace nq {
__global__ void __launch_bounds__(256) qdrop_gather_kernel(const uint4* __restrict__ inp_hi, const uint4* __restrict__ inp_lo,
                                                           const uint4* __restrict__ sym_hi, const uint4* __restrict__ sym_lo,
                                                           const int* __restrict__ frame_idx, const float4* __restrict__ rnd,
                                                           float prob, int n, int frame8, uint4* __restrict__ out_hi,
                                                           uint4* __restrict__ out_lo) {
  const int total = n * frame8;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int b = e / frame8, f = e - b * frame8;
    const size_t src = (size_t)frame_idx[b] * frame8 + f;
    uint4 h = inp_hi[src], l = inp_lo[src];
    if (sym_hi != nullptr) {
      const float4 r0 = rnd[2 * (size_t)e], r1 = rnd[2 * (size_t)e + 1];
      const uint4 sh = sym_hi[src], sl = sym_lo[src];
      const float rv[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
      uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
      const uint32_t shw[4] = {sh.x, sh.y, sh.z, sh.w}, slw[4] = {sl.x, sl.y, sl.z, sl.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // torch.where(rand < input_prob, cur_inp, cur_sym): per 16-bit half of each word
        const uint32_t m = (rv[2 * k] < prob ? 0x0000FFFFu : 0u) | (rv[2 * k + 1] < prob ? 0xFFFF0000u : 0u);
        hw[k] = (hw[k] & m) | (shw[k] & ~m);
        lw[k] = (lw[k] & m) | (slw[k] & ~m);
      }
      h = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      l = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
    out_hi[e] = h;
    out_lo[e] = l;
  }
}
}  // namespace nq

extern "C" int nq_qdrop_gather(const void* inp_split, const void* sym_split, const int32_t* frame_idx, const float* rnd,
                               float input_prob, int n, int n_cache, int64_t frame_elems, void* out_split, void* stream) {
  if (!inp_split || !frame_idx || !out_split || n <= 0 || n_cache <= 0 || frame_elems <= 0) return NQ_ERR_BAD_ARG;
  if ((sym_split != nullptr) != (rnd != nullptr)) return NQ_ERR_BAD_ARG;
  if (frame_elems % 8 || (int64_t)n * frame_elems / 8 >= (1LL << 31)) return NQ_ERR_BAD_SHAPE;
  const uint16_t* ih = reinterpret_cast<const uint16_t*>(inp_split);
  const uint16_t* sh = reinterpret_cast<const uint16_t*>(sym_split);
  uint16_t* oh = reinterpret_cast<uint16_t*>(out_split);
  const int64_t cache_plane = (int64_t)n_cache * frame_elems, out_plane = (int64_t)n * frame_elems;
  qdrop_gather_kernel<<<grid_for(out_plane / 8), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(ih), reinterpret_cast<const uint4*>(ih + cache_plane),
      sh ? reinterpret_cast<const uint4*>(sh) : nullptr, sh ? reinterpret_cast<const uint4*>(sh + cache_plane) : nullptr, frame_idx,
      reinterpret_cast<const float4*>(rnd), input_prob, n, (int)(frame_elems / 8), reinterpret_cast<uint4*>(oh),
      reinterpret_cast<uint4*>(oh + out_plane));
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

// ---------------------------------------------------------------------------------------------
// Fisher-weighted block losses (calib_block.py:66-72) on the same layouts.  F = cached |dL/dy| + 1 (data_utils.py:113),
// fp32 NHWC like the target cache and addressed through the same frame_idx.
//   MODE 1  fisher_diag : loss += sum d^2 F^2,               dy = 2 d F^2
//   MODE 2  fisher_full, pass A : frame_dot[b] += sum |d| F  (no gradient)
//   MODE 3  fisher_full, pass B : dy = 2 frame_dot[b] F sign(d)         (loss = sum_b frame_dot[b]^2, formed by the caller)
// blockIdx.y = batch entry, so that pass A reduces one frame per block.
// ---------------------------------------------------------------------------------------------
namespace nq {
template <int MODE>
__global__ void __launch_bounds__(256) block_fisher_kernel(const uint4* __restrict__ y_hi, const uint4* __restrict__ y_lo,
                                                           const float4* __restrict__ tgt, const float4* __restrict__ fisher,
                                                           const int* __restrict__ frame_idx, const float4* __restrict__ gprime,
                                                           int h, int w, int rh, int rw, int cg, float grad_scale,
                                                           float* __restrict__ loss_sum, float* __restrict__ frame_dot,
                                                           uint4* __restrict__ dz_hi, uint4* __restrict__ dz_lo) {
  __shared__ float red[32];
  const int c8n = cg >> 3;
  const int W2 = w * rw, H2 = h * rh;
  const int frame8 = H2 * W2 * c8n;
  const int b = blockIdx.y;
  const int src = frame_idx != nullptr ? frame_idx[b] : b;
  const float dot = MODE == 3 ? frame_dot[b] : 0.f;
  float acc = 0.f;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < frame8; f += gridDim.x * blockDim.x) {
    const size_t e = (size_t)b * frame8 + f, te = (size_t)src * frame8 + f;
    const int c8 = f % c8n;
    const int pix = f / c8n;
    const int x = pix % W2, yy = pix / W2;
    const uint4 yh = y_hi[e], yl = y_lo[e];
    const float4 t0 = tgt[2 * te], t1 = tgt[2 * te + 1];
    const float4 f0 = fisher[2 * te], f1 = fisher[2 * te + 1];
    float4 g0 = make_float4(1.f, 1.f, 1.f, 1.f), g1 = g0;
    if (MODE != 2 && gprime != nullptr) { g0 = gprime[2 * e]; g1 = gprime[2 * e + 1]; }
    const uint32_t hw[4] = {yh.x, yh.y, yh.z, yh.w}, lw[4] = {yl.x, yl.y, yl.z, yl.w};
    const float tv[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
    const float fv[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    uint32_t oh[4], ol[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gr[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint32_t hb = q ? (hw[k] & 0xFFFF0000u) : (hw[k] << 16), lb = q ? (lw[k] & 0xFFFF0000u) : (lw[k] << 16);
        const float d = (__uint_as_float(hb) + __uint_as_float(lb)) - tv[2 * k + q];
        const float F = fabsf(fv[2 * k + q]);
        float g = 0.f;
        if (MODE == 1) {
          acc += d * d * F * F;
          g = 2.0f * d * F * F;
        } else if (MODE == 2) {
          acc += fabsf(d) * F;
        } else {
          g = 2.0f * dot * F * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
        }
        gr[q] = g * grad_scale * gv[2 * k + q];
      }
      if (MODE != 2) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(gr[0]), h1 = __float2bfloat16_rn(gr[1]);
        oh[k] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        ol[k] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(gr[0] - __bfloat162float(h0))) |
                ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(gr[1] - __bfloat162float(h1))) << 16);
      }
    }
    if (MODE != 2) {
      const int qh = yy / rh, si = yy - qh * rh, qw = x / rw, sj = x - qw * rw;
      const size_t o = ((size_t)((b * h + qh) * w + qw) * (rh * rw) + (si * rw + sj)) * c8n + c8;
      dz_hi[o] = make_uint4(oh[0], oh[1], oh[2], oh[3]);
      dz_lo[o] = make_uint4(ol[0], ol[1], ol[2], ol[3]);
    }
  }
  if (MODE != 3) {
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) {
      if (MODE == 1 && loss_sum != nullptr) atomicAdd(loss_sum, acc);
      if (MODE == 2) atomicAdd(frame_dot + b, acc);
    }
  }
}
}  // namespace nq

extern "C" int nq_block_loss_bwd_fisher(const void* y_split, const float* tgt, const float* fisher, const int32_t* frame_idx,
                                        const float* gprime, int n, int h, int w, int rh, int rw, int cg, int mode,
                                        float grad_scale, float* loss_sum, float* frame_dot, void* dz_split, void* stream) {
  if (!y_split || !tgt || !fisher || !dz_split || n <= 0 || h <= 0 || w <= 0 || rh <= 0 || rw <= 0 || cg <= 0) return NQ_ERR_BAD_ARG;
  if (mode != 1 && mode != 2) return NQ_ERR_BAD_ARG;
  if (mode == 2 && !frame_dot) return NQ_ERR_BAD_ARG;
  if (cg % 8 || n > 65535) return NQ_ERR_BAD_SHAPE;
  const int64_t total = (int64_t)n * h * rh * w * rw * cg;
  if (total / 8 >= (1LL << 31)) return NQ_ERR_BAD_SHAPE;
  const int64_t frame8 = total / 8 / n;
  const uint16_t* yh = reinterpret_cast<const uint16_t*>(y_split);
  uint16_t* dh = reinterpret_cast<uint16_t*>(dz_split);
  // about two waves of 256-thread blocks over the batch
  int per_frame = (int)std::min<int64_t>((frame8 + 255) / 256, std::max<int64_t>(1, (2LL * 8 * sm_count()) / n));
  dim3 grid(per_frame, n);
  cudaStream_t st = as_stream(stream);
#define NQ_FISHER_ARGS                                                                                                         \
  reinterpret_cast<const uint4*>(yh), reinterpret_cast<const uint4*>(yh + total), reinterpret_cast<const float4*>(tgt),        \
      reinterpret_cast<const float4*>(fisher), frame_idx, reinterpret_cast<const float4*>(gprime), h, w, rh, rw, cg, grad_scale, \
      loss_sum, frame_dot, reinterpret_cast<uint4*>(dh), reinterpret_cast<uint4*>(dh + total)
  if (mode == 1) {
    block_fisher_kernel<1><<<grid, 256, 0, st>>>(NQ_FISHER_ARGS);
  } else {
    if (cudaMemsetAsync(frame_dot, 0, sizeof(float) * n, st) != cudaSuccess) return NQ_ERR_CUDA;
    block_fisher_kernel<2><<<grid, 256, 0, st>>>(NQ_FISHER_ARGS);
    NQ_LAUNCH_CHECK();
    block_fisher_kernel<3><<<grid, 256, 0, st>>>(NQ_FISHER_ARGS);
  }
#undef NQ_FISHER_ARGS
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

extern "C" int nq_block_loss_bwd(const void* y_split, const float* tgt, const int32_t* frame_idx, const float* gprime, int n, int h,
                                 int w, int rh, int rw, int cg, float p, float grad_scale, float* loss_sum, void* dz_split,
                                 void* stream) {
  if (!y_split || !tgt || !dz_split || n <= 0 || h <= 0 || w <= 0 || rh <= 0 || rw <= 0 || cg <= 0 || !(p > 0.f)) return NQ_ERR_BAD_ARG;
  if (cg % 8) return NQ_ERR_BAD_SHAPE;  // 16-byte channel groups (tensor-core stages pad cg to 16)
  const int64_t total = (int64_t)n * h * rh * w * rw * cg;
  if (total / 8 >= (1LL << 31)) return NQ_ERR_BAD_SHAPE;
  const uint16_t* yh = reinterpret_cast<const uint16_t*>(y_split);
  uint16_t* dh = reinterpret_cast<uint16_t*>(dz_split);
  block_loss_bwd_kernel<<<grid_for(total / 8), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(yh), reinterpret_cast<const uint4*>(yh + total), reinterpret_cast<const float4*>(tgt), frame_idx,
      reinterpret_cast<const float4*>(gprime), n, h, w, rh, rw, cg, p, grad_scale, loss_sum, reinterpret_cast<uint4*>(dh),
      reinterpret_cast<uint4*>(dh + total));
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

// ---------------------------------------------------------------------------------------------
// Packed weight codes (the quantised artefact, readme.md:125-127 "implementation-agnostic" step 4): integer codes
// 0 .. 2^bits - 1 (held in fp32 by the quantisers, quantizer.py:297) <-> a dense little-endian bit stream, element i in
// bits [i * bits, (i + 1) * bits).  Eight elements fill exactly `bits` bytes: one thread per group of eight.
// ---------------------------------------------------------------------------------------------
namespace nq {
__global__ void __launch_bounds__(256) pack_codes_kernel(const float* __restrict__ codes, int64_t numel, int bits,
                                                         uint8_t* __restrict__ out, int* __restrict__ bad) {
  const int64_t groups = (numel + 7) / 8;
  const float qmax = (float)((1 << bits) - 1);
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long acc = 0ull;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int64_t e = g * 8 + k;
      float v = e < numel ? codes[e] : 0.f;
      if (!(v >= 0.f && v <= qmax && v == rintf(v))) { if (bad) atomicOr(bad, 1); v = 0.f; }  // not an integer code
      acc |= (unsigned long long)(unsigned)v << (k * bits);
    }
    for (int b = 0; b < bits; ++b) out[g * bits + b] = (uint8_t)(acc >> (8 * b));
  }
}
__global__ void __launch_bounds__(256) unpack_codes_kernel(const uint8_t* __restrict__ in, int64_t numel, int bits,
                                                           float* __restrict__ codes) {
  const int64_t groups = (numel + 7) / 8;
  const unsigned mask = (1u << bits) - 1u;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long acc = 0ull;
    for (int b = 0; b < bits; ++b) acc |= (unsigned long long)in[g * bits + b] << (8 * b);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int64_t e = g * 8 + k;
      if (e < numel) codes[e] = (float)((unsigned)(acc >> (k * bits)) & mask);
    }
  }
}
}  // namespace nq

extern "C" int64_t nq_packed_bytes(int64_t numel, int n_bits) {
  if (numel < 0 || n_bits < 2 || n_bits > 8) return -1;
  return (numel + 7) / 8 * n_bits;
}
extern "C" int nq_pack_codes(const float* codes, int64_t numel, int n_bits, void* packed, int* not_integer_flag, void* stream) {
  if (!codes || !packed || numel <= 0 || n_bits < 2 || n_bits > 8) return NQ_ERR_BAD_ARG;
  pack_codes_kernel<<<grid_for((numel + 7) / 8), 256, 0, as_stream(stream)>>>(codes, numel, n_bits, reinterpret_cast<uint8_t*>(packed),
                                                                             not_integer_flag);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}
extern "C" int nq_unpack_codes(const void* packed, int64_t numel, int n_bits, float* codes, void* stream) {
  if (!codes || !packed || numel <= 0 || n_bits < 2 || n_bits > 8) return NQ_ERR_BAD_ARG;
  unpack_codes_kernel<<<grid_for((numel + 7) / 8), 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint8_t*>(packed), numel, n_bits,
                                                                               codes);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

extern "C" int nq_nchw_to_nhwc(const float* src, float* dst, int n, int c, int h, int w, int c_p, void* stream) {
  if (!src || !dst || n <= 0 || c <= 0 || h <= 0 || w <= 0 || c_p < c) return NQ_ERR_BAD_ARG;
  nchw_to_nhwc_kernel<<<grid_for((int64_t)n * h * w * c_p), 256, 0, as_stream(stream)>>>(src, dst, n, c, h, w, c_p);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}
extern "C" int nq_nhwc_to_nchw(const float* src, float* dst, int n, int c, int h, int w, int c_p, void* stream) {
  if (!src || !dst || n <= 0 || c <= 0 || h <= 0 || w <= 0 || c_p < c) return NQ_ERR_BAD_ARG;
  nhwc_to_nchw_kernel<<<grid_for((int64_t)n * c * h * w), 256, 0, as_stream(stream)>>>(src, dst, n, c, h, w, c_p);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

extern "C" int nq_lp_loss(const float* pred, const float* tgt, int64_t numel, float p, float grad_scale,
                          float* loss_sum, float* grad, void* stream) {
  if (!pred || !tgt || !loss_sum || numel <= 0 || !(p > 0.f)) return NQ_ERR_BAD_ARG;
  lp_loss_kernel<<<grid_for(numel), 256, 0, as_stream(stream)>>>(pred, tgt, numel, p, grad_scale, loss_sum, grad);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

// workspace for the two-level reductions lives in a small per-call device allocation owned by the
// caller in the general ABI; here the partial buffers are tiny (<= 16 * 148 floats), so they are
// carved out of the output-adjacent scratch the caller provides via `out` only for the final values.
// To stay stateless and allocation-free the partials use cudaMallocAsync on the given stream.
extern "C" int nq_multi_dot(const float* const* a_ptrs, const float* const* b_ptrs, const int64_t* sizes,
                            int n_tensors, int mode, float* out, void* stream) {
  if (!a_ptrs || !b_ptrs || !sizes || !out || n_tensors <= 0 || n_tensors > 16) return NQ_ERR_BAD_ARG;
  if (mode != 0 && mode != 1) return NQ_ERR_BAD_ARG;
  DotTable t;
  for (int i = 0; i < n_tensors; ++i) {
    if (!a_ptrs[i] || !b_ptrs[i] || sizes[i] <= 0) return NQ_ERR_BAD_ARG;
    t.a[i] = a_ptrs[i]; t.b[i] = b_ptrs[i]; t.n[i] = sizes[i];
  }
  cudaStream_t s = as_stream(stream);
  const int per = sm_count();
  float* partial = nullptr;
  NQ_CUDA_CHECK(cudaMallocAsync(&partial, sizeof(float) * per * n_tensors, s));
  multi_dot_partial_kernel<<<dim3(per, n_tensors), 256, 0, s>>>(t, mode, partial);
  multi_dot_final_kernel<<<n_tensors, 32, 0, s>>>(partial, per, out);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(partial, s);
  if (e != cudaSuccess) return cuda_fail(e);
  return NQ_OK;
}

extern "C" int nq_psnr(const float* a, const float* b, int n_frames, int64_t frame_numel, float* psnr, void* stream) {
  if (!a || !b || !psnr || n_frames <= 0 || frame_numel <= 0) return NQ_ERR_BAD_ARG;
  cudaStream_t s = as_stream(stream);
  int per = (int)((frame_numel + 256 * 16 - 1) / (256 * 16));
  if (per > sm_count()) per = sm_count();
  if (per < 1) per = 1;
  float* partial = nullptr;
  NQ_CUDA_CHECK(cudaMallocAsync(&partial, sizeof(float) * per * n_frames, s));
  psnr_partial_kernel<<<dim3(per, n_frames), 256, 0, s>>>(a, b, frame_numel, partial);
  psnr_final_kernel<<<n_frames, 32, 0, s>>>(partial, per, 1.0f / (float)frame_numel, psnr);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(partial, s);
  if (e != cudaSuccess) return cuda_fail(e);
  return NQ_OK;
}

// ---------------------------------------------------------------------------------------------
// Standalone activation-backward + un-shuffle (the tail of a stage whose consumer is not one of our
// dgrad kernels, e.g. a QuantNeRVBlock called on its own): dz[n,h,w,(i*rw+j)*cg+c] =
// dy[n,h*rh+i,w*rw+j,c] * act'(z[same]).
// ---------------------------------------------------------------------------------------------
namespace nq {
__global__ void __launch_bounds__(256) act_bwd_unshuffle_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                                int n, int h, int w, int rh, int rw, int cg, int act,
                                                                float* __restrict__ dz) {
  const int64_t total = (int64_t)n * h * rh * w * rw * cg;
  const int W2 = w * rw, H2 = h * rh;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % cg);
    const int64_t pix = e / cg;
    const int x = (int)(pix % W2), y = (int)((pix / W2) % H2), b = (int)(pix / ((int64_t)W2 * H2));
    float v = dy[e];
    if (act == 1) v *= gelu_grad_f(z[e]);
    else if (act == 2) v *= z[e];
    const int qh = y / rh, si = y - qh * rh, qw = x / rw, sj = x - qw * rw;
    dz[(((int64_t)b * h + qh) * w + qw) * ((int64_t)rh * rw * cg) + (int64_t)(si * rw + sj) * cg + c] = v;
  }
}
}  // namespace nq

extern "C" int nq_act_bwd_unshuffle(const float* dy, const float* z, int n, int h, int w, int rh, int rw, int cg, int act,
                                    float* dz, void* stream) {
  if (!dy || !dz || n <= 0 || h <= 0 || w <= 0 || rh <= 0 || rw <= 0 || cg <= 0) return NQ_ERR_BAD_ARG;
  if (act < 0 || act > 2) return NQ_ERR_BAD_ARG;
  if (act != 0 && !z) return NQ_ERR_BAD_ARG;
  const int64_t total = (int64_t)n * h * rh * w * rw * cg;
  act_bwd_unshuffle_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(dy, z, n, h, w, rh, rw, cg, act, dz);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

// ---------------------------------------------------------------------------------------------
// Frame ingest: uint8 frames as the data set stores them -> fp32 in [0, 1], value / 255 with an IEEE fp32 division
// (bit-identical to `read_image(...) / 255.0`, videosets/datasets.py:8-54).  16 values per thread: one 16-byte load,
// four 16-byte stores.  Used where a kernel wants fp32 targets (FFMA head, wide heads, PSNR); the default head kernel
// reads the uint8 frames directly (nq_head_fwd_loss_tapexp_u8).
// ---------------------------------------------------------------------------------------------
namespace nq {
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int64_t numel) {
  const int64_t n16 = numel >> 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float4* o = reinterpret_cast<float4*>(dst) + i * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      o[k] = make_float4(__fdiv_rn((float)(w[k] & 0xffu), 255.0f), __fdiv_rn((float)((w[k] >> 8) & 0xffu), 255.0f),
                         __fdiv_rn((float)((w[k] >> 16) & 0xffu), 255.0f), __fdiv_rn((float)(w[k] >> 24), 255.0f));
  }
  for (int64_t i = (n16 << 4) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride)
    dst[i] = __fdiv_rn((float)src[i], 255.0f);
}
}  // namespace nq

extern "C" int nq_u8_to_f32(const uint8_t* src, float* dst, int64_t numel, void* stream) {
  if (!src || !dst || numel <= 0) return NQ_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15)) return NQ_ERR_BAD_ARG;
  int64_t blocks = ((numel >> 4) + 255) / 256;
  const int64_t cap = (int64_t)nq::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  nq::u8_to_f32_kernel<<<(unsigned)blocks, 256, 0, nq::as_stream(stream)>>>(src, dst, numel);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

// ---------------------------------------------------------------------------------------------
// bit_assign as a SEARCH (methods/bit_assign.py:343-372 scores a hand-written candidate list one 15 s Hessian-vector
// product at a time).  Omega is a quadratic form, so for perturbations v_l(b) = W_l - Q_b(W_l) of layer l at b bits
//     Omega(b_0 .. b_{L-1}) = sum_l G[(l, b_l), (l, b_l)] + 2 sum_{l < m} G[(l, b_l), (m, b_m)]
// with the Gram table G[(l, b), (m, b')] = v_l(b)^T H_lm v_m(b') measured once (sensitivity.OmegaTable).  This kernel
// scores EVERY configuration -- nb^L of them, 7^7 = 823 543 for 7 layers x {2 .. 8} bits -- by table lookup, one thread
// each, and returns the admissible one (average bits <= budget) of smallest Omega; ties go to the smaller index, so the
// result does not depend on the launch geometry.
// ---------------------------------------------------------------------------------------------
namespace nq {
struct OmegaSearchParams {
  const double* G;        // (L * nb) x (L * nb), row-major, symmetric
  const double* bits_w;   // [L * nb]: n_params(l) * bits(b) / total_params  (contribution of the choice to the average bit-width)
  double budget;
  long long total;        // nb^L
  int L, nb;
  double* scores;         // optional: Omega of every configuration (total entries) or null
  double* block_best; long long* block_idx;
};

__device__ __forceinline__ double omega_of(const OmegaSearchParams& q, long long idx, double& avg_bits) {
  int c[8];
  for (int l = 0; l < q.L; ++l) { c[l] = (int)(idx % q.nb); idx /= q.nb; }
  double s = 0.0, ab = 0.0;
  const int n = q.L * q.nb;
  for (int l = 0; l < q.L; ++l) {
    const int a = l * q.nb + c[l];
    ab += q.bits_w[a];
    s += q.G[(size_t)a * n + a];
    for (int m = l + 1; m < q.L; ++m) s += 2.0 * q.G[(size_t)a * n + m * q.nb + c[m]];
  }
  avg_bits = ab;
  return s;
}

__global__ void __launch_bounds__(256) omega_search_kernel(const OmegaSearchParams q) {
  __shared__ double sb[256];
  __shared__ long long si[256];
  double best = INFINITY;
  long long bi = -1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < q.total; i += (long long)gridDim.x * blockDim.x) {
    double ab;
    const double s = omega_of(q, i, ab);
    if (q.scores) q.scores[i] = ab <= q.budget ? s : INFINITY;
    if (ab <= q.budget && (s < best || (s == best && i < bi))) { best = s; bi = i; }
  }
  sb[threadIdx.x] = best; si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const double s2 = sb[threadIdx.x + o]; const long long i2 = si[threadIdx.x + o];
      if (i2 >= 0 && (si[threadIdx.x] < 0 || s2 < sb[threadIdx.x] || (s2 == sb[threadIdx.x] && i2 < si[threadIdx.x]))) {
        sb[threadIdx.x] = s2; si[threadIdx.x] = i2;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { q.block_best[blockIdx.x] = sb[0]; q.block_idx[blockIdx.x] = si[0]; }
}

__global__ void omega_search_final_kernel(const double* block_best, const long long* block_idx, int nblocks, double* out_score,
                                          long long* out_idx) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double best = INFINITY; long long bi = -1;
  for (int b = 0; b < nblocks; ++b) {
    const long long i2 = block_idx[b];
    if (i2 >= 0 && (bi < 0 || block_best[b] < best || (block_best[b] == best && i2 < bi))) { best = block_best[b]; bi = i2; }
  }
  *out_score = best; *out_idx = bi;
}
}  // namespace nq

extern "C" int nq_omega_search_workspace(int n_layers, int n_options, int64_t* n_configs, int64_t* workspace_bytes) {
  if (n_layers < 1 || n_layers > 8 || n_options < 1 || n_options > 16 || !n_configs || !workspace_bytes) return NQ_ERR_BAD_ARG;
  long long total = 1;
  for (int l = 0; l < n_layers; ++l) total *= n_options;
  *n_configs = total;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)nq::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  *workspace_bytes = blocks * 16;
  return NQ_OK;
}

extern "C" int nq_omega_search(const double* gram, const double* bits_weight, int n_layers, int n_options, double budget,
                               double* scores, void* workspace, int64_t workspace_bytes, double* best_score,
                               int64_t* best_index, void* stream) {
  int64_t total = 0, need = 0;
  const int st = nq_omega_search_workspace(n_layers, n_options, &total, &need);
  if (st) return st;
  if (!gram || !bits_weight || !workspace || !best_score || !best_index) return NQ_ERR_BAD_ARG;
  if (workspace_bytes < need) return NQ_ERR_WORKSPACE;
  const int blocks = (int)(need / 16);
  nq::OmegaSearchParams q{};
  q.G = gram; q.bits_w = bits_weight; q.budget = budget; q.total = total; q.L = n_layers; q.nb = n_options; q.scores = scores;
  q.block_best = reinterpret_cast<double*>(workspace);
  q.block_idx = reinterpret_cast<long long*>(reinterpret_cast<uint8_t*>(workspace) + (size_t)blocks * 8);
  nq::omega_search_kernel<<<blocks, 256, 0, nq::as_stream(stream)>>>(q);
  NQ_LAUNCH_CHECK();
  nq::omega_search_final_kernel<<<1, 32, 0, nq::as_stream(stream)>>>(q.block_best, q.block_idx, blocks, best_score,
                                                                    reinterpret_cast<long long*>(best_index));
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}
