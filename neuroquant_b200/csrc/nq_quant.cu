// Quantiser kernels: UAQ 'max' scale init, fused fake-quant forward/backward (UAQ-STE, AdaRound
// soft/hard, rounding regulariser), AdaRound alpha init, Adam.  HBM-bound elementwise work:
// coalesced 1-D grids, one pass over each operand, reductions by warp shuffle.
// Reference: quantization/quantizer.py, quantization/calib_model.py:39-47.
#include "nq_common.cuh"

namespace nq {

// ---------------------------------------------------------------------------------------------
// quantizer.py:153-168 per row.  delta is formed in double from the fp32 extrema exactly as the
// reference's python floats do; zero_point = round(reciprocal(delta) * -x_min) because
// `-x_min / delta` with a python scalar on the left dispatches to Tensor.__rtruediv__.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) uaq_init_max_kernel(const float* __restrict__ x, int64_t row_len,
                                                           int n_levels, float* __restrict__ delta,
                                                           float* __restrict__ zp) {
  __shared__ float smin[8], smax[8];
  const float* row = x + (int64_t)blockIdx.x * row_len;
  float mn = INFINITY, mx = -INFINITY;
  for (int64_t i = threadIdx.x; i < row_len; i += blockDim.x) {
    const float v = row[i];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { smin[wid] = mn; smax[wid] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) { mn = fminf(mn, smin[i]); mx = fmaxf(mx, smax[i]); }
    const double x_min = fmin((double)mn, 0.0), x_max = fmax((double)mx, 0.0);
    float d = (float)((x_max - x_min) / (double)(n_levels - 1));
    d = fmaxf(d, 1e-8f);
    const float rcp = __fdiv_rn(1.0f, d);
    delta[blockIdx.x] = d;
    zp[blockIdx.x] = rintf(__fmul_rn(rcp, (float)(-x_min)));
  }
}

// 'mse' (L_3.5) / 'l1' range search and the 'gaussian' initialiser (quantizer.py:170-222), one block per row.  The ten
// candidate ranges are scored in ONE pass over the row; every operation the reference performs on fp32 tensors is a
// separately rounded fp32 operation here (no contraction), so the step sizes are the reference's bit for bit -- only the
// summation order of the score differs, which matters when two candidates tie to ~1e-7.
template <int METHOD>  // 1 mse, 2 l1, 3 gaussian
__global__ void __launch_bounds__(256) uaq_init_search_kernel(const float* __restrict__ x, int64_t row_len, int n_bits,
                                                              float* __restrict__ delta, float* __restrict__ zp) {
  __shared__ float smin[8], smax[8];
  __shared__ double ssum[8][10];
  __shared__ float bc[2];
  const float* row = x + (int64_t)blockIdx.x * row_len;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float qmax = (float)((1 << n_bits) - 1);
  if (METHOD == 3) {
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < row_len; i += blockDim.x) acc += (double)row[i];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) ssum[wid][0] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < nw; ++i) t += ssum[i][0];
      bc[0] = (float)(t / (double)row_len);
    }
    __syncthreads();
    const double mean = (double)bc[0];
    acc = 0.0;
    for (int64_t i = threadIdx.x; i < row_len; i += blockDim.x) { const double dlt = (double)row[i] - mean; acc += dlt * dlt; }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) ssum[wid][1] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < nw; ++i) t += ssum[i][1];
      const float mu = bc[0], var = (float)(t / (double)(row_len - 1));      // torch.var: unbiased
      const float six = __fmul_rn(6.0f, var);
      const float lo = fminf(__fsub_rn(mu, six), 0.0f), hi = fmaxf(__fadd_rn(mu, six), 0.0f);
      const float d = fmaxf(__fdiv_rn(__fsub_rn(hi, lo), qmax), 1e-8f);
      delta[blockIdx.x] = d;
      zp[blockIdx.x] = rintf(__fdiv_rn(-lo, d));
    }
    return;
  }
  float mn = INFINITY, mx = -INFINITY;
  for (int64_t i = threadIdx.x; i < row_len; i += blockDim.x) {
    const float v = row[i];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  if (lane == 0) { smin[wid] = mn; smax[wid] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < nw; ++i) { mn = fminf(mn, smin[i]); mx = fmaxf(mx, smax[i]); }
    bc[0] = mn; bc[1] = mx;
  }
  __syncthreads();
  mn = bc[0]; mx = bc[1];
  float d[10], z[10];
  double sc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const float f = (float)(1.0 - (double)i * 0.05);            // python float, cast when it meets the fp32 tensor
    const float hi = __fmul_rn(mx, f), lo = __fmul_rn(mn, f);
    d[i] = fmaxf(__fdiv_rn(__fsub_rn(hi, lo), qmax), 1e-8f);
    z[i] = rintf(__fdiv_rn(-lo, d[i]));
    sc[i] = 0.0;
  }
  for (int64_t e = threadIdx.x; e < row_len; e += blockDim.x) {
    const float v = row[e];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const float q = fminf(fmaxf(__fadd_rn(rintf(__fdiv_rn(v, d[i])), z[i]), 0.0f), qmax);
      const float err = fabsf(__fsub_rn(v, __fmul_rn(__fsub_rn(q, z[i]), d[i])));
      sc[i] += (double)(METHOD == 1 ? powf(err, 3.5f) : err);
    }
  }
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    double a = sc[i];
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) ssum[wid][i] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float best = 1e+10f;
    int pick = 0;
    for (int i = 0; i < 10; ++i) {
      double t = 0.0;
      for (int w = 0; w < nw; ++w) t += ssum[w][i];
      const float score = (float)(t / (double)row_len);
      if (score < best) { best = score; pick = i; }            // strict: the first of equal scores wins
    }
    delta[blockIdx.x] = d[pick];
    zp[blockIdx.x] = z[pick];
  }
}

__device__ __forceinline__ float soft_target_raw(float alpha, float& sig) {
  sig = sigmoid_f(alpha);
  return __fadd_rn(__fmul_rn(sig, kZeta - kGamma), kGamma);  // two roundings, as torch (no FMA)
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// one element of the fake-quantiser forward: returns the code, writes the rounding-regulariser term
template <int MODE>
__device__ __forceinline__ float fq_fwd_elem(float xv, float av, float d, float z, float qmax, bool want_reg, float reg_b,
                                             float& reg) {
  const float q = __fdiv_rn(xv, d);  // true fp32 divide
  float x_int;
  if (MODE == NQ_ROUND_NEAREST) {
    x_int = rintf(q);  // half-to-even == torch.round
  } else if (MODE == NQ_ROUND_SOFT) {
    float sig;
    const float h = fminf(fmaxf(soft_target_raw(av, sig), 0.f), 1.f);
    x_int = __fadd_rn(floorf(q), h);
    if (want_reg) {
      const float u = fabsf(h - 0.5f) * 2.0f;
      reg += 1.0f - powf(u, reg_b);
    }
  } else {
    x_int = __fadd_rn(floorf(q), av >= 0.f ? 1.f : 0.f);
  }
  return fminf(fmaxf(__fadd_rn(x_int, z), 0.f), qmax);
}

template <int MODE>
__global__ void __launch_bounds__(256) fakequant_fwd_kernel(
    const float* __restrict__ x, const float* __restrict__ alpha, const float* __restrict__ delta,
    const float* __restrict__ zp, int64_t numel, int row_len, int d_stride, float qmax,
    float* __restrict__ codes, float* __restrict__ deq, float* __restrict__ reg_sum, float reg_b) {
  __shared__ float red[32];
  float reg = 0.f;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < numel;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = (e / row_len) * d_stride;
    const float d = delta[r], z = zp[r];
    const float c = fq_fwd_elem<MODE>(x[e], MODE == NQ_ROUND_NEAREST ? 0.f : alpha[e], d, z, qmax, reg_sum != nullptr, reg_b, reg);
    if (codes != nullptr) codes[e] = c;
    if (deq != nullptr) deq[e] = __fmul_rn(__fsub_rn(c, z), d);
  }
  if (MODE == NQ_ROUND_SOFT && reg_sum != nullptr) {
    reg = block_sum(reg, red);
    if (threadIdx.x == 0) atomicAdd(reg_sum, reg);
  }
}

// ---------------------------------------------------------------------------------------------
// backward, AdaRound soft: elementwise d_alpha
// ---------------------------------------------------------------------------------------------
// one element of d(loss + regulariser)/d(alpha); the last product is an explicit rounding so that fusing the
// consumer (Adam) behind it cannot change the value
__device__ __forceinline__ float fq_bwd_soft_elem(float gv, float xv, float av, float d, float z, float qmax, float grad_scale,
                                                  float reg_w, float reg_b) {
  float sig;
  const float hraw = soft_target_raw(av, sig);
  const bool pass_h = (hraw >= 0.f) && (hraw <= 1.f);  // clamp backward is inclusive
  const float h = fminf(fmaxf(hraw, 0.f), 1.f);
  const float v = __fadd_rn(__fadd_rn(floorf(__fdiv_rn(xv, d)), h), z);
  const bool in_range = (v >= 0.f) && (v <= qmax);
  float dh = 0.f;
  if (in_range) dh = gv * grad_scale * d;
  if (reg_w != 0.f) {
    const float t = h - 0.5f;
    const float u = fabsf(t) * 2.0f;
    const float sgn = (t > 0.f) ? 1.f : ((t < 0.f) ? -1.f : 0.f);
    // d/dh (1 - u^b) = -b u^(b-1) * 2 sign(h - .5)
    dh += reg_w * (-reg_b * powf(u, reg_b - 1.0f) * 2.0f * sgn);
  }
  return pass_h ? __fmul_rn(dh * (kZeta - kGamma) * sig, 1.0f - sig) : 0.f;
}

// torch.optim.Adam, single-tensor path (torch/optim/adam.py _single_tensor_adam), one element
__device__ __forceinline__ void adam_elem(float& pv, float gv, float& mv_io, float& vv_io, float one_minus_b1, float b2,
                                          float one_minus_b2, float step_size, float bc2_sqrt, float eps) {
  const float mv = mv_io + one_minus_b1 * (gv - mv_io);  // lerp_
  const float vv = __fadd_rn(__fmul_rn(vv_io, b2), __fmul_rn(__fmul_rn(gv, gv), one_minus_b2));
  mv_io = mv;
  vv_io = vv;
  const float denom = __fadd_rn(__fdiv_rn(sqrtf(vv), bc2_sqrt), eps);
  pv = pv - step_size * __fdiv_rn(mv, denom);
}

__global__ void __launch_bounds__(256) fakequant_bwd_soft_kernel(
    const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ alpha,
    const float* __restrict__ delta, const float* __restrict__ zp, int64_t numel, int row_len,
    int d_stride, float qmax, float grad_scale, float reg_w, float reg_b, float* __restrict__ d_alpha,
    const float* __restrict__ hyper) {
  if (hyper != nullptr) {  // CUDA-graph replay: the schedule values live in device memory
    reg_w = hyper[0];
    reg_b = hyper[1];
  }
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < numel;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = (e / row_len) * d_stride;
    d_alpha[e] = fq_bwd_soft_elem(g[e], x[e], alpha[e], delta[r], zp[r], qmax, grad_scale, reg_w, reg_b);
  }
}

// ---------------------------------------------------------------------------------------------
// backward, UAQ-STE: per-row d_delta (one CTA per row; rows share a scale only when d_stride == 1,
// the per-tensor case uses a single CTA over the whole tensor)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fakequant_bwd_delta_kernel(
    const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ delta,
    const float* __restrict__ zp, int64_t row_len, float qmax, float grad_scale,
    float* __restrict__ d_delta) {
  __shared__ float red[32];
  const int64_t base = (int64_t)blockIdx.x * row_len;
  const float d = delta[blockIdx.x], z = zp[blockIdx.x];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < row_len; i += blockDim.x) {
    const float xv = x[base + i];
    const float q = __fdiv_rn(xv, d);
    const float v = __fadd_rn(rintf(q), z);
    const bool in_range = (v >= 0.f) && (v <= qmax);
    const float c = fminf(fmaxf(v, 0.f), qmax);
    acc += g[base + i] * ((c - z) - (in_range ? q : 0.f));
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) d_delta[blockIdx.x] = acc * grad_scale;
}

__global__ void __launch_bounds__(256) adaround_init_alpha_kernel(
    const float* __restrict__ x, const float* __restrict__ delta, int64_t numel, int row_len,
    int d_stride, float* __restrict__ alpha) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < numel;
       e += (int64_t)gridDim.x * blockDim.x) {
    const float d = delta[(e / row_len) * d_stride];
    const float q = __fdiv_rn(x[e], d);
    const float rest = __fsub_rn(q, floorf(q));
    // -log((zeta - gamma) / (rest - gamma) - 1)
    const float t = __fsub_rn(__fdiv_rn(kZeta - kGamma, __fsub_rn(rest, kGamma)), 1.0f);
    alpha[e] = -logf(t);
  }
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                   float one_minus_b1, float b2, float one_minus_b2,
                                                   float step_size, float bc2_sqrt, float eps,
                                                   const float* __restrict__ hyper) {
  if (hyper != nullptr) {
    step_size = hyper[2];
    bc2_sqrt = hyper[3];
  }
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) {
    float pv = p[e], mv = m[e], vv = v[e];
    adam_elem(pv, g[e], mv, vv, one_minus_b1, b2, one_minus_b2, step_size, bc2_sqrt, eps);
    m[e] = mv;
    v[e] = vv;
    p[e] = pv;
  }
}

// ---------------------------------------------------------------------------------------------
// Multi-tensor launches: the decoder has 7 stages x (weight, bias) quantisers of 12 .. 1.6 M elements; one launch
// each is launch-latency bound (4-5 us per kernel for < 1 us of work).  A block owns MULTI_CHUNK consecutive
// elements of one tensor; the block -> tensor table travels in the kernel parameters.
// ---------------------------------------------------------------------------------------------
constexpr int MULTI_CHUNK = 2048;  // elements per block (256 threads x 8)

struct FqMulti {
  nq_fq_task t[NQ_MULTI_MAX];
  int blk_start[NQ_MULTI_MAX + 1];
  int n;
  float* reg_sum;
  float reg_b;
};
struct AdaMulti {
  nq_ada_task t[NQ_MULTI_MAX];
  int blk_start[NQ_MULTI_MAX + 1];
  int n;
  float grad_scale, one_minus_b1, b2, one_minus_b2, eps;
  const float* hyper;
};

template <typename M>
__device__ __forceinline__ int multi_task_of_block(const M& m, int blk) {
  int t = 0;
  while (t + 1 < m.n && blk >= m.blk_start[t + 1]) ++t;
  return t;
}

__global__ void __launch_bounds__(256) fakequant_fwd_multi_kernel(const __grid_constant__ FqMulti m) {
  __shared__ float red[32];
  const int ti = multi_task_of_block(m, blockIdx.x);
  const nq_fq_task& t = m.t[ti];
  const int64_t numel = t.rows * t.row_len;
  const int64_t e0 = (int64_t)(blockIdx.x - m.blk_start[ti]) * MULTI_CHUNK;
  const int64_t e1 = e0 + MULTI_CHUNK < numel ? e0 + MULTI_CHUNK : numel;
  const float qmax = (float)((1 << t.n_bits) - 1);
  const bool want_reg = t.want_reg && m.reg_sum != nullptr && t.mode == NQ_ROUND_SOFT;
  float reg = 0.f;
  // 16-byte path: rows of a multiple of 4 elements (4 consecutive elements share their channel's step size), aligned tensors;
  // one 32-bit division per 4 elements.  Same element function, same values.
  const bool vec = (t.row_len & 3) == 0 && numel < (1LL << 31) &&
                   ((((uintptr_t)t.x | (uintptr_t)t.alpha | (uintptr_t)t.codes | (uintptr_t)t.deq) & 15) == 0);
  if (vec) {
    const int rl = (int)t.row_len;
    for (int e = (int)e0 + 4 * threadIdx.x; e < (int)e1; e += 4 * blockDim.x) {
      const int r = t.channel_wise ? e / rl : 0;
      const float d = t.delta[r], z = t.zero_point[r];
      const float4 xv = *reinterpret_cast<const float4*>(t.x + e);
      float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t.mode != NQ_ROUND_NEAREST) av = *reinterpret_cast<const float4*>(t.alpha + e);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, as[4] = {av.x, av.y, av.z, av.w};
      float c[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (t.mode == NQ_ROUND_NEAREST) c[k] = fq_fwd_elem<NQ_ROUND_NEAREST>(xs[k], 0.f, d, z, qmax, false, 0.f, reg);
        else if (t.mode == NQ_ROUND_SOFT) c[k] = fq_fwd_elem<NQ_ROUND_SOFT>(xs[k], as[k], d, z, qmax, want_reg, m.reg_b, reg);
        else c[k] = fq_fwd_elem<NQ_ROUND_HARD>(xs[k], as[k], d, z, qmax, false, 0.f, reg);
      }
      if (t.codes != nullptr) *reinterpret_cast<float4*>(t.codes + e) = make_float4(c[0], c[1], c[2], c[3]);
      if (t.deq != nullptr)
        *reinterpret_cast<float4*>(t.deq + e) = make_float4(__fmul_rn(__fsub_rn(c[0], z), d), __fmul_rn(__fsub_rn(c[1], z), d),
                                                             __fmul_rn(__fsub_rn(c[2], z), d), __fmul_rn(__fsub_rn(c[3], z), d));
    }
  } else
  for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
    const int64_t r = t.channel_wise ? e / t.row_len : 0;
    const float d = t.delta[r], z = t.zero_point[r];
    const float xv = t.x[e];
    float c;
    if (t.mode == NQ_ROUND_NEAREST) c = fq_fwd_elem<NQ_ROUND_NEAREST>(xv, 0.f, d, z, qmax, false, 0.f, reg);
    else if (t.mode == NQ_ROUND_SOFT) c = fq_fwd_elem<NQ_ROUND_SOFT>(xv, t.alpha[e], d, z, qmax, want_reg, m.reg_b, reg);
    else c = fq_fwd_elem<NQ_ROUND_HARD>(xv, t.alpha[e], d, z, qmax, false, 0.f, reg);
    if (t.codes != nullptr) t.codes[e] = c;
    if (t.deq != nullptr) t.deq[e] = __fmul_rn(__fsub_rn(c, z), d);
  }
  if (want_reg) {  // block-uniform
    reg = block_sum(reg, red);
    if (threadIdx.x == 0) atomicAdd(m.reg_sum, reg);
  }
}

// d_alpha (fakequant_bwd_soft_kernel) and the Adam update of alpha (adam_kernel) in one pass: d_alpha never
// leaves the registers.  Same element functions as the single-tensor kernels: identical values.
__global__ void __launch_bounds__(256) adaround_step_multi_kernel(const __grid_constant__ AdaMulti m) {
  const int ti = multi_task_of_block(m, blockIdx.x);
  const nq_ada_task& t = m.t[ti];
  const int64_t numel = t.rows * t.row_len;
  const int64_t e0 = (int64_t)(blockIdx.x - m.blk_start[ti]) * MULTI_CHUNK;
  const int64_t e1 = e0 + MULTI_CHUNK < numel ? e0 + MULTI_CHUNK : numel;
  const float qmax = (float)((1 << t.n_bits) - 1);
  const float reg_w = t.use_reg ? m.hyper[0] : 0.f, reg_b = t.use_reg ? m.hyper[1] : 0.f;
  const float step_size = m.hyper[2], bc2_sqrt = m.hyper[3];
  const bool vec = (t.row_len & 3) == 0 && numel < (1LL << 31) &&
                   ((((uintptr_t)t.g | (uintptr_t)t.x | (uintptr_t)t.alpha | (uintptr_t)t.exp_avg | (uintptr_t)t.exp_avg_sq) & 15) == 0);
  if (vec) {  // 16-byte path, see fakequant_fwd_multi_kernel
    const int rl = (int)t.row_len;
    for (int e = (int)e0 + 4 * threadIdx.x; e < (int)e1; e += 4 * blockDim.x) {
      const int r = t.channel_wise ? e / rl : 0;
      const float d = t.delta[r], z = t.zero_point[r];
      const float4 g4 = *reinterpret_cast<const float4*>(t.g + e), x4 = *reinterpret_cast<const float4*>(t.x + e);
      float4 a4 = *reinterpret_cast<const float4*>(t.alpha + e), m4 = *reinterpret_cast<const float4*>(t.exp_avg + e),
             v4 = *reinterpret_cast<const float4*>(t.exp_avg_sq + e);
      float* ap = &a4.x; float* mp = &m4.x; float* vp = &v4.x;
      const float* gp = &g4.x; const float* xp = &x4.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gv = fq_bwd_soft_elem(gp[k], xp[k], ap[k], d, z, qmax, m.grad_scale, reg_w, reg_b);
        adam_elem(ap[k], gv, mp[k], vp[k], m.one_minus_b1, m.b2, m.one_minus_b2, step_size, bc2_sqrt, m.eps);
      }
      *reinterpret_cast<float4*>(t.exp_avg + e) = m4;
      *reinterpret_cast<float4*>(t.exp_avg_sq + e) = v4;
      *reinterpret_cast<float4*>(t.alpha + e) = a4;
    }
  } else
  for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
    const int64_t r = t.channel_wise ? e / t.row_len : 0;
    float av = t.alpha[e], mv = t.exp_avg[e], vv = t.exp_avg_sq[e];
    const float gv = fq_bwd_soft_elem(t.g[e], t.x[e], av, t.delta[r], t.zero_point[r], qmax, m.grad_scale, reg_w, reg_b);
    adam_elem(av, gv, mv, vv, m.one_minus_b1, m.b2, m.one_minus_b2, step_size, bc2_sqrt, m.eps);
    t.exp_avg[e] = mv;
    t.exp_avg_sq[e] = vv;
    t.alpha[e] = av;
  }
}

static inline int grid_for(int64_t numel, int block = 256) {
  int64_t b = (numel + block - 1) / block;
  const int64_t cap = (int64_t)sm_count() * 16;  // multiple of the SM count, grid-stride beyond
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace nq

using namespace nq;

extern "C" int nq_uaq_init_max(const float* x, int64_t rows, int64_t row_len, int n_bits, float* delta,
                               float* zero_point, void* stream) {
  if (!x || !delta || !zero_point || rows <= 0 || row_len <= 0) return NQ_ERR_BAD_ARG;
  if (n_bits < 2 || n_bits > 8) return NQ_ERR_BAD_ARG;
  uaq_init_max_kernel<<<(unsigned)rows, 256, 0, as_stream(stream)>>>(x, row_len, 1 << n_bits, delta, zero_point);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

extern "C" int nq_uaq_init_search(const float* x, int64_t rows, int64_t row_len, int n_bits, int method, float* delta,
                                  float* zero_point, void* stream) {
  if (!x || !delta || !zero_point || rows <= 0 || row_len <= 0) return NQ_ERR_BAD_ARG;
  if (n_bits < 2 || n_bits > 8 || method < 1 || method > 3) return NQ_ERR_BAD_ARG;
  cudaStream_t s = as_stream(stream);
  if (method == 1) uaq_init_search_kernel<1><<<(unsigned)rows, 256, 0, s>>>(x, row_len, n_bits, delta, zero_point);
  else if (method == 2) uaq_init_search_kernel<2><<<(unsigned)rows, 256, 0, s>>>(x, row_len, n_bits, delta, zero_point);
  else uaq_init_search_kernel<3><<<(unsigned)rows, 256, 0, s>>>(x, row_len, n_bits, delta, zero_point);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

extern "C" int nq_fakequant_fwd(const float* x, const float* alpha, const float* delta, const float* zero_point,
                                int64_t rows, int64_t row_len, int d_stride, int n_bits, int mode, float* codes,
                                float* deq, float* reg_sum, float reg_b, void* stream) {
  if (!x || !delta || !zero_point || rows <= 0 || row_len <= 0) return NQ_ERR_BAD_ARG;
  if (n_bits < 2 || n_bits > 8 || (d_stride != 0 && d_stride != 1)) return NQ_ERR_BAD_ARG;
  if (mode != NQ_ROUND_NEAREST && !alpha) return NQ_ERR_BAD_ARG;
  if (row_len > 0x7fffffffLL) return NQ_ERR_BAD_SHAPE;
  const int64_t numel = rows * row_len;
  const float qmax = (float)((1 << n_bits) - 1);
  const int grid = grid_for(numel);
  cudaStream_t s = as_stream(stream);
  switch (mode) {
    case NQ_ROUND_NEAREST:
      fakequant_fwd_kernel<NQ_ROUND_NEAREST><<<grid, 256, 0, s>>>(x, alpha, delta, zero_point, numel, (int)row_len,
                                                                  d_stride, qmax, codes, deq, nullptr, 0.f);
      break;
    case NQ_ROUND_SOFT:
      fakequant_fwd_kernel<NQ_ROUND_SOFT><<<grid, 256, 0, s>>>(x, alpha, delta, zero_point, numel, (int)row_len,
                                                               d_stride, qmax, codes, deq, reg_sum, reg_b);
      break;
    case NQ_ROUND_HARD:
      fakequant_fwd_kernel<NQ_ROUND_HARD><<<grid, 256, 0, s>>>(x, alpha, delta, zero_point, numel, (int)row_len,
                                                               d_stride, qmax, codes, deq, nullptr, 0.f);
      break;
    default:
      return NQ_ERR_BAD_ARG;
  }
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

extern "C" int nq_fakequant_bwd(const float* g, const float* x, const float* alpha, const float* delta,
                                const float* zero_point, int64_t rows, int64_t row_len, int d_stride, int n_bits,
                                int mode, float grad_scale, float reg_w, float reg_b, float* d_alpha,
                                float* d_delta, void* stream) {
  if (!g || !x || !delta || !zero_point || rows <= 0 || row_len <= 0) return NQ_ERR_BAD_ARG;
  if (n_bits < 2 || n_bits > 8 || (d_stride != 0 && d_stride != 1)) return NQ_ERR_BAD_ARG;
  if (row_len > 0x7fffffffLL) return NQ_ERR_BAD_SHAPE;
  const float qmax = (float)((1 << n_bits) - 1);
  cudaStream_t s = as_stream(stream);
  if (mode == NQ_ROUND_SOFT) {
    if (!alpha || !d_alpha) return NQ_ERR_BAD_ARG;
    const int64_t numel = rows * row_len;
    fakequant_bwd_soft_kernel<<<grid_for(numel), 256, 0, s>>>(g, x, alpha, delta, zero_point, numel, (int)row_len,
                                                             d_stride, qmax, grad_scale, reg_w, reg_b, d_alpha, nullptr);
  } else if (mode == NQ_ROUND_NEAREST) {
    if (!d_delta) return NQ_ERR_BAD_ARG;
    if (d_stride == 1)
      fakequant_bwd_delta_kernel<<<(unsigned)rows, 256, 0, s>>>(g, x, delta, zero_point, row_len, qmax, grad_scale, d_delta);
    else  // one shared scale: a single row spanning the tensor
      fakequant_bwd_delta_kernel<<<1, 256, 0, s>>>(g, x, delta, zero_point, rows * row_len, qmax, grad_scale, d_delta);
  } else {
    return NQ_ERR_BAD_ARG;  // hard rounding has no learnable
  }
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

extern "C" int nq_adaround_init_alpha(const float* x, const float* delta, int64_t rows, int64_t row_len,
                                      int d_stride, float* alpha, void* stream) {
  if (!x || !delta || !alpha || rows <= 0 || row_len <= 0) return NQ_ERR_BAD_ARG;
  if (d_stride != 0 && d_stride != 1) return NQ_ERR_BAD_ARG;
  if (row_len > 0x7fffffffLL) return NQ_ERR_BAD_SHAPE;
  const int64_t numel = rows * row_len;
  adaround_init_alpha_kernel<<<grid_for(numel), 256, 0, as_stream(stream)>>>(x, delta, numel, (int)row_len, d_stride, alpha);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

extern "C" int nq_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                            double lr, double beta1, double beta2, double eps, int step, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || n <= 0 || step < 1) return NQ_ERR_BAD_ARG;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  adam_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, (float)(1.0 - beta1),
                                                          (float)beta2, (float)(1.0 - beta2), step_size, bc2_sqrt,
                                                          (float)eps, nullptr);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

// Variants for CUDA-graph capture of the calibration iteration: the per-iteration scalars (regulariser weight and
// temperature, Adam's bias-corrected step size and sqrt(1 - beta2^t)) are read from hyper_dev[0..3] at run time.
extern "C" int nq_fakequant_bwd_soft_dev(const float* g, const float* x, const float* alpha, const float* delta,
                                         const float* zero_point, int64_t rows, int64_t row_len, int d_stride,
                                         int n_bits, float grad_scale, int use_reg, const float* hyper_dev,
                                         float* d_alpha, void* stream) {
  if (!g || !x || !alpha || !delta || !zero_point || !d_alpha || !hyper_dev || rows <= 0 || row_len <= 0) return NQ_ERR_BAD_ARG;
  if (n_bits < 2 || n_bits > 8 || (d_stride != 0 && d_stride != 1)) return NQ_ERR_BAD_ARG;
  if (row_len > 0x7fffffffLL) return NQ_ERR_BAD_SHAPE;
  const int64_t numel = rows * row_len;
  const float qmax = (float)((1 << n_bits) - 1);
  if (use_reg)
    fakequant_bwd_soft_kernel<<<grid_for(numel), 256, 0, as_stream(stream)>>>(g, x, alpha, delta, zero_point, numel, (int)row_len,
                                                                             d_stride, qmax, grad_scale, 0.f, 0.f, d_alpha, hyper_dev);
  else
    fakequant_bwd_soft_kernel<<<grid_for(numel), 256, 0, as_stream(stream)>>>(g, x, alpha, delta, zero_point, numel, (int)row_len,
                                                                             d_stride, qmax, grad_scale, 0.f, 0.f, d_alpha, nullptr);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

extern "C" int nq_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                double beta1, double beta2, double eps, const float* hyper_dev, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !hyper_dev || n <= 0) return NQ_ERR_BAD_ARG;
  adam_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, (float)(1.0 - beta1),
                                                          (float)beta2, (float)(1.0 - beta2), 0.f, 1.f, (float)eps, hyper_dev);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

extern "C" int nq_fakequant_fwd_multi(const nq_fq_task* tasks, int n_tasks, float* reg_sum, float reg_b, void* stream) {
  if (!tasks || n_tasks <= 0) return NQ_ERR_BAD_ARG;
  for (int i0 = 0; i0 < n_tasks; i0 += NQ_MULTI_MAX) {
    FqMulti m{};
    m.n = n_tasks - i0 < NQ_MULTI_MAX ? n_tasks - i0 : NQ_MULTI_MAX;
    m.reg_sum = reg_sum;
    m.reg_b = reg_b;
    int blocks = 0;
    for (int i = 0; i < m.n; ++i) {
      const nq_fq_task& t = tasks[i0 + i];
      if (!t.x || !t.delta || !t.zero_point || t.rows <= 0 || t.row_len <= 0) return NQ_ERR_BAD_ARG;
      if (t.n_bits < 2 || t.n_bits > 8 || t.mode < NQ_ROUND_NEAREST || t.mode > NQ_ROUND_HARD) return NQ_ERR_BAD_ARG;
      if (t.mode != NQ_ROUND_NEAREST && !t.alpha) return NQ_ERR_BAD_ARG;
      if (t.row_len > 0x7fffffffLL) return NQ_ERR_BAD_SHAPE;
      m.t[i] = t;
      m.blk_start[i] = blocks;
      blocks += (int)((t.rows * t.row_len + MULTI_CHUNK - 1) / MULTI_CHUNK);
    }
    m.blk_start[m.n] = blocks;
    fakequant_fwd_multi_kernel<<<blocks, 256, 0, as_stream(stream)>>>(m);
    NQ_LAUNCH_CHECK();
  }
  return NQ_OK;
}

extern "C" int nq_adaround_step_multi(const nq_ada_task* tasks, int n_tasks, float grad_scale, double beta1, double beta2,
                                      double eps, const float* hyper_dev, void* stream) {
  if (!tasks || n_tasks <= 0 || !hyper_dev) return NQ_ERR_BAD_ARG;
  for (int i0 = 0; i0 < n_tasks; i0 += NQ_MULTI_MAX) {
    AdaMulti m{};
    m.n = n_tasks - i0 < NQ_MULTI_MAX ? n_tasks - i0 : NQ_MULTI_MAX;
    m.grad_scale = grad_scale;
    m.one_minus_b1 = (float)(1.0 - beta1);
    m.b2 = (float)beta2;
    m.one_minus_b2 = (float)(1.0 - beta2);
    m.eps = (float)eps;
    m.hyper = hyper_dev;
    int blocks = 0;
    for (int i = 0; i < m.n; ++i) {
      const nq_ada_task& t = tasks[i0 + i];
      if (!t.g || !t.x || !t.alpha || !t.delta || !t.zero_point || !t.exp_avg || !t.exp_avg_sq) return NQ_ERR_BAD_ARG;
      if (t.rows <= 0 || t.row_len <= 0 || t.n_bits < 2 || t.n_bits > 8) return NQ_ERR_BAD_ARG;
      if (t.row_len > 0x7fffffffLL) return NQ_ERR_BAD_SHAPE;
      m.t[i] = t;
      m.blk_start[i] = blocks;
      blocks += (int)((t.rows * t.row_len + MULTI_CHUNK - 1) / MULTI_CHUNK);
    }
    m.blk_start[m.n] = blocks;
    adaround_step_multi_kernel<<<blocks, 256, 0, as_stream(stream)>>>(m);
    NQ_LAUNCH_CHECK();
  }
  return NQ_OK;
}
