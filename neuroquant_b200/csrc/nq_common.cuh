// Shared device/host helpers for libnq_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include <stdint.h>
#include <math.h>
#include "../../include/neuroquant_b200.h"

namespace nq {

extern thread_local int g_last_cuda_error;

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = (int)e;
  return NQ_ERR_CUDA;
}

#define NQ_CUDA_CHECK(expr)                               \
  do {                                                    \
    cudaError_t _e = (expr);                              \
    if (_e != cudaSuccess) return ::nq::cuda_fail(_e);    \
  } while (0)

#define NQ_LAUNCH_CHECK() NQ_CUDA_CHECK(cudaGetLastError())

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();  // cached per device

// Programmatic dependent launch.  The persistent tensor-core kernels spend their first microseconds on chip only (barrier
// initialisation, TMEM allocation, constant operand fill); launched with the programmatic-serialisation attribute their
// CTAs take over an SM as soon as a CTA of the kernel before them exits and do that part under its tail.  pdl_wait()
// returns when the preceding kernel has completed and its writes are visible -- nothing before it may touch global
// memory; without the attribute it returns at once.  pdl_trigger() lets the NEXT kernel's CTAs be scheduled.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
inline bool pdl_enabled() {  // NQ_PDL=0: plain stream order
  static const bool on = !(getenv("NQ_PDL") && atoi(getenv("NQ_PDL")) == 0);
  return on;
}

constexpr float kZeta = 1.1f, kGamma = -0.1f;  // quantizer.py:274

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; result valid in thread 0.  `red` is >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    v = lane < nw ? red[lane] : 0.f;
    v = warp_sum(v);
  }
  return v;
}

// exact-erf GELU and its derivative (nn.GELU default, models/_layers.py:105)
__device__ __forceinline__ float gelu_f(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Branch-free GELU / GELU' for the tensor-core epilogues.  Phi(x) = 0.5 erfc(-x/sqrt2) from the
// Abramowitz-Stegun 7.1.26 rational (|abs err| <= 1.5e-7 on erfc), evaluated on |x| and mirrored so
// that the small tail is never formed by cancellation; GELU' reuses the same exponential.
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& e) {
  const float u = fabsf(x) * 0.70710678118654752440f;
  float t;  // one MUFU.RCP (1 ulp) instead of the ~8-instruction IEEE reciprocal: far inside the rational's own error
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, u, 1.0f)));
  e = __expf(-u * u);  // = exp(-x^2 / 2)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float half_erfc = 0.5f * poly * t * e;  // 0.5 erfc(|x|/sqrt2)
  cdf = x < 0.f ? half_erfc : 1.0f - half_erfc;
}
__device__ __forceinline__ float gelu_fast(float x) {
  float cdf, e;
  gelu_parts(x, cdf, e);
  return x * cdf;
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
  float cdf, e;
  gelu_parts(x, cdf, e);
  return fmaf(x * 0.39894228040143267794f, e, cdf);
}

// GELU and GELU' of the same argument from one exponential (forward epilogue that keeps the derivative)
__device__ __forceinline__ void gelu_both_fast(float x, float& y, float& g) {
  float cdf, e;
  gelu_parts(x, cdf, e);
  y = x * cdf;
  g = fmaf(x * 0.39894228040143267794f, e, cdf);
}

__device__ __forceinline__ float sigmoid_f(float a) { return 1.0f / (1.0f + expf(-a)); }

}  // namespace nq
