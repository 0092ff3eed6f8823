// Second-order forward-mode ("jet") kernels for the Omega sensitivity of bit_assign.
// Reference: methods/bit_assign.py:57-118,171-203 computes Omega = v^T H v with a double backward pass;
// here the same number is d^2/d eps^2 L(w + eps v) at eps = 0, obtained by propagating (y, y', y'')
// FORWARD through the decoder: convolutions are bilinear in (input, weight), so each stage
// needs only forward convolutions (run by the tensor-core conv kernel) plus the elementwise chain rule
// below.  Nothing is differentiated backwards.
#include "nq_common.cuh"

namespace nq {

// y = f(z), y' = f'(z) z', y'' = f''(z) z'^2 + f'(z) z''  with  z' = zd1 + zd2,  z'' = zdd1 + 2 zdd2
// (zd1 = conv(x'; w), zd2 = conv(x; v), zdd1 = conv(x''; w), zdd2 = conv(x'; v); any may be null = 0).
__global__ void __launch_bounds__(256) jet_act_kernel(const float* __restrict__ z, const float* __restrict__ zd1,
                                                      const float* __restrict__ zd2, const float* __restrict__ zdd1,
                                                      const float* __restrict__ zdd2, int64_t numel, int act,
                                                      float* __restrict__ y, float* __restrict__ yd,
                                                      float* __restrict__ ydd, int split) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < numel; e += (int64_t)gridDim.x * blockDim.x) {
    const float x = z[e];
    const float d1 = (zd1 ? zd1[e] : 0.f) + (zd2 ? zd2[e] : 0.f);
    const float d2 = (zdd1 ? zdd1[e] : 0.f) + 2.0f * (zdd2 ? zdd2[e] : 0.f);
    float f0 = x, f1 = 1.f, f2 = 0.f;
    if (act == 1) {  // exact-erf GELU: f = x Phi, f' = Phi + x phi, f'' = phi (2 - x^2)
      const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
      const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
      f0 = x * cdf;
      f1 = cdf + x * pdf;
      f2 = pdf * (2.0f - x * x);
    }
    const float o0 = f0, o1 = f1 * d1, o2 = f2 * d1 * d1 + f1 * d2;
    if (!split) {
      y[e] = o0;
      yd[e] = o1;
      ydd[e] = o2;
    } else {  // split-bf16 planes (hi at [0, numel), lo at [numel, 2 numel)) for the tensor-core convolutions
      const float vals[3] = {o0, o1, o2};
      float* outs[3] = {y, yd, ydd};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        uint16_t* hp = reinterpret_cast<uint16_t*>(outs[k]);
        uint32_t u = __float_as_uint(vals[k]);
        u += 0x7FFFu + ((u >> 16) & 1u);
        const uint16_t hb = (uint16_t)(u >> 16);
        const float rem = vals[k] - __uint_as_float((uint32_t)hb << 16);
        uint32_t u2 = __float_as_uint(rem);
        u2 += 0x7FFFu + ((u2 >> 16) & 1u);
        hp[e] = hb;
        hp[numel + e] = (uint16_t)(u2 >> 16);
      }
    }
  }
}

// Head: o = OutImg(z) per channel; L = mean over n*3*h*w of (o - t)^2 (nn.MSELoss, bit_assign.py:192);
// L'' = mean(2 o'^2 + 2 (o - t) o'').  z tensors are NHWC with 4 channels (3 used); target is NCHW.
__global__ void __launch_bounds__(256) jet_head_kernel(const float* __restrict__ z, const float* __restrict__ zd1,
                                                       const float* __restrict__ zd2, const float* __restrict__ zdd1,
                                                       const float* __restrict__ zdd2, const float* __restrict__ tgt,
                                                       int n, int h, int w, int out_bias, double* __restrict__ out) {
  __shared__ float red[32];
  const int64_t pixels = (int64_t)n * h * w;
  const int64_t plane = (int64_t)h * w;
  float acc = 0.f;
  for (int64_t px = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; px < pixels; px += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = px / plane, r = px - b * plane;
    const float4 zz = *reinterpret_cast<const float4*>(z + px * 4);
    const float4 a1 = zd1 ? *reinterpret_cast<const float4*>(zd1 + px * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 a2 = zd2 ? *reinterpret_cast<const float4*>(zd2 + px * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 b1 = zdd1 ? *reinterpret_cast<const float4*>(zdd1 + px * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 b2 = zdd2 ? *reinterpret_cast<const float4*>(zdd2 + px * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float zs[3] = {zz.x, zz.y, zz.z};
    const float d1s[3] = {a1.x + a2.x, a1.y + a2.y, a1.z + a2.z};
    const float d2s[3] = {b1.x + 2.f * b2.x, b1.y + 2.f * b2.y, b1.z + 2.f * b2.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float o, o1, o2;
      if (out_bias == 0) {  // 0.5 tanh + 0.5
        const float t = tanhf(zs[c]);
        const float s = 1.0f - t * t;
        o = 0.5f * t + 0.5f;
        o1 = 0.5f * s * d1s[c];
        o2 = 0.5f * (s * d2s[c] - 2.0f * t * s * d1s[c] * d1s[c]);
      } else {  // sigmoid
        const float sg = sigmoid_f(zs[c]);
        const float s1 = sg * (1.0f - sg);
        o = sg;
        o1 = s1 * d1s[c];
        o2 = s1 * d2s[c] + s1 * (1.0f - 2.0f * sg) * d1s[c] * d1s[c];
      }
      const float tg = tgt[(b * 3 + c) * plane + r];
      acc += 2.0f * o1 * o1 + 2.0f * (o - tg) * o2;
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, (double)acc / (double)(pixels * 3));
}

}  // namespace nq

using namespace nq;

extern "C" int nq_jet_act(const float* z, const float* zd1, const float* zd2, const float* zdd1, const float* zdd2,
                          int64_t numel, int act, void* y, void* yd, void* ydd, int split_out, void* stream) {
  if (!z || !y || !yd || !ydd || numel <= 0 || (act != 0 && act != 1)) return NQ_ERR_BAD_ARG;
  int64_t blocks = (numel + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  jet_act_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(z, zd1, zd2, zdd1, zdd2, numel, act,
                                                                  reinterpret_cast<float*>(y), reinterpret_cast<float*>(yd),
                                                                  reinterpret_cast<float*>(ydd), split_out);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

extern "C" int nq_jet_head(const float* z, const float* zd1, const float* zd2, const float* zdd1, const float* zdd2,
                           const float* target, int n, int h, int w, int out_bias, double* omega_acc, void* stream) {
  if (!z || !target || !omega_acc || n <= 0 || h <= 0 || w <= 0 || (out_bias != 0 && out_bias != 1)) return NQ_ERR_BAD_ARG;
  const int64_t pixels = (int64_t)n * h * w;
  int64_t blocks = (pixels + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  jet_head_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(z, zd1, zd2, zdd1, zdd2, target, n, h, w, out_bias, omega_acc);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}
