// Exact-fp32 (FFMA) implicit-GEMM convolution kernels, NHWC: forward with fused bias + up-shuffle +
// GELU, data gradient with fused GELU' + un-shuffle, weight/bias gradient with deterministic split-K.
// These are the "NQ_PREC_FP32" path: bit-for-bit fp32 products and fp32 accumulation, used for the
// small / HBM-bound stages (stem, head) and as the exact mode of the heavy stages.
// Reference ops replaced: F.conv2d (quant_layer.py:80) + nn.PixelShuffle + nn.GELU (quant_block.py:31-35)
// and their autograd twins.
#include <stdlib.h>
#include "nq_common.cuh"

namespace nq {

int check_conv_desc(const nq_conv_desc* d);

constexpr int BM = 128;  // GEMM-M tile
constexpr int BK = 8;    // GEMM-K step
constexpr int AS_STRIDE = BM + 4;
constexpr int NTHREADS = 256;

enum { EPI_FWD = 0, EPI_DGRAD = 1 };

struct IgemmParams {
  const float* in;     // (n, h, w, C) NHWC
  const float* wmat;   // [K][N]
  const float* bias;   // [N] (fwd) or null
  const float* zprev;  // dgrad: pre-activation of the previous stage, (n, h, w, N), or null
  float* out_z;        // fwd: pre-activation (may be null)
  float* out_y;        // fwd: activated output; dgrad: dz_prev
  int n, h, w, C, ks, pad, N, K, M;
  int rh, rw, cg, act;
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int BN, int EPI>
__global__ void __launch_bounds__(NTHREADS, 2) conv_igemm_kernel(const IgemmParams p) {
  constexpr int TN = BN / 16;        // 8 or 4 output columns per thread
  constexpr int NB4 = BN / 4;        // float4 per B row
  __shared__ __align__(16) float As[2][BK][AS_STRIDE];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;

  // ---- A loader state: one float4 (4 channels of one tap) of one pixel row per thread per k-step
  const int a_row = tid >> 1, a_kq = tid & 1;
  const int a_m = m0 + a_row;
  const bool a_valid_m = a_m < p.M;
  int a_n = 0, a_h = 0, a_w = 0;
  if (a_valid_m) {
    a_w = a_m % p.w;
    const int t = a_m / p.w;
    a_h = t % p.h;
    a_n = t / p.h;
  }
  int a_k = a_kq * 4;  // running k of this thread's float4
  int a_c = a_k % p.C, a_tap = a_k / p.C;
  int a_kh = a_tap / p.ks, a_kw = a_tap % p.ks;

  // ---- B loader state
  const int b_row = tid / NB4, b_c4 = tid % NB4;
  const bool b_active = tid < BK * NB4;
  const int b_col = n0 + b_c4 * 4;

  float4 a_reg = make_float4(0.f, 0.f, 0.f, 0.f), b_reg = make_float4(0.f, 0.f, 0.f, 0.f);

  auto load_tiles = [&](int k0) {
    a_reg = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a_valid_m && a_k < p.K) {
      const int ih = a_h + a_kh - p.pad, iw = a_w + a_kw - p.pad;
      if ((unsigned)ih < (unsigned)p.h && (unsigned)iw < (unsigned)p.w)
        a_reg = ldg4(p.in + ((int64_t)(a_n * p.h + ih) * p.w + iw) * p.C + a_c);
    }
    // advance to the next k-step
    a_k += BK;
    a_c += BK;
    while (a_c >= p.C) {
      a_c -= p.C;
      if (++a_kw == p.ks) { a_kw = 0; ++a_kh; }
    }
    b_reg = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b_active) {
      const int kr = k0 + b_row;
      if (kr < p.K && b_col < p.N) b_reg = ldg4(p.wmat + (int64_t)kr * p.N + b_col);
    }
  };
  auto store_tiles = [&](int buf) {
    As[buf][a_kq * 4 + 0][a_row] = a_reg.x;
    As[buf][a_kq * 4 + 1][a_row] = a_reg.y;
    As[buf][a_kq * 4 + 2][a_row] = a_reg.z;
    As[buf][a_kq * 4 + 3][a_row] = a_reg.w;
    if (b_active) *reinterpret_cast<float4*>(&Bs[buf][b_row][b_c4 * 4]) = b_reg;
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int KT = (p.K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  int cur = 0;
  for (int kt = 0; kt < KT; ++kt) {
    if (kt + 1 < KT) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][64 + ty * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
      {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
        if (TN == 8) {
          const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][kk][(BN / 2) + tx * 4]);
          b[TN - 4] = b1.x; b[TN - 3] = b1.y; b[TN - 2] = b1.z; b[TN - 1] = b1.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < KT) store_tiles(cur ^ 1);
    __syncthreads();
    cur ^= 1;
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
    const int pw = m % p.w;
    const int t = m / p.w;
    const int ph = t % p.h, pn = t / p.h;
#pragma unroll
    for (int jg = 0; jg < TN / 4; ++jg) {
      const int col = n0 + (jg == 0 ? tx * 4 : (BN / 2) + tx * 4);
      if (col >= p.N) continue;
      float4 v = make_float4(acc[i][jg * 4 + 0], acc[i][jg * 4 + 1], acc[i][jg * 4 + 2], acc[i][jg * 4 + 3]);
      if (EPI == EPI_FWD) {
        const float4 bb = ldg4(p.bias + col);
        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        const int grp = col / p.cg, c = col - grp * p.cg;
        const int si = grp / p.rw, sj = grp - si * p.rw;
        const int64_t o = (((int64_t)pn * (p.h * p.rh) + (ph * p.rh + si)) * (p.w * p.rw) + (pw * p.rw + sj)) * p.cg + c;
        if (p.out_z) {  // act 2: keep GELU'(z) for the backward pass instead of z
          const float4 zs = p.act == 2 ? make_float4(gelu_grad_f(v.x), gelu_grad_f(v.y), gelu_grad_f(v.z), gelu_grad_f(v.w)) : v;
          *reinterpret_cast<float4*>(p.out_z + o) = zs;
        }
        if (p.act != 0) { v.x = gelu_f(v.x); v.y = gelu_f(v.y); v.z = gelu_f(v.z); v.w = gelu_f(v.w); }
        *reinterpret_cast<float4*>(p.out_y + o) = v;
      } else {
        if (p.zprev) {
          const float4 z = ldg4(p.zprev + (int64_t)m * p.N + col);
          if (p.act == 1) {
            v.x *= gelu_grad_f(z.x); v.y *= gelu_grad_f(z.y); v.z *= gelu_grad_f(z.z); v.w *= gelu_grad_f(z.w);
          } else if (p.act == 2) {  // z_prev already holds the derivative
            v.x *= z.x; v.y *= z.y; v.z *= z.z; v.w *= z.w;
          }
        }
        const int qh = ph / p.rh, si = ph - qh * p.rh;
        const int qw = pw / p.rw, sj = pw - qw * p.rw;
        const int64_t o = (((int64_t)pn * (p.h / p.rh) + qh) * (p.w / p.rw) + qw) * ((int64_t)p.rh * p.rw * p.N) +
                          (int64_t)(si * p.rw + sj) * p.N + col;
        *reinterpret_cast<float4*>(p.out_y + o) = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient: D[kf][n'] = sum_pixels X[pixel + off(tap(kf))][ci(kf)] * dZ[pixel][n'],
// row kf == K carries the bias gradient (X == 1).  blockIdx.z = pixel split.
// ---------------------------------------------------------------------------------------------
struct WgradParams {
  const float* x;   // (n, h, w, C)
  const float* dz;  // (n, h, w, N)
  float* out;       // [splits][(K+4)][N]
  int n, h, w, C, ks, pad, N, K, P;  // P = n*h*w pixels
  int chunk;                         // pixels per split (multiple of BK)
};

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 2) conv_wgrad_kernel(const WgradParams p) {
  constexpr int TN = BN / 16;
  constexpr int NB4 = BN / 4;
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;
  const int rows = p.K + 4;

  const int p_begin = blockIdx.z * p.chunk;
  const int p_end = min(p.P, p_begin + p.chunk);

  // A loader: pixel a_pk of the step, rows kf = m0 + a_m4*4 .. +3 (4 channels of one tap)
  const int a_pk = tid >> 5, a_m4 = tid & 31;
  const int a_kf = m0 + a_m4 * 4;
  const bool a_is_w = a_kf < p.K;
  const bool a_is_bias = a_kf == p.K;
  int a_dh = 0, a_dw = 0, a_c = 0;
  if (a_is_w) {
    const int tap = a_kf / p.C;
    a_c = a_kf - tap * p.C;
    a_dh = tap / p.ks - p.pad;
    a_dw = tap % p.ks - p.pad;
  }
  const int b_pk = tid / NB4, b_c4 = tid % NB4;
  const bool b_active = tid < BK * NB4;
  const int b_col = n0 + b_c4 * 4;

  float4 a_reg, b_reg;
  auto load_tiles = [&](int pix0) {
    a_reg = make_float4(0.f, 0.f, 0.f, 0.f);
    const int pa = pix0 + a_pk;
    if (pa < p_end) {
      if (a_is_w) {
        const int pw = pa % p.w;
        const int t = pa / p.w;
        const int ph = t % p.h, pn = t / p.h;
        const int ih = ph + a_dh, iw = pw + a_dw;
        if ((unsigned)ih < (unsigned)p.h && (unsigned)iw < (unsigned)p.w)
          a_reg = ldg4(p.x + ((int64_t)(pn * p.h + ih) * p.w + iw) * p.C + a_c);
      } else if (a_is_bias) {
        a_reg.x = 1.0f;
      }
    }
    b_reg = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b_active) {
      const int pb = pix0 + b_pk;
      if (pb < p_end && b_col < p.N) b_reg = ldg4(p.dz + (int64_t)pb * p.N + b_col);
    }
  };
  auto store_tiles = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][a_pk][a_m4 * 4]) = a_reg;
    if (b_active) *reinterpret_cast<float4*>(&Bs[buf][b_pk][b_c4 * 4]) = b_reg;
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int KT = (p_end - p_begin + BK - 1) / BK;
  if (KT > 0) {
    load_tiles(p_begin);
    store_tiles(0);
  }
  __syncthreads();
  int cur = 0;
  for (int kt = 0; kt < KT; ++kt) {
    if (kt + 1 < KT) load_tiles(p_begin + (kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][64 + ty * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
      {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
        if (TN == 8) {
          const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][kk][(BN / 2) + tx * 4]);
          b[TN - 4] = b1.x; b[TN - 3] = b1.y; b[TN - 2] = b1.z; b[TN - 1] = b1.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < KT) store_tiles(cur ^ 1);
    __syncthreads();
    cur ^= 1;
  }

  float* out = p.out + (int64_t)blockIdx.z * rows * p.N;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= rows) continue;
#pragma unroll
    for (int jg = 0; jg < TN / 4; ++jg) {
      const int col = n0 + (jg == 0 ? tx * 4 : (BN / 2) + tx * 4);
      if (col >= p.N) continue;
      *reinterpret_cast<float4*>(out + (int64_t)m * p.N + col) =
          make_float4(acc[i][jg * 4 + 0], acc[i][jg * 4 + 1], acc[i][jg * 4 + 2], acc[i][jg * 4 + 3]);
    }
  }
}

// fixed-order sum over splits (deterministic)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int64_t numel4, int splits,
                                                            float* __restrict__ out) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < numel4;
       e += (int64_t)gridDim.x * blockDim.x) {
    float4 s = reinterpret_cast<const float4*>(ws)[e];
    for (int k = 1; k < splits; ++k) {
      const float4 v = reinterpret_cast<const float4*>(ws)[(int64_t)k * numel4 + e];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(out)[e] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// Head: 3x3 conv to 3 (padded 4) channels + OutImg + loss.  HBM-bound (about 23 FLOP/B): one pixel
// per thread, the input tile with its halo is staged through shared memory in 8-channel chunks
// (pixel stride 12 floats -> conflict-free float4 reads), weights of the chunk in shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int HT_H = 8, HT_W = 32, H_CH = 8, H_PS = 12;  // tile, channel chunk, smem pixel stride

struct HeadParams {
  const float* x; const float* w; const float* bias; const float* target;
  float* img; float* loss_sum; float* dz;
  int n, h, w_, C, out_bias;
  float p, inv_mean;
  // split-bf16 variants (tensor-core engine): input planes (hi, lo) of (n, h, w, C) bf16; gradient planes of
  // (n, h, w, 8) bf16 (3 real channels + zeros: one 16-byte chunk per pixel for the cp.async consumers)
  const uint16_t* x_hi; const uint16_t* x_lo;
  uint16_t* dz_hi; uint16_t* dz_lo;
};

__device__ __forceinline__ float bf16_bits_to_f(uint32_t b) { return __uint_as_float(b << 16); }
__device__ __forceinline__ uint16_t f_to_bf16_bits(float v) {  // round to nearest even
  uint32_t u = __float_as_uint(v);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// ---------------------------------------------------------------------------------------------
// Head forward, 4 pixels per thread.  (A one-pixel-per-thread kernel issued 16 shared-memory wavefronts -- 2 input
// float4 + 8 broadcast weight float4 -- per 24 FMAs and was bound by the shared-memory pipe: 92 % busy, 0.27 ms; removed.)  Here a thread
// owns a 4-row strip of one column: the 8 weight loads of a (tap, channel quad) feed 4 pixels, and the 6 input rows
// of a strip are loaded once per kernel column -- 72 wavefronts per 288 FMAs, balanced with the FMA pipe.
// Same staging (pixel-major, 8-channel chunks), same per-pixel epilogue.
// ---------------------------------------------------------------------------------------------
constexpr int HV_TW = 32, HV_TH = 32, HV_THREADS = 256;  // thread = (column lx, strip ly of 4 rows)
constexpr int HV_SMEM = ((HV_TH + 2) * (HV_TW + 2) * H_PS + 9 * H_CH * 4) * 4;

__global__ void __launch_bounds__(HV_THREADS) head_fwd_loss_strip_kernel(const HeadParams q) {
  extern __shared__ __align__(16) float hsm[];
  float* xs = hsm;                                         // [(HV_TH + 2)][(HV_TW + 2)][H_PS]
  float* ws = hsm + (HV_TH + 2) * (HV_TW + 2) * H_PS;      // [9][H_CH][4]
  __shared__ float red[32];
  const int tid = threadIdx.x;
  const int lx = tid % HV_TW, ly = tid / HV_TW;
  const int tiles_w = (q.w_ + HV_TW - 1) / HV_TW, tiles_h = (q.h + HV_TH - 1) / HV_TH;
  int b = blockIdx.x;
  const int tw = b % tiles_w; b /= tiles_w;
  const int th = b % tiles_h;
  const int bn = b / tiles_h;
  const int x0 = tw * HV_TW, y0 = th * HV_TH;

  float acc[4][3];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = 0.f;

  for (int c0 = 0; c0 < q.C; c0 += H_CH) {
    const int cw = min(H_CH, q.C - c0);  // 8 or 4
    const int c4n = cw >> 2;
    __syncthreads();
#pragma unroll 3
    for (int i = tid; i < (HV_TH + 2) * (HV_TW + 2) * 2; i += HV_THREADS) {
      const int c4 = i & 1, pix = i >> 1;
      const int sy = pix / (HV_TW + 2), sx = pix - sy * (HV_TW + 2);
      const int gx = x0 + sx - 1, gy = y0 + sy - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c4 < c4n && (unsigned)gx < (unsigned)q.w_ && (unsigned)gy < (unsigned)q.h) {
        const int64_t o = (int64_t)((bn * q.h + gy) * q.w_ + gx) * q.C + c0 + c4 * 4;
        if (q.x_hi) {
          const uint2 hb = __ldg(reinterpret_cast<const uint2*>(q.x_hi + o)), lb = __ldg(reinterpret_cast<const uint2*>(q.x_lo + o));
          v.x = bf16_bits_to_f(hb.x & 0xFFFFu) + bf16_bits_to_f(lb.x & 0xFFFFu);
          v.y = bf16_bits_to_f(hb.x >> 16) + bf16_bits_to_f(lb.x >> 16);
          v.z = bf16_bits_to_f(hb.y & 0xFFFFu) + bf16_bits_to_f(lb.y & 0xFFFFu);
          v.w = bf16_bits_to_f(hb.y >> 16) + bf16_bits_to_f(lb.y >> 16);
        } else {
          v = ldg4(q.x + o);
        }
      }
      *reinterpret_cast<float4*>(&xs[pix * H_PS + c4 * 4]) = v;
    }
    for (int i = tid; i < 9 * H_CH; i += HV_THREADS) {
      const int tap = i / H_CH, c = i % H_CH;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < cw) v = ldg4(q.w + ((int64_t)tap * q.C + c0 + c) * 4);
      *reinterpret_cast<float4*>(&ws[i * 4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const float* xcol = &xs[((4 * ly) * (HV_TW + 2) + (lx + kw)) * H_PS];
#pragma unroll
      for (int c4 = 0; c4 < 2; ++c4) {
        float4 xv[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) xv[r] = *reinterpret_cast<const float4*>(xcol + r * (HV_TW + 2) * H_PS + c4 * 4);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const float* wp = &ws[((kh * 3 + kw) * H_CH + c4 * 4) * 4];
          const float4 w0 = *reinterpret_cast<const float4*>(wp), w1 = *reinterpret_cast<const float4*>(wp + 4),
                       w2 = *reinterpret_cast<const float4*>(wp + 8), w3 = *reinterpret_cast<const float4*>(wp + 12);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 x = xv[i + kh];
            acc[i][0] = fmaf(x.x, w0.x, acc[i][0]); acc[i][1] = fmaf(x.x, w0.y, acc[i][1]); acc[i][2] = fmaf(x.x, w0.z, acc[i][2]);
            acc[i][0] = fmaf(x.y, w1.x, acc[i][0]); acc[i][1] = fmaf(x.y, w1.y, acc[i][1]); acc[i][2] = fmaf(x.y, w1.z, acc[i][2]);
            acc[i][0] = fmaf(x.z, w2.x, acc[i][0]); acc[i][1] = fmaf(x.z, w2.y, acc[i][1]); acc[i][2] = fmaf(x.z, w2.z, acc[i][2]);
            acc[i][0] = fmaf(x.w, w3.x, acc[i][0]); acc[i][1] = fmaf(x.w, w3.y, acc[i][1]); acc[i][2] = fmaf(x.w, w3.z, acc[i][2]);
          }
        }
      }
    }
  }

  float loss = 0.f;
  const int px = x0 + lx;
  const int64_t plane = (int64_t)q.h * q.w_;
  const float b0 = q.bias[0], b1 = q.bias[1], b2 = q.bias[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int py = y0 + 4 * ly + i;
    if (px >= q.w_ || py >= q.h) continue;
    const float v[3] = {acc[i][0] + b0, acc[i][1] + b1, acc[i][2] + b2};
    float g[3] = {0.f, 0.f, 0.f};
    const int64_t o = (int64_t)bn * 3 * plane + (int64_t)py * q.w_ + px;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float outv, dout;
      if (q.out_bias == 0) {
        const float t = tanhf(v[c]);
        outv = t * 0.5f + 0.5f;
        dout = 0.5f * (1.0f - t * t);
      } else {
        outv = sigmoid_f(v[c]);
        dout = outv * (1.0f - outv);
      }
      if (q.img) q.img[o + c * plane] = outv;
      if (q.target) {
        const float dlt = outv - __ldg(q.target + o + c * plane);
        const float a = fabsf(dlt);
        if (q.p == 2.0f) {
          loss += dlt * dlt;
          g[c] = 2.0f * dlt * q.inv_mean * dout;
        } else {
          loss += powf(a, q.p);
          const float sgn = dlt > 0.f ? 1.f : (dlt < 0.f ? -1.f : 0.f);
          g[c] = q.p * powf(a, q.p - 1.0f) * sgn * q.inv_mean * dout;
        }
      }
    }
    const int64_t pix = (int64_t)(bn * q.h + py) * q.w_ + px;
    if (q.dz) *reinterpret_cast<float4*>(q.dz + pix * 4) = make_float4(g[0], g[1], g[2], 0.f);
    if (q.dz_hi) {
      uint16_t hb[3], lb[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        hb[c] = f_to_bf16_bits(g[c]);
        lb[c] = f_to_bf16_bits(g[c] - bf16_bits_to_f(hb[c]));
      }
      *reinterpret_cast<uint4*>(q.dz_hi + pix * 8) = make_uint4((uint32_t)hb[0] | ((uint32_t)hb[1] << 16), hb[2], 0u, 0u);
      *reinterpret_cast<uint4*>(q.dz_lo + pix * 8) = make_uint4((uint32_t)lb[0] | ((uint32_t)lb[1] << 16), lb[2], 0u, 0u);
    }
  }
  if (q.target && q.loss_sum) {
    loss = block_sum(loss, red);
    if (tid == 0) atomicAdd(q.loss_sum, loss);
  }
}

// ---------------------------------------------------------------------------------------------
// Head forward for the split-bf16 engine, asynchronous.  Both kernels above stage every 8-channel chunk
// synchronously (global -> registers -> convert -> shared, two block barriers per chunk) and sit at ~0.26 ms whatever
// their inner loop does: neither DRAM (19 %) nor the FMA pipe is busy, the loads are simply exposed.  Here persistent
// CTAs stream the RAW bf16 planes with cp.async through a 4-deep ring of (tile, chunk) items, the hi + lo sum is
// formed in registers when a pixel is read, and all weights stay resident: no staging pass, one barrier per item.
// Thread mapping and inner loop as head_fwd_loss_strip_kernel (4-row strip of one column).
// ---------------------------------------------------------------------------------------------
constexpr int HA_TW = 32, HA_TH = 32, HA_THREADS = 256, HA_STAGES = 4;
constexpr int HA_NPIX = (HA_TH + 2) * (HA_TW + 2);
constexpr int HA_STAGE_BYTES = HA_NPIX * 16 * 2;  // hi plane + lo plane, 8 channels (16 bytes) per pixel each
constexpr int HA_MAXC = 64;

__device__ __forceinline__ void ha_cp16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(HA_THREADS, 1) head_fwd_loss_async_kernel(const HeadParams q) {
  extern __shared__ __align__(16) uint8_t hraw[];
  float* ws = reinterpret_cast<float*>(hraw + HA_STAGES * HA_STAGE_BYTES);  // [9][C][4] all channels, loaded once
  __shared__ float red[32];
  const int tid = threadIdx.x;
  const int lx = tid % HA_TW, ly = tid / HA_TW;
  const int tiles_w = (q.w_ + HA_TW - 1) / HA_TW, tiles_h = (q.h + HA_TH - 1) / HA_TH;
  const int tiles = tiles_w * tiles_h * q.n;
  const int nch = q.C / 8;  // 8-channel chunks per tile (C % 8 == 0 on this path)
  const int my_tiles = ((int)blockIdx.x < tiles) ? (tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int items = my_tiles * nch;
  const uint32_t raw0 = (uint32_t)__cvta_generic_to_shared(hraw);

  for (int i = tid; i < 9 * q.C; i += HA_THREADS) *reinterpret_cast<float4*>(&ws[i * 4]) = ldg4(q.w + (int64_t)i * 4);

  auto issue = [&](int item) {  // one (tile, chunk) item -> ring slot item % HA_STAGES
    if (item < items) {
      const int t = blockIdx.x + (item / nch) * gridDim.x, c0 = (item % nch) * 8;
      const int tw = t % tiles_w, th = (t / tiles_w) % tiles_h, bn = t / (tiles_w * tiles_h);
      const int x0 = tw * HA_TW, y0 = th * HA_TH;
      const uint32_t dst = raw0 + (item % HA_STAGES) * HA_STAGE_BYTES;
      for (int pix = tid; pix < HA_NPIX; pix += HA_THREADS) {
        const int sy = pix / (HA_TW + 2), sx = pix - sy * (HA_TW + 2);
        const int gx = x0 + sx - 1, gy = y0 + sy - 1;
        const bool ok = (unsigned)gx < (unsigned)q.w_ && (unsigned)gy < (unsigned)q.h;
        const int64_t o = ok ? (int64_t)((bn * q.h + gy) * q.w_ + gx) * q.C + c0 : 0;
        ha_cp16(dst + pix * 16, q.x_hi + o, ok ? 16u : 0u);
        ha_cp16(dst + HA_NPIX * 16 + pix * 16, q.x_lo + o, ok ? 16u : 0u);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // (empty groups keep the wait counts uniform)
  };

#pragma unroll
  for (int k = 0; k < HA_STAGES - 1; ++k) issue(k);

  float acc[4][3];
  float loss = 0.f;
  const int64_t plane = (int64_t)q.h * q.w_;
  for (int item = 0; item < items; ++item) {
    const int ch = item % nch;
    if (ch == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = 0.f;
    }
    asm volatile("cp.async.wait_group %0;" ::"n"(HA_STAGES - 2) : "memory");
    __syncthreads();               // item's data visible to all; everyone is done with the slot refilled next
    issue(item + HA_STAGES - 1);
    const uint8_t* hi = hraw + (item % HA_STAGES) * HA_STAGE_BYTES;
    const uint8_t* lo = hi + HA_NPIX * 16;
    const float* wc = ws + ch * 8 * 4;  // weights of this chunk: ws[(tap * C + ch * 8 + c) * 4]
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int p0 = (4 * ly) * (HA_TW + 2) + lx + kw;
      float xv[6][8];
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        const uint4 h = *reinterpret_cast<const uint4*>(hi + (p0 + r * (HA_TW + 2)) * 16);
        const uint4 l = *reinterpret_cast<const uint4*>(lo + (p0 + r * (HA_TW + 2)) * 16);
        xv[r][0] = __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
        xv[r][1] = __uint_as_float(h.x & 0xFFFF0000u) + __uint_as_float(l.x & 0xFFFF0000u);
        xv[r][2] = __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
        xv[r][3] = __uint_as_float(h.y & 0xFFFF0000u) + __uint_as_float(l.y & 0xFFFF0000u);
        xv[r][4] = __uint_as_float(h.z << 16) + __uint_as_float(l.z << 16);
        xv[r][5] = __uint_as_float(h.z & 0xFFFF0000u) + __uint_as_float(l.z & 0xFFFF0000u);
        xv[r][6] = __uint_as_float(h.w << 16) + __uint_as_float(l.w << 16);
        xv[r][7] = __uint_as_float(h.w & 0xFFFF0000u) + __uint_as_float(l.w & 0xFFFF0000u);
      }
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const float* wp = wc + (kh * 3 + kw) * q.C * 4;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 wv = *reinterpret_cast<const float4*>(wp + c * 4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[i][0] = fmaf(xv[i + kh][c], wv.x, acc[i][0]);
            acc[i][1] = fmaf(xv[i + kh][c], wv.y, acc[i][1]);
            acc[i][2] = fmaf(xv[i + kh][c], wv.z, acc[i][2]);
          }
        }
      }
    }
    if (ch != nch - 1) continue;
    // ---- per-pixel epilogue of the finished tile (OutImg + loss + dL/dz), as in the kernels above
    const int t = blockIdx.x + (item / nch) * gridDim.x;
    const int tw = t % tiles_w, th = (t / tiles_w) % tiles_h, bn = t / (tiles_w * tiles_h);
    const int px = tw * HA_TW + lx;
    const float b0 = q.bias[0], b1 = q.bias[1], b2 = q.bias[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int py = th * HA_TH + 4 * ly + i;
      if (px >= q.w_ || py >= q.h) continue;
      const float v[3] = {acc[i][0] + b0, acc[i][1] + b1, acc[i][2] + b2};
      float g[3] = {0.f, 0.f, 0.f};
      const int64_t o = (int64_t)bn * 3 * plane + (int64_t)py * q.w_ + px;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float outv, dout;
        if (q.out_bias == 0) {
          const float tt = tanhf(v[c]);
          outv = tt * 0.5f + 0.5f;
          dout = 0.5f * (1.0f - tt * tt);
        } else {
          outv = sigmoid_f(v[c]);
          dout = outv * (1.0f - outv);
        }
        if (q.img) q.img[o + c * plane] = outv;
        if (q.target) {
          const float dlt = outv - __ldg(q.target + o + c * plane);
          const float a = fabsf(dlt);
          if (q.p == 2.0f) {
            loss += dlt * dlt;
            g[c] = 2.0f * dlt * q.inv_mean * dout;
          } else {
            loss += powf(a, q.p);
            const float sgn = dlt > 0.f ? 1.f : (dlt < 0.f ? -1.f : 0.f);
            g[c] = q.p * powf(a, q.p - 1.0f) * sgn * q.inv_mean * dout;
          }
        }
      }
      if (q.dz_hi) {
        const int64_t pix = (int64_t)(bn * q.h + py) * q.w_ + px;
        uint16_t hb[3], lb[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          hb[c] = f_to_bf16_bits(g[c]);
          lb[c] = f_to_bf16_bits(g[c] - bf16_bits_to_f(hb[c]));
        }
        *reinterpret_cast<uint4*>(q.dz_hi + pix * 8) = make_uint4((uint32_t)hb[0] | ((uint32_t)hb[1] << 16), hb[2], 0u, 0u);
        *reinterpret_cast<uint4*>(q.dz_lo + pix * 8) = make_uint4((uint32_t)lb[0] | ((uint32_t)lb[1] << 16), lb[2], 0u, 0u);
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (q.target && q.loss_sum) {
    loss = block_sum(loss, red);
    if (tid == 0) atomicAdd(q.loss_sum, loss);
  }
}

static inline int64_t hcdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
// (A register-blocked variant -- 2 x 4 pixels per thread over a channel-planar shared-memory tile -- was measured at
// 0.31 ms against 0.27 ms for this kernel at 1280x640x2: with 44 KB of staging per CTA its synchronous tile loads are
// exposed, and the head's input arrives as split-bf16 NHWC, which needs a converting transpose.  Not kept.)
static int launch_head_fwd(const HeadParams& q, cudaStream_t s) {
  if (q.x_hi && !q.target && q.C % 8 == 0 && q.C <= HA_MAXC) {
    // decode (no target): the asynchronous kernel; with the loss epilogue its 8 warps per SM cannot hide the target
    // loads and gradient stores, and the synchronous strip kernel below is faster (0.26 vs 0.29 ms at 1280x640x2)
    const int smem = HA_STAGES * HA_STAGE_BYTES + 9 * q.C * 16;
    NQ_CUDA_CHECK(cudaFuncSetAttribute(head_fwd_loss_async_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t tiles = hcdiv(q.w_, HA_TW) * hcdiv(q.h, HA_TH) * q.n;
    const int64_t grid = tiles < sm_count() ? tiles : sm_count();
    head_fwd_loss_async_kernel<<<(unsigned)grid, HA_THREADS, smem, s>>>(q);
  } else {
    NQ_CUDA_CHECK(cudaFuncSetAttribute(head_fwd_loss_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HV_SMEM));
    const int64_t blocks = hcdiv(q.w_, HV_TW) * hcdiv(q.h, HV_TH) * q.n;
    head_fwd_loss_strip_kernel<<<(unsigned)blocks, HV_THREADS, HV_SMEM, s>>>(q);
  }
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

// Head weight gradient.  Persistent CTAs loop over pixel tiles; inside a tile the 9*8 (tap, channel)
// pairs of a chunk (+1 bias slot) are spread over three thread groups that split the tile rows;
// per-CTA partials are then summed in a fixed order.
constexpr int HW_MAXCHUNK = 16;  // supports C <= 128 (head inputs: 24..40 channels in the 3M configs, 112 in the 12M one)
constexpr int HW_SLOTS = 9 * H_CH + 1;

struct HeadWgradParams {
  const float* x; const float* dz; float* out;
  int n, h, w_, C, tiles;
};

__global__ void __launch_bounds__(HT_H * HT_W) head_wgrad_kernel(const HeadWgradParams q) {
  __shared__ __align__(16) float xs[(HT_H + 2) * (HT_W + 2) * H_PS];
  __shared__ __align__(16) float gs[HT_H * HT_W * 4];
  __shared__ float comb[3][HW_SLOTS][4];
  const int tid = threadIdx.x;
  const int grp = tid / HW_SLOTS, slot = tid % HW_SLOTS;  // grp 0..2 active, (tid >= 3*HW_SLOTS idle)
  const bool active = grp < 3;
  const bool is_bias = slot == HW_SLOTS - 1;
  const int tap = is_bias ? 0 : slot / H_CH, cc = is_bias ? 0 : slot % H_CH;
  const int kh = tap / 3, kw = tap % 3;
  const int tiles_w = (q.w_ + HT_W - 1) / HT_W, tiles_h = (q.h + HT_H - 1) / HT_H;
  const int n_chunks = (q.C + H_CH - 1) / H_CH;

  float acc[HW_MAXCHUNK][3];
#pragma unroll
  for (int i = 0; i < HW_MAXCHUNK; ++i) acc[i][0] = acc[i][1] = acc[i][2] = 0.f;

  for (int tile = blockIdx.x; tile < q.tiles; tile += gridDim.x) {
    int b = tile;
    const int tw = b % tiles_w; b /= tiles_w;
    const int th = b % tiles_h;
    const int bn = b / tiles_h;
    const int x0 = tw * HT_W, y0 = th * HT_H;
    __syncthreads();
    {  // gradient tile (zero outside the image)
      const int lx = tid % HT_W, ly = tid / HT_W;
      const int gx = x0 + lx, gy = y0 + ly;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gx < q.w_ && gy < q.h) v = ldg4(q.dz + ((int64_t)(bn * q.h + gy) * q.w_ + gx) * 4);
      *reinterpret_cast<float4*>(&gs[tid * 4]) = v;
    }
#pragma unroll
    for (int ch = 0; ch < HW_MAXCHUNK; ++ch) {
      if (ch >= n_chunks) break;
      const int c0 = ch * H_CH;
      const int c4n = min(H_CH, q.C - c0) >> 2;
      __syncthreads();
      for (int i = tid; i < (HT_H + 2) * (HT_W + 2) * 2; i += HT_H * HT_W) {
        const int c4 = i & 1, pix = i >> 1;
        const int sx = pix % (HT_W + 2), sy = pix / (HT_W + 2);
        const int gx = x0 + sx - 1, gy = y0 + sy - 1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4 < c4n && (unsigned)gx < (unsigned)q.w_ && (unsigned)gy < (unsigned)q.h)
          v = ldg4(q.x + ((int64_t)(bn * q.h + gy) * q.w_ + gx) * q.C + c0 + c4 * 4);
        *reinterpret_cast<float4*>(&xs[pix * H_PS + c4 * 4]) = v;
      }
      __syncthreads();
      if (active && (!is_bias || ch == 0)) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        // group g takes tile rows g, g+3, g+6
        for (int ly = grp; ly < HT_H; ly += 3) {
#pragma unroll 8
          for (int lx = 0; lx < HT_W; ++lx) {
            const float4 gv = *reinterpret_cast<const float4*>(&gs[(ly * HT_W + lx) * 4]);
            const float xv = is_bias ? 1.0f : xs[((ly + kh) * (HT_W + 2) + (lx + kw)) * H_PS + cc];
            s0 = fmaf(xv, gv.x, s0);
            s1 = fmaf(xv, gv.y, s1);
            s2 = fmaf(xv, gv.z, s2);
          }
        }
        acc[ch][0] += s0; acc[ch][1] += s1; acc[ch][2] += s2;
      }
    }
  }
  // combine the three groups and write this CTA's partial [(9*C + 4)][4]
  const int rows = 9 * q.C + 4;
  float* out = q.out + (int64_t)blockIdx.x * rows * 4;
  for (int i = tid; i < rows * 4; i += HT_H * HT_W) out[i] = 0.f;
#pragma unroll
  for (int ch = 0; ch < HW_MAXCHUNK; ++ch) {
    if (ch >= n_chunks) break;
    __syncthreads();
    if (active) { comb[grp][slot][0] = acc[ch][0]; comb[grp][slot][1] = acc[ch][1]; comb[grp][slot][2] = acc[ch][2]; }
    __syncthreads();
    if (tid < HW_SLOTS) {
      const int c = ch * H_CH + cc;
      const bool bias_slot = tid == HW_SLOTS - 1;
      if ((bias_slot && ch == 0) || (!bias_slot && c < q.C)) {
        const int row = bias_slot ? 9 * q.C : tap * q.C + c;
#pragma unroll
        for (int o = 0; o < 3; ++o) out[row * 4 + o] = (comb[0][tid][o] + comb[1][tid][o]) + comb[2][tid][o];
      }
    }
  }
}

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

template <int EPI>
static int launch_igemm(const IgemmParams& p, cudaStream_t s) {
  const int rem = p.N % 128;
  const bool use64 = (p.N <= 64) || (rem > 0 && rem <= 64);
  dim3 grid((unsigned)cdiv(p.M, BM), (unsigned)cdiv(p.N, use64 ? 64 : 128));
  if (use64) conv_igemm_kernel<64, EPI><<<grid, NTHREADS, 0, s>>>(p);
  else conv_igemm_kernel<128, EPI><<<grid, NTHREADS, 0, s>>>(p);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

}  // namespace nq

using namespace nq;

extern "C" int nq_conv_fwd(const nq_conv_desc* d, const float* x, const float* wk, const float* bias_packed, float* z,
                           float* y, void* stream) {
  int st = check_conv_desc(d);
  if (st) return st;
  if (!x || !wk || !bias_packed || !y) return NQ_ERR_BAD_ARG;
  IgemmParams p{};
  p.in = x; p.wmat = wk; p.bias = bias_packed; p.zprev = nullptr; p.out_z = z; p.out_y = y;
  p.n = d->n; p.h = d->h; p.w = d->w; p.C = d->cin_p; p.ks = d->ksize; p.pad = d->ksize / 2;
  p.N = d->rh * d->rw * d->cg; p.K = d->ksize * d->ksize * d->cin_p;
  const int64_t M = (int64_t)d->n * d->h * d->w;
  if (M > 0x7fffffffLL) return NQ_ERR_BAD_SHAPE;
  p.M = (int)M;
  p.rh = d->rh; p.rw = d->rw; p.cg = d->cg; p.act = d->act;
  return launch_igemm<EPI_FWD>(p, as_stream(stream));
}

extern "C" int nq_conv_dgrad(const nq_conv_desc* d, const float* dz, const float* wt, const float* z_prev, int prev_rh,
                             int prev_rw, int prev_act, float* dz_prev, void* stream) {
  int st = check_conv_desc(d);
  if (st) return st;
  if (!dz || !wt || !dz_prev || prev_rh <= 0 || prev_rw <= 0) return NQ_ERR_BAD_ARG;
  if (d->h % prev_rh || d->w % prev_rw) return NQ_ERR_BAD_SHAPE;
  if (prev_act < 0 || prev_act > 2) return NQ_ERR_BAD_ARG;
  IgemmParams p{};
  p.in = dz; p.wmat = wt; p.bias = nullptr; p.zprev = z_prev; p.out_z = nullptr; p.out_y = dz_prev;
  p.n = d->n; p.h = d->h; p.w = d->w; p.C = d->rh * d->rw * d->cg; p.ks = d->ksize; p.pad = d->ksize / 2;
  p.N = d->cin_p; p.K = d->ksize * d->ksize * p.C;
  const int64_t M = (int64_t)d->n * d->h * d->w;
  if (M > 0x7fffffffLL) return NQ_ERR_BAD_SHAPE;
  p.M = (int)M;
  p.rh = prev_rh; p.rw = prev_rw; p.cg = d->cin_p; p.act = prev_act;
  return launch_igemm<EPI_DGRAD>(p, as_stream(stream));
}

extern "C" int nq_conv_wgrad(const nq_conv_desc* d, const float* x, const float* dz, float* dwk, float* workspace,
                             int64_t workspace_floats, int splits, void* stream) {
  int st = check_conv_desc(d);
  if (st) return st;
  if (!x || !dz || !dwk || splits < 1) return NQ_ERR_BAD_ARG;
  WgradParams p{};
  p.x = x; p.dz = dz;
  p.n = d->n; p.h = d->h; p.w = d->w; p.C = d->cin_p; p.ks = d->ksize; p.pad = d->ksize / 2;
  p.N = d->rh * d->rw * d->cg; p.K = d->ksize * d->ksize * d->cin_p;
  const int64_t P = (int64_t)d->n * d->h * d->w;
  if (P > 0x7fffffffLL) return NQ_ERR_BAD_SHAPE;
  p.P = (int)P;
  const int rows = p.K + 4;
  const int64_t tile_elems = (int64_t)rows * p.N;
  if (splits > 1 && (!workspace || workspace_floats < tile_elems * splits)) return NQ_ERR_WORKSPACE;
  p.chunk = (int)(cdiv(cdiv(P, splits), BK) * BK);
  p.out = splits > 1 ? workspace : dwk;
  cudaStream_t s = as_stream(stream);
  const int rem = p.N % 128;
  const bool use64 = (p.N <= 64) || (rem > 0 && rem <= 64);
  dim3 grid((unsigned)cdiv(rows, BM), (unsigned)cdiv(p.N, use64 ? 64 : 128), (unsigned)splits);
  if (use64) conv_wgrad_kernel<64><<<grid, NTHREADS, 0, s>>>(p);
  else conv_wgrad_kernel<128><<<grid, NTHREADS, 0, s>>>(p);
  NQ_LAUNCH_CHECK();
  if (splits > 1) {
    const int64_t n4 = tile_elems / 4;
    int64_t blocks = cdiv(n4, 256);
    if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
    splitk_reduce_kernel<<<(unsigned)blocks, 256, 0, s>>>(workspace, n4, splits, dwk);
    NQ_LAUNCH_CHECK();
  }
  return NQ_OK;
}

static int check_head(const nq_conv_desc* d) {
  int st = check_conv_desc(d);
  if (st) return st;
  if (d->ksize != 3 || d->rh != 1 || d->rw != 1 || d->cout != 3 || d->cg != 4) return NQ_ERR_BAD_SHAPE;
  return NQ_OK;
}

extern "C" int nq_head_fwd_loss(const nq_conv_desc* d, const float* x, const float* w_head, const float* bias_head,
                                int out_bias, const float* target, float p, float mean_pixels, float* img,
                                float* loss_sum, float* dz_head, void* stream) {
  int st = check_head(d);
  if (st) return st;
  if (!x || !w_head || !bias_head) return NQ_ERR_BAD_ARG;
  if (out_bias != 0 && out_bias != 1) return NQ_ERR_UNSUPPORTED;
  if (target && (!(p > 0.f) || !(mean_pixels > 0.f))) return NQ_ERR_BAD_ARG;
  if (!target && !img) return NQ_ERR_BAD_ARG;
  HeadParams q{};
  q.x = x; q.w = w_head; q.bias = bias_head; q.target = target; q.img = img; q.loss_sum = loss_sum; q.dz = dz_head;
  q.n = d->n; q.h = d->h; q.w_ = d->w; q.C = d->cin_p; q.out_bias = out_bias; q.p = p;
  q.inv_mean = target ? 1.0f / mean_pixels : 0.f;
  return launch_head_fwd(q, as_stream(stream));
}

extern "C" int nq_head_fwd_loss_split(const nq_conv_desc* d, const void* x_split, const float* w_head, const float* bias_head,
                                      int out_bias, const float* target, float p, float mean_pixels, float* img,
                                      float* loss_sum, void* dz_head_split, void* stream) {
  int st = check_head(d);
  if (st) return st;
  if (!x_split || !w_head || !bias_head) return NQ_ERR_BAD_ARG;
  if (out_bias != 0 && out_bias != 1) return NQ_ERR_UNSUPPORTED;
  if (target && (!(p > 0.f) || !(mean_pixels > 0.f))) return NQ_ERR_BAD_ARG;
  if (!target && !img) return NQ_ERR_BAD_ARG;
  if (d->cin_p % 8) return NQ_ERR_BAD_SHAPE;
  HeadParams q{};
  const int64_t pix = (int64_t)d->n * d->h * d->w;
  q.x = nullptr; q.w = w_head; q.bias = bias_head; q.target = target; q.img = img; q.loss_sum = loss_sum; q.dz = nullptr;
  q.x_hi = reinterpret_cast<const uint16_t*>(x_split);
  q.x_lo = q.x_hi + pix * d->cin_p;
  q.dz_hi = reinterpret_cast<uint16_t*>(dz_head_split);
  q.dz_lo = q.dz_hi ? q.dz_hi + pix * 8 : nullptr;
  q.n = d->n; q.h = d->h; q.w_ = d->w; q.C = d->cin_p; q.out_bias = out_bias; q.p = p;
  q.inv_mean = target ? 1.0f / mean_pixels : 0.f;
  return launch_head_fwd(q, as_stream(stream));
}

extern "C" int nq_head_wgrad_blocks(const nq_conv_desc* d) {
  if (check_head(d)) return 0;
  const int64_t tiles = cdiv(d->w, HT_W) * cdiv(d->h, HT_H) * d->n;
  const int64_t cap = (int64_t)sm_count() * 4;
  return (int)(tiles < cap ? tiles : cap);
}

extern "C" int nq_head_wgrad(const nq_conv_desc* d, const float* x, const float* dz_head, float* dwk_head,
                             float* workspace, int64_t workspace_floats, void* stream) {
  int st = check_head(d);
  if (st) return st;
  if (!x || !dz_head || !dwk_head || !workspace) return NQ_ERR_BAD_ARG;
  if (d->cin_p > H_CH * HW_MAXCHUNK) return NQ_ERR_UNSUPPORTED;
  const int blocks = nq_head_wgrad_blocks(d);
  const int64_t rows4 = (int64_t)(9 * d->cin_p + 4) * 4;
  if (workspace_floats < rows4 * blocks) return NQ_ERR_WORKSPACE;
  HeadWgradParams q{};
  q.x = x; q.dz = dz_head; q.out = workspace;
  q.n = d->n; q.h = d->h; q.w_ = d->w; q.C = d->cin_p;
  q.tiles = (int)(cdiv(d->w, HT_W) * cdiv(d->h, HT_H) * d->n);
  cudaStream_t s = as_stream(stream);
  head_wgrad_kernel<<<blocks, HT_H * HT_W, 0, s>>>(q);
  NQ_LAUNCH_CHECK();
  const int64_t n4 = rows4 / 4;
  splitk_reduce_kernel<<<(unsigned)cdiv(n4, 256), 256, 0, s>>>(workspace, n4, blocks, dwk_head);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}
