// Head forward (3x3 conv to 3 channels + OutImg + lp_loss + dL/dz) on the tensor cores, "tap-expanded".
// Replaces the head's F.conv2d + OutImg (quant_layer.py:80, models/_layers.py:10-16) and lp_loss (quantizer.py:66-73).
//
// The head has 3 output channels: as a GEMM with N = 3 (padded to 16) every MMA re-reads a 4 KB activation tile for 3
// useful columns, 9 taps x 3 channel blocks x 3 passes = 81 MMAs per 128 pixels (nq_tc_head_fwd_loss), and as FFMA it is
// 1080 multiply-adds per pixel behind a shared-memory pipe (head_fwd_loss_strip_kernel).  Here the taps move from K to
// N:   D[pixel q][(tap t, channel o)] = sum_c x[q][c] * w[t][c][o]          N = 27 (padded 32), K = C, NO halo in A
// and the convolution is finished by nine shifted adds   y[p][o] = sum_t D[p + offset(t)][(t, o)].
// 9 MMAs (3 channel blocks x 3 split passes) of N = 32 per 128 INPUT pixels instead of 81; the adds are 27 per pixel.
//
// A CTA step: a 16 x 32 block of input pixels (4 MMA tiles of 4 rows; linear pixel order, so the 8-row groups of the
// K-major no-swizzle operand are simply contiguous) -> the 14 x 30 outputs inside it.  The tiles stream through a ring
// of six 128-pixel buffers, so the copies of the next tiles (and of the next block) run under the MMAs of this one.
// Roles (576 threads): warps 0-7 move the accumulators TMEM -> shared staging [pixel][33] and then finish the outputs (bias, OutImg, loss, dL/dz), warps
// 8-15 copy the split-bf16 activations with cp.async, warp 16 issues the MMAs, warp 17 owns the TMEM allocation.
// Accumulators are double buffered in TMEM (2 x 128 columns): the epilogue of block i overlaps the loads and MMAs of
// block i+1.  Weights (5.8 KB fp32) are split into bf16 hi / lo planes in shared memory once per CTA.
#include <cuda_bf16.h>
#include <stdint.h>
#include "nq_common.cuh"

namespace nq {

constexpr int HT_RH = 16, HT_RW = 32;             // input block (with the 1-pixel apron of its outputs)
constexpr int HT_NPIX = HT_RH * HT_RW;            // 512 = 4 MMA tiles
constexpr int HT_OH = HT_RH - 2, HT_OW = HT_RW - 2;
constexpr int HT_THREADS = 576;
constexpr int HT_NB = 6;                          // ring of 128-pixel activation buffers (one MMA tile each)
constexpr int HT_CGS = 128 * 16 + 64;             // bytes per 8-channel group of a tile buffer (+64: lane-pair stores conflict free)
constexpr int HT_STG = 33;                        // floats per staged pixel row (27 used; odd stride: conflict-free scalar access)
constexpr int HT_MAXC = 64;

struct HeadTcParams {
  const uint8_t* x;          // split-bf16 planes (n, h, w, Cs)
  size_t x_plane_bytes;
  const float* wk;           // [9][Cs][4] fp32 head weights as nq_pack_weight lays them out
  const float* bias;         // [>= 3]
  const float* target;       // (n, 3, h, w) or null
  const uint8_t* target_u8;  // (n, 3, h, w) uint8 frames (value / 255, datasets.py:8-54) or null; at most one of the two
  float* img;                // (n, 3, h, w) or null
  float* loss_sum;
  uint8_t* dz;               // split-bf16 planes (n, h, w, 8) or null
  size_t dz_plane_bytes;
  int n, h, w, Cs, C16;      // stored channels (multiple of 8), GEMM K (multiple of 16)
  int tiles_x, tiles_y, total;
  int out_bias;
  float p, inv_mean;
};

__device__ __forceinline__ uint32_t hsm(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void hbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void hbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void hcp16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hcp_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool helect() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void hmma(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                     uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void hcommit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(HT_THREADS, 1) head_tapexp_kernel(const __grid_constant__ HeadTcParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  // uint8 targets: v / 255 with an IEEE fp32 division (bit-identical to the reference's `read_image / 255.0`), tabulated once
  __shared__ float u8lut[256];
  if (threadIdx.x < 256) u8lut[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.0f);
  // [0,256) barriers + TMEM pointer | HT_NB activation tile buffers (2 planes each) | B planes | staging
  const uint32_t bar0 = hsm(smem);
  const uint32_t A_FULL = bar0, A_EMPTY = bar0 + HT_NB * 8, T_FULL = bar0 + 2 * HT_NB * 8, T_EMPTY = T_FULL + 16;  // T_*: two slots
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 240);
  const int ncg = p.C16 / 8;
  const int a_plane = ncg * HT_CGS;
  const int a_tile = 2 * a_plane;
  const int b_plane = ncg * 32 * 16;
  uint8_t* a_buf = smem + 256;
  uint8_t* b_buf = a_buf + HT_NB * a_tile;
  float* stg = reinterpret_cast<float*>(b_buf + 2 * b_plane);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < HT_NB; ++i) {
      hbar_init(A_FULL + i * 8, 256);
      hbar_init(A_EMPTY + i * 8, 1);
    }
    for (int i = 0; i < 2; ++i) {
      hbar_init(T_FULL + i * 8, 1);
      hbar_init(T_EMPTY + i * 8, 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(hsm(tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // weights -> bf16 hi / lo planes [k-group][32 columns][8 channels]; column n = tap * 3 + channel (27 used)
  for (int e = threadIdx.x; e < ncg * 32 * 8; e += HT_THREADS) {
    const int i = e & 7, n = (e >> 3) & 31, kg = e >> 8;
    const int c = kg * 8 + i, t = n / 3, o = n - t * 3;
    const float v = (n < 27 && c < p.Cs) ? p.wk[((size_t)t * p.Cs + c) * 4 + o] : 0.f;
    const __nv_bfloat16 hv = __float2bfloat16_rn(v);
    reinterpret_cast<__nv_bfloat16*>(b_buf)[e] = hv;
    reinterpret_cast<__nv_bfloat16*>(b_buf + b_plane)[e] = __float2bfloat16_rn(v - __bfloat162float(hv));
  }
  // channel groups beyond the stored channels (C16 > Cs) are never copied: zero them once
  for (int e = threadIdx.x; e < (p.C16 - p.Cs) / 8 * (HT_CGS / 16) * 2 * HT_NB; e += HT_THREADS) {
    const int per = (p.C16 - p.Cs) / 8 * (HT_CGS / 16), pl = e / per, r = e % per;  // pl = (buffer, plane)
    reinterpret_cast<uint4*>(a_buf + pl * a_plane + (p.Cs / 8) * HT_CGS)[r] = make_uint4(0u, 0u, 0u, 0u);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 16) {
    // ===================== MMA issuer =====================
    const bool leader = helect();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_hi32 = (128u >> 4) | (1u << 14);   // SBO: 8 consecutive pixels = 128 bytes
    const uint32_t b_hi32 = (128u >> 4) | (1u << 14);   // SBO: 8 columns = 128 bytes
    const uint32_t a0 = ((hsm(a_buf) & 0x3FFFFu) >> 4) | ((uint32_t)(HT_CGS >> 4) << 16);
    const uint32_t b0 = ((hsm(b_buf) & 0x3FFFFu) >> 4) | ((uint32_t)(32 * 16 >> 4) << 16);
    const uint32_t a_plane16 = (uint32_t)a_plane >> 4, b_plane16 = (uint32_t)b_plane >> 4;
    const int nk = p.C16 / 16;
    uint32_t it = 0, ab = 0, aph = 0;
    for (int t = blockIdx.x; t < p.total; t += gridDim.x, ++it) {
      const uint32_t acc = it & 1;
      hbar_wait(T_EMPTY + acc * 8, ((it >> 1) & 1) ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        hbar_wait(A_FULL + ab * 8, aph);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem_base + acc * 128 + j * 32;
        const uint32_t at = a0 + ab * ((uint32_t)a_tile >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k >= nk) break;
          const uint32_t a = at + (uint32_t)k * 2u * (HT_CGS >> 4);
          const uint32_t b = b0 + (uint32_t)k * 2u * (32 * 16 >> 4);
          if (leader) {
            hmma(d, a, a_hi32, b, b_hi32, idesc, k ? 1u : 0u);
            hmma(d, a + a_plane16, a_hi32, b, b_hi32, idesc, 1);
            hmma(d, a, a_hi32, b + b_plane16, b_hi32, idesc, 1);
          }
        }
        if (leader) hcommit(A_EMPTY + ab * 8);
        if (++ab == HT_NB) { ab = 0; aph ^= 1; }
      }
      if (leader) hcommit(T_FULL + acc * 8);
    }
  } else if (warp >= 8 && warp < 16) {
    // ===================== loaders: split-bf16 NHWC -> [channel group][pixel][8] =====================
    const int ltid = threadIdx.x - 256;
    const int ncs = p.Cs / 8, npair = (ncs + 1) >> 1, tasks = 128 * npair;
    const uint32_t a_sm = hsm(a_buf);
    uint32_t ab = 0, aph = 0;
    for (int t = blockIdx.x; t < p.total; t += gridDim.x) {
      const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, img = t / (p.tiles_x * p.tiles_y);
      const int y0 = ty * HT_OH - 1, x0 = tx * HT_OW - 1;
      for (int mt = 0; mt < 4; ++mt) {  // MMA tile mt = block rows 4 mt .. 4 mt + 3
        hbar_wait(A_EMPTY + ab * 8, aph ^ 1);
        const uint32_t dst = a_sm + ab * a_tile;
        for (int j = ltid >> 1; j < tasks; j += 128) {
          const int cpi = j >> 7, pix = j & 127;
          const int cg = 2 * cpi + (ltid & 1);
          if (cg < ncs) {
            const int ry = 4 * mt + (pix >> 5), rx = pix & 31;
            const int gy = y0 + ry, gx = x0 + rx;
            const bool ok = (unsigned)gy < (unsigned)p.h && (unsigned)gx < (unsigned)p.w;
            const uint8_t* src = ok ? p.x + ((size_t)((img * p.h + gy) * p.w + gx) * p.Cs + cg * 8) * 2 : p.x;
            const uint32_t d = dst + cg * HT_CGS + pix * 16;
            hcp16(d, src, ok ? 16u : 0u);
            hcp16(d + a_plane, src + p.x_plane_bytes, ok ? 16u : 0u);
          }
        }
        hcp_arrive(A_FULL + ab * 8);
        if (++ab == HT_NB) { ab = 0; aph ^= 1; }
      }
    }
  } else if (warp < 8) {
    // ===================== epilogue: TMEM -> staging, then the nine shifted adds per output =====================
    const int q = warp & 3, half = warp >> 2;
    const int64_t plane = (int64_t)p.h * p.w;
    const bool has_target = p.target != nullptr || p.target_u8 != nullptr;
    const float b0 = p.bias[0], b1 = p.bias[1], b2 = p.bias[2];
    float loss = 0.f;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.total; t += gridDim.x, ++it) {
      const uint32_t acc = it & 1;
      const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, img = t / (p.tiles_x * p.tiles_y);
      // the target pixels of this thread's (at most two) outputs are requested before the accumulator wait: the loss
      // epilogue otherwise pays the DRAM latency of these loads once per output round
      // (uint8 targets stay raw bytes here -- converting would wait for the load -- and go through the table below)
      uint32_t tgraw[2][3];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int o = threadIdx.x + 256 * r;
        const int oy = o / HT_OW, ox = o - oy * HT_OW;
        const int py = ty * HT_OH + oy, px = tx * HT_OW + ox;
        const bool ok = has_target && o < HT_OH * HT_OW && py < p.h && px < p.w;
        const int64_t off = (int64_t)img * 3 * plane + (int64_t)py * p.w + px;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          tgraw[r][c] = !ok ? 0u : (p.target ? __float_as_uint(__ldg(p.target + off + c * plane))
                                             : (uint32_t)__ldg(p.target_u8 + off + c * plane));
      }
      hbar_wait(T_FULL + acc * 8, (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float tg[2][3];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) tg[r][c] = p.target_u8 ? u8lut[tgraw[r][c] & 255u] : __uint_as_float(tgraw[r][c]);
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = half + 2 * jj;
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
            "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(tmem_base + acc * 128 + j * 32 + ((uint32_t)(q * 32) << 16))
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float* row = stg + (j * 128 + q * 32 + lane) * HT_STG;
#pragma unroll
        for (int c = 0; c < 27; ++c) row[c] = __uint_as_float(v[c]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) hbar_arrive(T_EMPTY + acc * 8);
      asm volatile("bar.sync 1, 256;" ::: "memory");  // staging complete (epilogue warps only)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int o = threadIdx.x + 256 * r;
        if (o >= HT_OH * HT_OW) continue;
        const int oy = o / HT_OW, ox = o - oy * HT_OW;
        const int py = ty * HT_OH + oy, px = tx * HT_OW + ox;
        if (py >= p.h || px >= p.w) continue;
        float y0 = b0, y1 = b1, y2 = b2;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float* r = stg + ((oy + kh) * HT_RW + ox + kw) * HT_STG + (kh * 3 + kw) * 3;
            y0 += r[0]; y1 += r[1]; y2 += r[2];
          }
        const float v[3] = {y0, y1, y2};
        float g[3] = {0.f, 0.f, 0.f};
        const int64_t off = (int64_t)img * 3 * plane + (int64_t)py * p.w + px;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float outv, dout;
          if (p.out_bias == 0) {
            const float th = tanhf(v[c]);
            outv = th * 0.5f + 0.5f;
            dout = 0.5f * (1.0f - th * th);
          } else {
            outv = sigmoid_f(v[c]);
            dout = outv * (1.0f - outv);
          }
          if (p.img) p.img[off + c * plane] = outv;
          if (has_target) {
            const float dlt = outv - tg[r][c];
            const float a = fabsf(dlt);
            if (p.p == 2.0f) {
              loss += dlt * dlt;
              g[c] = 2.0f * dlt * p.inv_mean * dout;
            } else {
              loss += powf(a, p.p);
              const float sgn = dlt > 0.f ? 1.f : (dlt < 0.f ? -1.f : 0.f);
              g[c] = p.p * powf(a, p.p - 1.0f) * sgn * p.inv_mean * dout;
            }
          }
        }
        if (p.dz) {
          uint32_t hb[3], lb[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const __nv_bfloat16 hv = __float2bfloat16_rn(g[c]);
            hb[c] = __bfloat16_as_ushort(hv);
            lb[c] = __bfloat16_as_ushort(__float2bfloat16_rn(g[c] - __bfloat162float(hv)));
          }
          const size_t pix = ((size_t)(img * p.h + py) * p.w + px) * 16;
          *reinterpret_cast<uint4*>(p.dz + pix) = make_uint4(hb[0] | (hb[1] << 16), hb[2], 0u, 0u);
          *reinterpret_cast<uint4*>(p.dz + p.dz_plane_bytes + pix) = make_uint4(lb[0] | (lb[1] << 16), lb[2], 0u, 0u);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // staging free for the next block
    }
    if (has_target && p.loss_sum) {
      loss = warp_sum(loss);
      if (lane == 0 && loss != 0.f) atomicAdd(p.loss_sum, loss);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 17) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

int check_conv_desc(const nq_conv_desc* d);

}  // namespace nq

using namespace nq;

// Same contract as nq_head_fwd_loss_split (include/neuroquant_b200.h).
static int head_tapexp_launch(const nq_conv_desc* d, const void* x_split, const float* w_head, const float* bias_head,
                              int out_bias, const float* target_f32, const uint8_t* target_u8, float p, float mean_pixels, float* img,
                              float* loss_sum, void* dz_head_split, void* stream) {
  const void* target = target_f32 ? (const void*)target_f32 : (const void*)target_u8;
  int st = check_conv_desc(d);
  if (st) return st;
  if (d->ksize != 3 || d->rh != 1 || d->rw != 1 || d->cout != 3 || d->cg != 4) return NQ_ERR_BAD_SHAPE;
  if (!x_split || !w_head || !bias_head) return NQ_ERR_BAD_ARG;
  if (out_bias != 0 && out_bias != 1) return NQ_ERR_UNSUPPORTED;
  if (target && (!(p > 0.f) || !(mean_pixels > 0.f))) return NQ_ERR_BAD_ARG;
  if (!target && !img) return NQ_ERR_BAD_ARG;
  if (d->cin_p % 8 || d->cin_p > HT_MAXC) return NQ_ERR_UNSUPPORTED;
  HeadTcParams q{};
  const size_t pix = (size_t)d->n * d->h * d->w;
  q.x = reinterpret_cast<const uint8_t*>(x_split);
  q.x_plane_bytes = pix * d->cin_p * 2;
  q.wk = w_head; q.bias = bias_head; q.target = target_f32; q.target_u8 = target_u8; q.img = img; q.loss_sum = loss_sum;
  q.dz = reinterpret_cast<uint8_t*>(dz_head_split);
  q.dz_plane_bytes = pix * 16;
  q.n = d->n; q.h = d->h; q.w = d->w; q.Cs = d->cin_p; q.C16 = (d->cin_p + 15) / 16 * 16;
  q.tiles_x = (d->w + HT_OW - 1) / HT_OW; q.tiles_y = (d->h + HT_OH - 1) / HT_OH;
  const long long total = (long long)q.tiles_x * q.tiles_y * d->n;
  if (total >= (1LL << 31) || pix >= (1ULL << 31)) return NQ_ERR_BAD_SHAPE;
  q.total = (int)total;
  q.out_bias = out_bias; q.p = p; q.inv_mean = target ? 1.0f / mean_pixels : 0.f;
  const int ncg = q.C16 / 8;
  const int smem = 256 + HT_NB * 2 * ncg * HT_CGS + 2 * ncg * 32 * 16 + HT_NPIX * HT_STG * 4;
  if (smem + 1024 > 227 * 1024) return NQ_ERR_UNSUPPORTED;  // + the static 1 KB uint8 table
  NQ_CUDA_CHECK(cudaFuncSetAttribute(head_tapexp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int grid = sm_count();
  if (grid > q.total) grid = q.total;
  head_tapexp_kernel<<<grid, HT_THREADS, smem, as_stream(stream)>>>(q);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

extern "C" int nq_head_fwd_loss_tapexp(const nq_conv_desc* d, const void* x_split, const float* w_head, const float* bias_head,
                                       int out_bias, const float* target, float p, float mean_pixels, float* img,
                                       float* loss_sum, void* dz_head_split, void* stream) {
  return head_tapexp_launch(d, x_split, w_head, bias_head, out_bias, target, nullptr, p, mean_pixels, img, loss_sum, dz_head_split, stream);
}

// Same with the target frames as the data set stores them: uint8 (n, 3, h, w); the kernel evaluates value / 255 in fp32
// (IEEE division, bit-identical to `read_image(...) / 255.0`, videosets/datasets.py:8-54) -- a quarter of the bytes over
// PCIe and out of HBM.
extern "C" int nq_head_fwd_loss_tapexp_u8(const nq_conv_desc* d, const void* x_split, const float* w_head, const float* bias_head,
                                          int out_bias, const uint8_t* target_u8, float p, float mean_pixels, float* img,
                                          float* loss_sum, void* dz_head_split, void* stream) {
  if (!target_u8) return NQ_ERR_BAD_ARG;
  return head_tapexp_launch(d, x_split, w_head, bias_head, out_bias, nullptr, target_u8, p, mean_pixels, img, loss_sum, dz_head_split, stream);
}

// ================================================================================================
// Head weight gradient, tap-expanded:   dW[t][c][o] = sum_p x[p + off(t)][c] * dz[p][o]
//                                                   = sum_q x[q][c] * dzs[q][(t, o)],   dzs[q][(t, o)] = dz[q - off(t)][o]
// One GEMM with the INPUT pixels q as K, the channels as M (plain copy of the activation tile: no shifted copies, no
// halo rows) and the 27 (tap, channel) pairs as N; the nine shifts move to the tiny dZ operand (3 channels), which eight
// warps assemble in shared memory from a raw halo tile.  Per 16 pixels: 2 MMAs (x_hi x [dzs_hi | dzs_lo], x_lo x dzs_hi)
// instead of 6 in wgrad_tc_kernel, and 4x less shared-memory traffic.  One extra row of ones gives the bias gradient
// (column of the centre tap).  Persistent CTAs over row segments of 128 pixels; accumulators stay in TMEM; every CTA
// writes one partial in the layout nq_tc_wgrad_finish_multi reads ((9 * cin_p + 4) rows x 16 columns).
// ================================================================================================
namespace nq {

constexpr int WT_TW = 128, WT_THREADS = 576, WT_NB = 2;
constexpr int WT_CGS = WT_TW * 16 + 64;          // bytes per 8-row group (A: 8 channels, B: 8 columns) of a tile buffer
constexpr int WT_RAW_ROW = (WT_TW + 2) * 16;     // raw dZ halo row: 130 pixels x (8 channels bf16)
constexpr int WT_A_PLANE = 16 * WT_CGS, WT_B_PLANE = 4 * WT_CGS, WT_RAW_PLANE = 3 * WT_RAW_ROW;

struct HeadWgParams {
  const uint8_t* x; size_t x_plane_bytes;     // split-bf16 (n, h, w, Cs)
  const uint8_t* dz; size_t dz_plane_bytes;   // split-bf16 (n, h, w, 8)
  float* ws;                                  // (grid, 9 * Cs + 4, 16) partial sums
  int n, h, w, Cs;
  int tiles_x, total, tiles_per_cta;
};

__global__ void __launch_bounds__(WT_THREADS, 1) head_wgrad_tapexp_kernel(const __grid_constant__ HeadWgParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t bar0 = hsm(smem);
  const uint32_t LOAD_FULL = bar0, B_FULL = bar0 + 16, SLOT_EMPTY = bar0 + 32, DONE = bar0 + 48;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 64);
  uint8_t* a_buf = smem + 256;                                  // [slot][plane][16 groups][WT_CGS]
  uint8_t* b_buf = a_buf + WT_NB * 2 * WT_A_PLANE;              // [slot][plane][4 groups][WT_CGS]
  uint8_t* raw = b_buf + WT_NB * 2 * WT_B_PLANE;                // [slot][plane][3 rows][130 px][16 B]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncs = p.Cs / 8;
  const int t_begin = blockIdx.x * p.tiles_per_cta, t_end = min(p.total, t_begin + p.tiles_per_cta);

  if (threadIdx.x == 0) {
    for (int i = 0; i < WT_NB; ++i) {
      hbar_init(LOAD_FULL + i * 8, 256);
      hbar_init(B_FULL + i * 8, 8);
      hbar_init(SLOT_EMPTY + i * 8, 1);
    }
    hbar_init(DONE, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(hsm(tmem_slot)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // constant rows of the activation operand: group ncs = ones in channel 0 (bias gradient), groups above = zero
  for (int e = threadIdx.x; e < WT_NB * 2 * (16 - ncs) * (WT_CGS / 16); e += WT_THREADS) {
    const int per = (16 - ncs) * (WT_CGS / 16), sp = e / per, r = e % per;  // sp = (slot, plane)
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if ((sp & 1) == 0 && r < WT_CGS / 16) v.x = 0x3F80u;  // bf16 1.0, hi plane, first constant group
    reinterpret_cast<uint4*>(a_buf + sp * WT_A_PLANE + ncs * WT_CGS)[r] = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 16) {
    // ===================== MMA issuer =====================
    const bool leader = helect();
    const uint32_t idesc32 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);   // D fp32, A/B bf16 both MN-major, M = 128
    const uint32_t idesc64 = (idesc32 & ~(0x3Fu << 17)) | ((uint32_t)(64 >> 3) << 17);
    const uint32_t hi32 = ((uint32_t)WT_CGS >> 4) | (1u << 14);  // SBO: next 8-row group
    const uint32_t lbo = (128u >> 4) << 16;                       // LBO: second 8-pixel half of a K = 16 step
    uint32_t it = 0;
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const uint32_t s = it % WT_NB, ph = (it / WT_NB) & 1;
      hbar_wait(LOAD_FULL + s * 8, ph);
      hbar_wait(B_FULL + s * 8, ph);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a16 = ((hsm(a_buf + s * 2 * WT_A_PLANE) & 0x3FFFFu) >> 4) | lbo;
      const uint32_t b16 = ((hsm(b_buf + s * 2 * WT_B_PLANE) & 0x3FFFFu) >> 4) | lbo;
#pragma unroll
      for (int k = 0; k < WT_TW / 16; ++k) {
        const uint32_t a = a16 + k * 16, b = b16 + k * 16;  // 16 pixels = 256 bytes
        if (leader) {
          hmma(tmem_base, a, hi32, b, hi32, idesc64, (it | k) ? 1u : 0u);
          hmma(tmem_base, a + (WT_A_PLANE >> 4), hi32, b, hi32, idesc32, 1);
        }
      }
      if (leader) hcommit(SLOT_EMPTY + s * 8);
    }
    if (leader) hcommit(DONE);
  } else if (warp >= 8 && warp < 16) {
    // ===================== loaders: activation tile + raw dZ halo =====================
    const int ltid = threadIdx.x - 256;
    const int npair = (ncs + 1) >> 1;
    uint32_t it = 0;
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const uint32_t s = it % WT_NB, ph = (it / WT_NB) & 1;
      const int tx = t % p.tiles_x, row = t / p.tiles_x;  // row = img * h + y
      const int y = row % p.h, x0 = tx * WT_TW;
      hbar_wait(SLOT_EMPTY + s * 8, ph ^ 1);
      const uint32_t a_sm = hsm(a_buf + s * 2 * WT_A_PLANE);
      for (int j = ltid >> 1; j < WT_TW * npair; j += 128) {
        const int cpi = j >> 7, px = j & 127;
        const int cg = 2 * cpi + (ltid & 1);
        if (cg < ncs) {
          const bool ok = x0 + px < p.w;
          const uint8_t* src = ok ? p.x + ((size_t)(row * p.w + x0 + px) * p.Cs + cg * 8) * 2 : p.x;
          const uint32_t d = a_sm + cg * WT_CGS + px * 16;
          hcp16(d, src, ok ? 16u : 0u);
          hcp16(d + WT_A_PLANE, src + p.x_plane_bytes, ok ? 16u : 0u);
        }
      }
      const uint32_t r_sm = hsm(raw + s * 2 * WT_RAW_PLANE);
      for (int j = ltid; j < 3 * (WT_TW + 2); j += 256) {
        const int ry = j / (WT_TW + 2), rx = j - ry * (WT_TW + 2);
        const int gy = y + ry - 1, gx = x0 + rx - 1;
        const bool ok = (unsigned)gy < (unsigned)p.h && (unsigned)gx < (unsigned)p.w;
        const uint8_t* src = ok ? p.dz + (size_t)((row + ry - 1) * p.w + gx) * 16 : p.dz;
        const uint32_t d = r_sm + ry * WT_RAW_ROW + rx * 16;
        hcp16(d, src, ok ? 16u : 0u);
        hcp16(d + WT_RAW_PLANE, src + p.dz_plane_bytes, ok ? 16u : 0u);
      }
      hcp_arrive(LOAD_FULL + s * 8);
    }
  } else if (warp < 8) {
    // ===================== dZ operand builders: dzs[q][(t, o)] = dz[q - off(t)][o], columns n = 3 t + o =====================
    uint32_t it = 0;
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const uint32_t s = it % WT_NB, ph = (it / WT_NB) & 1;
      hbar_wait(LOAD_FULL + s * 8, ph);
      const uint16_t* rw = reinterpret_cast<const uint16_t*>(raw + s * 2 * WT_RAW_PLANE);
      uint8_t* bb = b_buf + s * 2 * WT_B_PLANE;
      {  // one thread = one pixel of one plane: all 27 (tap, channel) columns, every index a compile-time constant
        const int px = threadIdx.x & 127, pl = threadIdx.x >> 7;
        const uint16_t* rp = rw + (pl * WT_RAW_PLANE + (px + 2) * 16) / 2;
        uint8_t* dst = bb + pl * WT_B_PLANE + px * 16;
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
          uint32_t w4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int nn = gi * 8 + j;
            if (nn < 27) {
              const int tt = nn / 3, o = nn - tt * 3, kh = tt / 3, kw = tt - kh * 3;
              const uint32_t v = rp[((2 - kh) * WT_RAW_ROW - kw * 16) / 2 + o];
              w4[j >> 1] |= v << (16 * (j & 1));
            }
          }
          *reinterpret_cast<uint4*>(dst + gi * WT_CGS) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) hbar_arrive(B_FULL + s * 8);
    }
    // ===================== epilogue (warps 0-3): accumulator -> this CTA's partial =====================
    const int rows_total = 9 * p.Cs + 4;
    float* out = p.ws + (size_t)blockIdx.x * rows_total * 16;
    if (warp < 4) {
      for (int e = threadIdx.x; e < rows_total * 4; e += 128) reinterpret_cast<float4*>(out)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (t_end > t_begin) {
        hbar_wait(DONE, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int rowm = warp * 32 + lane;  // accumulator row = channel (or the ones row)
        uint32_t v[32], v2[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
            "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
            "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v2[0]), "=r"(v2[1]), "=r"(v2[2]), "=r"(v2[3]), "=r"(v2[4]), "=r"(v2[5]), "=r"(v2[6]), "=r"(v2[7]), "=r"(v2[8]),
              "=r"(v2[9]), "=r"(v2[10]), "=r"(v2[11]), "=r"(v2[12]), "=r"(v2[13]), "=r"(v2[14]), "=r"(v2[15]), "=r"(v2[16]),
              "=r"(v2[17]), "=r"(v2[18]), "=r"(v2[19]), "=r"(v2[20]), "=r"(v2[21]), "=r"(v2[22]), "=r"(v2[23]), "=r"(v2[24]),
              "=r"(v2[25]), "=r"(v2[26]), "=r"(v2[27]), "=r"(v2[28]), "=r"(v2[29]), "=r"(v2[30]), "=r"(v2[31])
            : "r"(taddr + 32)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const bool is_ch = rowm < p.Cs, is_ones = rowm == p.Cs;
        if (is_ch || is_ones) {
#pragma unroll
          for (int nn = 0; nn < 27; ++nn) {
            const float val = __uint_as_float(v[nn]) + __uint_as_float(v2[nn]);
            const int tt = nn / 3, o = nn - tt * 3;
            if (is_ch) out[((size_t)tt * p.Cs + rowm) * 16 + o] = val;
            else if (tt == 4) out[(size_t)9 * p.Cs * 16 + o] = val;  // bias gradient: unshifted (centre-tap) column
          }
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 17) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64) : "memory");
  }
}

}  // namespace nq

extern "C" int nq_head_wgrad_tapexp_splits(const nq_conv_desc* d) {
  if (check_conv_desc(d)) return 0;
  const long long total = (long long)d->n * d->h * ((d->w + WT_TW - 1) / WT_TW);
  return (int)(total < sm_count() ? total : sm_count());
}

extern "C" int nq_head_wgrad_tapexp(const nq_conv_desc* d, const void* x_split, const void* dz_split, float* workspace,
                                    int64_t workspace_floats, void* stream) {
  int st = check_conv_desc(d);
  if (st) return st;
  if (d->ksize != 3 || d->rh != 1 || d->rw != 1 || d->cout != 3) return NQ_ERR_BAD_SHAPE;
  if (!x_split || !dz_split || !workspace) return NQ_ERR_BAD_ARG;
  if (d->cin_p % 8 || d->cin_p > HT_MAXC) return NQ_ERR_UNSUPPORTED;
  const int grid = nq_head_wgrad_tapexp_splits(d);
  const int rows_total = 9 * d->cin_p + 4;
  if (workspace_floats < (int64_t)grid * rows_total * 16) return NQ_ERR_WORKSPACE;
  const size_t pix = (size_t)d->n * d->h * d->w;
  if (pix >= (1ULL << 31)) return NQ_ERR_BAD_SHAPE;
  HeadWgParams q{};
  q.x = reinterpret_cast<const uint8_t*>(x_split); q.x_plane_bytes = pix * d->cin_p * 2;
  q.dz = reinterpret_cast<const uint8_t*>(dz_split); q.dz_plane_bytes = pix * 16;
  q.ws = workspace; q.n = d->n; q.h = d->h; q.w = d->w; q.Cs = d->cin_p;
  q.tiles_x = (d->w + WT_TW - 1) / WT_TW;
  q.total = d->n * d->h * q.tiles_x;
  q.tiles_per_cta = (q.total + grid - 1) / grid;
  const int smem = 256 + WT_NB * 2 * (WT_A_PLANE + WT_B_PLANE + WT_RAW_PLANE);
  NQ_CUDA_CHECK(cudaFuncSetAttribute(head_wgrad_tapexp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  head_wgrad_tapexp_kernel<<<grid, WT_THREADS, smem, as_stream(stream)>>>(q);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}
