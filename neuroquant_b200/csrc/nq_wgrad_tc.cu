// tcgen05 / TMEM weight (+ bias) gradient of a stride-1 "same" convolution, NHWC fp32 in, fp32 out.
// Replaces the wgrad half of autograd(F.conv2d) (quant_layer.py:80) inside the calibration loop
// (calib_model.py:160-162, :221-223).
//
//   dW[(kh,kw)][ci][n'] = sum_pixels X[pixel + (kh,kw)][ci] * dZ[pixel][n']        (+ db[n'] = sum dZ)
//
// The reduction runs over pixels, so pixels are the MMA K dimension and both operands are "MN-major"
// (channels contiguous) -- which NHWC gives for free: an 8-channel group of 8 consecutive x positions is
// exactly one 128-byte no-swizzle core matrix.  Mapping:
//   * a CTA owns a group of consecutive kernel rows kh, one slice of the output channels n' and a contiguous
//     range of 16-pixel-wide x TR-row pixel tiles; its accumulators (one per kh and 128-row block, side by
//     side in the 512 TMEM columns) stay in TMEM for the whole range (persistent split-K), then are written
//     once as a partial dW and summed in a fixed order.  Grouping kernel rows lets one staged tile of dZ and of
//     the shifted inputs feed several kh (the kh shift is a row offset in the descriptor): the L2 -> SM traffic,
//     which bounds this kernel, drops by the group size
//   * A (GEMM M) = shifted input: rows (kw, channel group, 8 channels).  The ks horizontal shifts are
//     materialised as ks copies of the 16-pixel row segment in shared memory ([kw][group][row][x][8]),
//     which makes the 8-row groups uniformly strided, so M = 128 covers 16 (kw, group) pairs per MMA
//   * one extra A group holds the constant 1 in channel 0: its accumulator row is the bias gradient
//   * B (GEMM N) = dZ tile [group][row][x][8]; K = 16 pixels = one row segment per MMA
//   * operands are bf16 hi (+ lo) planes of the fp32 tensors, accumulation fp32
#include <cuda_bf16.h>

#include <stdio.h>
#include <stdlib.h>

#include "nq_common.cuh"

namespace nq {

constexpr int WG_THREADS = 640;
constexpr int WG_LOADERS = 512;  // warps 0-15
constexpr int WG_TW = 16;  // pixels per MMA K step

struct WgParams {
  const uint8_t* x;   // split-bf16 (hi plane, lo plane), each (n, h, w, C) bf16
  const uint8_t* dz;  // split-bf16, each (n, h, w, dz_stride) bf16
  size_t x_plane_bytes, dz_plane_bytes;
  float* ws;        // [psplits][(ks*ks*C + 4)][N] partial gradients (bias row at ks*ks*C)
  int n, h, w, C, N, ks, pad;
  int ncg, G, MB;   // channel groups, (kw, group) pairs, 128-row blocks
  int NC, nsplits;  // output columns per CTA
  int TR;           // tile rows
  int nkh, khg, AR; // kernel rows per CTA, number of kernel-row groups, staged input rows (TR + nkh - 1)
  int msplit, ncg_c; // channel-group slices of the input (GEMM-M split across CTAs), groups per slice
  int bcat;          // the two dZ planes (adjacent in the column-group dimension of a buffer) as ONE operand of 2 NC columns
  int tiles_x, tiles_y, tiles_total, psplits, tiles_per_split;
  int a_planes, b_planes;
  int CGS_A, CGS_B, a_plane_bytes, b_plane_bytes, buf_bytes, nbuf;
  int dz_stride, n_valid;  // channels per pixel stored in dz (<= N); columns >= n_valid are zero
  int a_sts;               // shifted input copies staged through registers (one global load, <= ks st.shared) instead of ks cp.async
  int skip;                // NQ_WG_SKIP (debug): bit 0 shifted-input copies, bit 1 dZ copies (bit 2, MMAs, is masked off at launch)
  long long* dbg;          // NQ_TC_DBG: {SM cycles, ns} of CTA 0
};

// ---- PTX helpers (same conventions as nq_conv_tc.cu) ----
__device__ __forceinline__ uint32_t wsmem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void wbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void wbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t wdesc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void wmma(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void wmma_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                       uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
      : "memory");
}
// same MMA, keeping the A operand in the collector for the next MMA (same A, other B) / taking it from there: the 4 KB
// A tile is read from shared memory once for both
#define NQ_WMMA(NAME, QUAL)                                                                                               \
  __device__ __forceinline__ void NAME(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,          \
                                       uint32_t idesc, uint32_t accum) {                                                    \
    asm volatile(                                                                                                           \
        "{\n\t"                                                                                                             \
        ".reg .pred p;\n\t"                                                                                                 \
        ".reg .b64 da, db;\n\t"                                                                                             \
        "setp.ne.b32 p, %6, 0;\n\t"                                                                                         \
        "mov.b64 da, {%1, %2};\n\t"                                                                                         \
        "mov.b64 db, {%3, %4};\n\t"                                                                                         \
        "tcgen05.mma.cta_group::1.kind::f16" QUAL " [%0], da, db, %5, p;\n\t"                                              \
        "}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)                               \
        : "memory");                                                                                                        \
  }
NQ_WMMA(wmma_w_keep, ".collector::a::fill")
NQ_WMMA(wmma_w_reuse, ".collector::a::lastuse")
#undef NQ_WMMA
__device__ __forceinline__ void wcp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void wcp_async16_ca(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void wcp_async_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool welect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void wcommit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void wsplit8(const float4& a, const float4& b, uint4& hi, uint4& lo) {
  const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * i]), h1 = __float2bfloat16_rn(x[2 * i + 1]);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(x[2 * i] - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(x[2 * i + 1] - __bfloat162float(h1));
    h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// BCAT / PASSES are compile-time copies of WgParams::bcat and of (a_planes == 2) | (b_planes == 2) << 1: the per-MMA tests
// of runtime flags were a third of the unrolled issue sequence (22 -> 15 instructions per 3 MMAs).  Measured effect on
// conv_wgrad[5] of this and of fetching x_hi once for both dz planes (A collector): none (0.480 -> 0.477 ms) -- that kernel
// sits at the shared-memory wavefront limit, tensor-core operand reads 50 % + cp.async writes 53 % of the SM's wavefront
// slots (profiles/r02r_wgrad.md); the collector moved the first share from 86 M to 63 M wavefronts.
template <int BCAT, int PASSES>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t bar0 = wsmem_u32(smem);
  const uint32_t FULL = bar0, EMPTY = bar0 + 4 * 8, DONE = bar0 + 8 * 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 9 * 8);
  uint8_t* bufs = smem + 128;
  const uint32_t buf0 = wsmem_u32(bufs);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // block -> (pixel split, column split, kernel-row group)
  int b = blockIdx.x;
  const int kh0 = (b % p.khg) * p.nkh;
  const int nkh = min(p.nkh, p.ks - kh0);
  b /= p.khg;
  const int nsplit = b % p.nsplits;
  b /= p.nsplits;
  const int cg0 = (b % p.msplit) * p.ncg_c;            // first input channel group of this CTA
  const int ncg_c = min(p.ncg_c, p.ncg - cg0);         // its channel groups
  const int Gc = p.ks * ncg_c;                         // its (kw, group) pairs
  const int psplit = b / p.msplit;
  const int n0 = nsplit * p.NC;
  const int nc = min(p.NC, p.N - n0);
  const int t_begin = psplit * p.tiles_per_split;
  const int t_end = min(p.tiles_total, t_begin + p.tiles_per_split);

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nbuf; ++i) {
      wbar_init(FULL + i * 8, WG_LOADERS);  // one deferred cp.async arrival per loader thread
      wbar_init(EMPTY + i * 8, 1);
    }
    wbar_init(DONE, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 18) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(wsmem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // constant part of every A buffer: the "ones" group (bias gradient) and zeroed tail groups
  {
    const int tail_groups = p.MB * 16 - Gc;  // >= 1
    const int per_group16 = p.AR * WG_TW;     // 16-byte units per group
    for (int bi = 0; bi < p.nbuf; ++bi)
      for (int pl = 0; pl < p.a_planes; ++pl) {
        uint8_t* base = bufs + (size_t)bi * p.buf_bytes + (size_t)pl * p.a_plane_bytes;
        for (int i = threadIdx.x; i < tail_groups * per_group16; i += WG_THREADS) {
          const int g = Gc + i / per_group16, u = i % per_group16;
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (g == Gc && pl == 0) v.x = 0x3F80u;  // bf16 1.0 in channel 0
          *reinterpret_cast<uint4*>(base + (size_t)g * p.CGS_A + (size_t)u * 16) = v;
        }
      }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_wait();     // everything above ran under the tail of the kernel before this one (nq_common.cuh)
  pdl_trigger();
  const uint32_t tmem_base = *tmem_slot;
  long long dbg_c0 = 0, dbg_t0 = 0;
  if (p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    dbg_c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
  }

  if (warp == 17) {
    // ===================== MMA issuer =====================
    // Highest warp id of its scheduler partition; warp-uniform loop with an elect.sync leader so that the
    // descriptors stay in uniform registers (see nq_conv_tc.cu).
    const bool leader = welect_one() && !(p.skip & 4);
    // D fp32, A/B bf16, both MN-major (bits 15, 16), M = 128, N = nc
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(nc >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc2 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)((2 * nc) >> 3) << 17);  // N = 2 nc
    const uint32_t acc_cols = (uint32_t)p.NC << BCAT;
    const uint32_t a_hi32 = ((uint32_t)p.CGS_A >> 4) | (1u << 14), b_hi32 = ((uint32_t)p.CGS_B >> 4) | (1u << 14);
    const uint32_t lbo_bits = (128u >> 4) << 16;
    const uint32_t a_plane16 = (uint32_t)p.a_plane_bytes >> 4, b_plane16 = (uint32_t)p.b_plane_bytes >> 4;
    const uint32_t mb_step16 = (uint32_t)(16 * p.CGS_A) >> 4, row16 = (WG_TW * 16) >> 4;
    constexpr int passes = PASSES;
    uint32_t accum = 0, bi = 0, ph = 0;
    for (int t = t_begin; t < t_end; ++t) {
      wbar_wait(FULL + bi * 8, ph);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // cp.async (generic proxy) -> MMA (async proxy)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_buf = buf0 + bi * p.buf_bytes;
      const uint32_t a16 = ((a_buf & 0x3FFFFu) >> 4) | lbo_bits;
      const uint32_t b16 = (((a_buf + p.a_planes * p.a_plane_bytes) & 0x3FFFFu) >> 4) | lbo_bits;
      if (p.MB == 1) {
        // One 128-row block: fully unrolled, every descriptor an independent add off the buffer base.  The loop nest
        // below carries them through dependent uniform-datapath adds, which binds the narrow-N plans (the head:
        // 36 MMAs of N = 16 per tile, ~100 cycles of issue each).
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (r >= p.TR) break;
#pragma unroll
          for (int khl = 0; khl < 7; ++khl) {
            if (khl >= nkh) break;
            const uint32_t a_lo = a16 + (r + khl) * row16, b_lo = b16 + r * row16, d = tmem_base + khl * acc_cols;
            if (leader) {
              if (BCAT) {  // x_hi * [dz_hi | dz_lo] in one MMA of 2 NC columns, then x_lo * dz_hi: 2 A-tile reads, not 3
                wmma_w(d, a_lo, a_hi32, b_lo, b_hi32, idesc2, r == 0 ? accum : 1u);
                wmma_w(d, a_lo + a_plane16, a_hi32, b_lo, b_hi32, idesc, 1);
              } else if (passes & 2) {  // x_hi meets both dz planes: fetched once (A collector)
                wmma_w_keep(d, a_lo, a_hi32, b_lo, b_hi32, idesc, r == 0 ? accum : 1u);
                wmma_w_reuse(d, a_lo, a_hi32, b_lo + b_plane16, b_hi32, idesc, 1);
                if (passes & 1) wmma_w(d, a_lo + a_plane16, a_hi32, b_lo, b_hi32, idesc, 1);
              } else {
                wmma_w(d, a_lo, a_hi32, b_lo, b_hi32, idesc, r == 0 ? accum : 1u);
                if (passes & 1) wmma_w(d, a_lo + a_plane16, a_hi32, b_lo, b_hi32, idesc, 1);
              }
            }
          }
        }
        accum = 1;
      } else
      for (int r = 0; r < p.TR; ++r) {
        const uint32_t b_lo = b16 + r * row16;
        uint32_t d = tmem_base;
        for (int khl = 0; khl < nkh; ++khl) {
          uint32_t a_lo = a16 + (r + khl) * row16;  // the kh shift is a row offset into the staged input rows
#pragma unroll 1
          for (int mb = 0; mb < p.MB; ++mb) {
            if (leader) {
              if (passes & 2) {
                wmma_w_keep(d, a_lo, a_hi32, b_lo, b_hi32, idesc, accum);
                wmma_w_reuse(d, a_lo, a_hi32, b_lo + b_plane16, b_hi32, idesc, 1);
              } else {
                wmma_w(d, a_lo, a_hi32, b_lo, b_hi32, idesc, accum);
              }
              if (passes & 1) wmma_w(d, a_lo + a_plane16, a_hi32, b_lo, b_hi32, idesc, 1);
            }
            a_lo += mb_step16;
            d += acc_cols;
          }
        }
        accum = 1;
      }
      if (leader) wcommit(EMPTY + bi * 8);
      if (++bi == (uint32_t)p.nbuf) { bi = 0; ph ^= 1; }
    }
    if (leader) wcommit(DONE);
  } else if (warp < WG_LOADERS / 32) {
    // ===================== loaders (16 warps): fp32 NHWC -> bf16 planes =====================
    const int ltid = threadIdx.x;
    // Lanes 2k and 2k+1 copy the two 16-byte halves (adjacent channel groups) of one 32-byte sector: full sector
    // efficiency on the L2 -> SM path, which bounds this kernel.  All index walking is incremental (no division).
    const int cgp = ltid & 1;
    const int xl = (ltid >> 1) & (WG_TW - 1);
    const int lgrp = ltid >> 5, nlgrp = WG_LOADERS >> 5;  // 16 pixels x 2 parities per warp
    const int npair = (ncg_c + 1) >> 1;
    const int QA = p.AR * p.ks * npair;                   // (input row, kw, group pair) items, pair fastest
    const int slots = p.TR * WG_TW;
    const int ncg_b = nc >> 3;
    // output-gradient tile: thread -> (pixel slot, parity), walks group pairs
    const int bslot = (ltid >> 1) & (slots - 1);
    const int bgrp = ltid / (2 * slots), nbgrp = WG_LOADERS / (2 * slots);
    const int br = bslot / WG_TW, bxl = bslot % WG_TW;
    // The (input row, kw, group) items a thread copies are the same for every tile: decode them ONCE into
    // registers (source offset relative to the tile origin, destination offset, dy, dx); the per-tile work is then
    // a bounds test and two adds per copy.  (The decode loop per tile made this kernel issue bound.)
    constexpr int MAXI = 4;  // (the register-staged path below covers the 3x3 and 5x5 stages; this one the 1x1 stages)
    int a_so[MAXI], a_do[MAXI], a_dyx[MAXI];  // source element offset, destination byte offset, (dy << 16) | (dx & 0xffff)
    int na = 0;
    bool a_table = true;
    {
      int cpi = lgrp, kw = 0, ar = 0;
      while (cpi >= npair) { cpi -= npair; if (++kw == p.ks) { kw = 0; ++ar; } }
      for (int q = lgrp; q < QA; q += nlgrp) {
        const int cg = 2 * cpi + cgp;
        if (cg < ncg_c) {
          if (na < MAXI) {
            const int dy = ar + kh0 - p.pad, dx = xl + kw - p.pad;
#pragma unroll
            for (int k = 0; k < MAXI; ++k)
              if (k == na) {
                a_so[k] = (dy * p.w + dx) * p.C + (cg0 + cg) * 8;
                a_do[k] = (kw * ncg_c + cg) * p.CGS_A + ar * (WG_TW * 16) + xl * 16;
                a_dyx[k] = (dy << 16) | (dx & 0xffff);
              }
            ++na;
          } else {
            a_table = false;  // more items than registers: generic path below
          }
        }
        cpi += nlgrp;
        while (cpi >= npair) { cpi -= npair; if (++kw == p.ks) { kw = 0; ++ar; } }
      }
    }
    // Register-staged input copies (p.a_sts): every source 16 bytes is loaded from global memory ONCE (before the buffer is
    // even free) and written to its <= ks shifted destinations with st.shared.  The ks cp.async.ca copies per source cost 8
    // shared-memory / L1 wavefronts per warp instruction each and took 28-35 % of this kernel (profiles/r02r_wgrad.md: with
    // the input copies skipped conv_wgrad[5] drops from 819 k to 591 k cycles); a 16-lane st.shared run costs 2-3.
    // Warp items: phase 0 = (staged row, group pair) x source columns 0..15; phase 1 = (block of 4 staged rows, group pair) x
    // the ks - 1 halo columns 16.. of each row.
    constexpr int MAXS = 4;
    int s_so[MAXS], s_do[MAXS], s_dyx[MAXS];  // source element offset, destination of the kw = 0 copy, (dy + 64) << 16 | (dx + 64), or -1: lane unused
    int ns = 0;
    const int q0n = p.AR * npair, q1n = ((p.AR + 3) >> 2) * npair;
    // one decision for the whole CTA: every warp's share of the items fits the registers
    const bool a_sts = p.a_sts != 0 && (q0n + q1n + nlgrp - 1) / nlgrp <= MAXS;
    if (a_sts) {
      for (int q = lgrp; q < q0n + q1n; q += nlgrp) {
        int ar, xs, cpi;
        bool ok;
        if (q < q0n) {
          ar = q / npair; cpi = q - ar * npair; xs = xl; ok = true;
        } else {
          const int q1 = q - q0n, arb = q1 / npair;
          cpi = q1 - arb * npair; ar = arb * 4 + (xl >> 2); xs = WG_TW + (xl & 3);
          ok = ar < p.AR && (xl & 3) < p.ks - 1;
        }
        const int cg = 2 * cpi + cgp;
        ok = ok && cg < ncg_c;
        const int dy = ar + kh0 - p.pad, dx = xs - p.pad;
#pragma unroll
        for (int k = 0; k < MAXS; ++k)
          if (k == ns) {
            s_so[k] = (dy * p.w + dx) * p.C + (cg0 + cg) * 8;
            s_do[k] = cg * p.CGS_A + ar * (WG_TW * 16) + xs * 16;
            s_dyx[k] = ok ? (((dy + 64) << 16) | (dx + 64)) : -1;  // both biased by 64 (|dy|, |dx| < 64): a valid item is >= 0
          }
        ++ns;
      }
    }
    const uint16_t* xh = reinterpret_cast<const uint16_t*>(p.x);
    const uint16_t* zh = reinterpret_cast<const uint16_t*>(p.dz);
    uint32_t bi = 0, ph = 0;
    // tiles are consecutive: decode the first one, then step (no per-tile divisions; 32-bit pixel indices)
    int tx = t_begin % p.tiles_x, ty = (t_begin / p.tiles_x) % p.tiles_y, img_next = t_begin / (p.tiles_x * p.tiles_y);
    for (int t = t_begin; t < t_end; ++t) {
      const int y0 = ty * p.TR, x0 = tx * WG_TW, img = img_next;
      if (++tx == p.tiles_x) { tx = 0; if (++ty == p.tiles_y) { ty = 0; ++img_next; } }
      // register-staged copies go in rounds of two items per thread: round 0 is loaded while the MMAs may still be reading the
      // buffer (before the wait below), later rounds (wide channel slices only) after it
      uint4 s_hi[2], s_lo[2];
      const uint16_t* s_origin = xh + (ptrdiff_t)((img * p.h + y0) * p.w + x0) * p.C;
      auto s_load = [&](int base) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          s_hi[k] = make_uint4(0u, 0u, 0u, 0u);
          s_lo[k] = make_uint4(0u, 0u, 0u, 0u);
          const int dyx = k + base < ns ? (k + base == 0 ? s_dyx[0] : k + base == 1 ? s_dyx[1] : k + base == 2 ? s_dyx[2] : s_dyx[3]) : -1;
          if (dyx >= 0) {
            const int gy = y0 + (dyx >> 16) - 64, gx = x0 + (dyx & 0xffff) - 64;
            if ((unsigned)gy < (unsigned)p.h && (unsigned)gx < (unsigned)p.w) {
              const int so = k + base == 0 ? s_so[0] : k + base == 1 ? s_so[1] : k + base == 2 ? s_so[2] : s_so[3];
              const uint16_t* src = s_origin + so;
              s_hi[k] = __ldg(reinterpret_cast<const uint4*>(src));
              if (p.a_planes == 2) s_lo[k] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(src) + p.x_plane_bytes));
            }
          }
        }
      };
      if (a_sts && !(p.skip & 1)) s_load(0);
      wbar_wait(EMPTY + bi * 8, ph ^ 1);
      const uint32_t buf = buf0 + bi * p.buf_bytes;
      // ---- shifted input copies: staged row ar holds image row y0 + ar + kh0 - pad; copy kw holds source column
      //      x0 + xl + kw - pad; the kw copies of one source sector come from L1 (cp.async.ca)
      if (p.skip & 1) {
      } else if (a_sts) {
        const uint32_t kw_step = (uint32_t)(ncg_c * p.CGS_A) - 16u;  // next shift: next (kw, group) block, one pixel to the left
        for (int base = 0; base < ns; base += 2) {
          if (base > 0) s_load(base);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int dyx = k + base < ns ? (k + base == 0 ? s_dyx[0] : k + base == 1 ? s_dyx[1] : k + base == 2 ? s_dyx[2] : s_dyx[3]) : -1;
            if (dyx >= 0) {
              const int dof = k + base == 0 ? s_do[0] : k + base == 1 ? s_do[1] : k + base == 2 ? s_do[2] : s_do[3];
              uint32_t d = buf + (uint32_t)dof;
              const int xs = (dyx & 0xffff) - 64 + p.pad;  // source column within the staged row: 0 .. 15 + ks - 1
              for (int kw = 0; kw < p.ks; ++kw, d += kw_step) {
                const int xd = xs - kw;                    // its destination column in shift kw
                if ((unsigned)xd < (unsigned)WG_TW) {
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(s_hi[k].x), "r"(s_hi[k].y), "r"(s_hi[k].z), "r"(s_hi[k].w) : "memory");
                  if (p.a_planes == 2)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + p.a_plane_bytes), "r"(s_lo[k].x), "r"(s_lo[k].y), "r"(s_lo[k].z), "r"(s_lo[k].w) : "memory");
                }
              }
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // st.shared (generic proxy) -> MMA operand reads (async proxy)
      } else if (a_table) {
        const uint16_t* origin = xh + (ptrdiff_t)((img * p.h + y0) * p.w + x0) * p.C;
#pragma unroll
        for (int k = 0; k < MAXI; ++k) {
          if (k < na) {
            const int gy = y0 + (a_dyx[k] >> 16), gx = x0 + (int)(int16_t)(a_dyx[k] & 0xffff);
            const bool ok = (unsigned)gy < (unsigned)p.h && (unsigned)gx < (unsigned)p.w;
            const uint16_t* src = ok ? origin + a_so[k] : xh;
            const uint32_t d = buf + (uint32_t)a_do[k];
            wcp_async16_ca(d, src, ok ? 16u : 0u);
            if (p.a_planes == 2)
              wcp_async16_ca(d + p.a_plane_bytes, reinterpret_cast<const uint8_t*>(src) + p.x_plane_bytes, ok ? 16u : 0u);  // plane 1 of the dummy address is valid too
          }
        }
      } else {
        const uint32_t a_dst = buf + xl * 16;
        int cpi = lgrp, kw = 0, ar = 0;
        while (cpi >= npair) { cpi -= npair; if (++kw == p.ks) { kw = 0; ++ar; } }
        for (int q = lgrp; q < QA; q += nlgrp) {
          const int cg = 2 * cpi + cgp;
          if (cg < ncg_c) {
            const int gy = y0 + ar + kh0 - p.pad, gx = x0 + xl + kw - p.pad;
            const bool ok = (unsigned)gy < (unsigned)p.h && (unsigned)gx < (unsigned)p.w;
            const uint8_t* src = ok ? p.x + ((size_t)((img * p.h + gy) * p.w + gx) * p.C + (cg0 + cg) * 8) * 2 : p.x;
            const uint32_t d = a_dst + (uint32_t)(kw * ncg_c + cg) * p.CGS_A + (uint32_t)ar * (WG_TW * 16);
            wcp_async16_ca(d, src, ok ? 16u : 0u);
            if (p.a_planes == 2) wcp_async16_ca(d + p.a_plane_bytes, src + p.x_plane_bytes, ok ? 16u : 0u);
          }
          cpi += nlgrp;
          while (cpi >= npair) { cpi -= npair; if (++kw == p.ks) { kw = 0; ++ar; } }
        }
      }
      // ---- output-gradient tile
      if (!(p.skip & 2)) {
        const uint32_t b_dst = buf + p.a_planes * p.a_plane_bytes + bslot * 16;
        const int oy = y0 + br, ox = x0 + bxl;
        const bool pix_ok = oy < p.h && ox < p.w;
        const uint16_t* zpix = zh + (size_t)((img * p.h + (pix_ok ? oy : 0)) * p.w + (pix_ok ? ox : 0)) * p.dz_stride + n0;
#pragma unroll 2
        for (int cg = 2 * bgrp + cgp; cg < ncg_b; cg += 2 * nbgrp) {
          const bool ok = pix_ok && (n0 + cg * 8) < p.n_valid;
          const uint16_t* src = ok ? zpix + cg * 8 : zh;
          const uint32_t d = b_dst + (uint32_t)cg * p.CGS_B;
          wcp_async16(d, src, ok ? 16u : 0u);
          if (p.b_planes == 2)
            wcp_async16(d + p.b_plane_bytes, reinterpret_cast<const uint8_t*>(src) + p.dz_plane_bytes, ok ? 16u : 0u);
        }
      }
      wcp_async_arrive(FULL + bi * 8);
      if (++bi == (uint32_t)p.nbuf) { bi = 0; ph ^= 1; }
    }
    // ===================== epilogue (warps 0-3): TMEM -> partial dW =====================
    if (warp < 4) {
      wbar_wait(DONE, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int q = warp & 3;
      const int rows_total = p.ks * p.ks * p.C + 4;
      float* out = p.ws + (size_t)psplit * rows_total * p.N;
      const bool have_work = t_end > t_begin;
      for (int khl = 0; khl < nkh; ++khl)
        for (int mb = 0; mb < p.MB; ++mb) {
          const int kh = kh0 + khl;
          const int row = mb * 128 + q * 32 + lane;  // (kw, group, channel)
          const int g = row >> 3, ch = row & 7;
          int orow = -1;
          if (g < Gc) {
            const int kw = g / ncg_c, cg = g - kw * ncg_c;
            orow = (kh * p.ks + kw) * p.C + (cg0 + cg) * 8 + ch;
          } else if (g == Gc && ch < 4 && kh == 0 && cg0 == 0) {
            orow = p.ks * p.ks * p.C + ch;  // bias gradient row (+ 3 zero rows)
          }
          const uint32_t taddr = tmem_base + (khl * p.MB + mb) * (p.NC << BCAT) + ((uint32_t)(q * 32) << 16);
          for (int c0 = 0; c0 < nc; c0 += 16) {
            uint32_t v[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr + c0)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (BCAT) {  // + the x_hi * dz_lo partial sums, NC columns further
              uint32_t v2[16];
              asm volatile(
                  "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                  : "=r"(v2[0]), "=r"(v2[1]), "=r"(v2[2]), "=r"(v2[3]), "=r"(v2[4]), "=r"(v2[5]), "=r"(v2[6]), "=r"(v2[7]),
                    "=r"(v2[8]), "=r"(v2[9]), "=r"(v2[10]), "=r"(v2[11]), "=r"(v2[12]), "=r"(v2[13]), "=r"(v2[14]), "=r"(v2[15])
                  : "r"(taddr + p.NC + c0)
                  : "memory");
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int k = 0; k < 16; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) + __uint_as_float(v2[k]));
            }
            if (orow >= 0) {
              float4* dst = reinterpret_cast<float4*>(out + (size_t)orow * p.N + n0 + c0);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                dst[k] = have_work ? make_float4(__uint_as_float(v[4 * k]), __uint_as_float(v[4 * k + 1]),
                                                 __uint_as_float(v[4 * k + 2]), __uint_as_float(v[4 * k + 3]))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    p.dbg[0] = clock64() - dbg_c0;
    p.dbg[1] = t1 - dbg_t0;
  }
  if (warp == 18) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

__global__ void __launch_bounds__(256) wg_reduce_kernel(const float* __restrict__ ws, int64_t numel4, int splits,
                                                        float* __restrict__ out) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < numel4; e += (int64_t)gridDim.x * blockDim.x) {
    float4 s = reinterpret_cast<const float4*>(ws)[e];
    for (int k = 1; k < splits; ++k) {
      const float4 v = reinterpret_cast<const float4*>(ws)[(int64_t)k * numel4 + e];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(out)[e] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// Finish of all stages' weight gradients in one launch: fixed-order sum of the pixel-split partials (read
// coalesced, in packed order) and scatter into the reference layout (cout, cin_dst, k, k) + bias gradient.
// Replaces 7 x (wg_reduce_kernel + unpack_wgrad_kernel): 14 launches of 7-10 us each on 0.1 .. 30 MB.
// ---------------------------------------------------------------------------------------------
struct WgFinishItem {
  const float* ws;
  float* dw_ref;
  float* db_ref;
  int psplits, N, rows_total;      // partial buffers, columns per row, rows (k*k*cin_p + 4)
  int kk, cin, cin_p, cin_dst, cout, rr, cg;
  int n4, pad_elems;               // float4 per partial; elements of the zero-filled channel pad (cin..cin_dst)
};
struct WgFinishMulti {
  WgFinishItem t[NQ_MULTI_MAX];
  int blk_start[NQ_MULTI_MAX + 1];
  int n;
};

__global__ void __launch_bounds__(256) wg_finish_multi_kernel(const __grid_constant__ WgFinishMulti m) {
  int ti = 0;
  while (ti + 1 < m.n && (int)blockIdx.x >= m.blk_start[ti + 1]) ++ti;
  const WgFinishItem& t = m.t[ti];
  const int e = (blockIdx.x - m.blk_start[ti]) * 256 + threadIdx.x;
  if (e < t.pad_elems) {  // channels cin .. cin_dst of the (rotated-weight) layout carry no gradient
    const int per_co = (t.cin_dst - t.cin) * t.kk;
    const int co = e / per_co, r = e - co * per_co;
    t.dw_ref[((size_t)co * t.cin_dst + t.cin) * t.kk + r] = 0.f;
  }
  if (e >= t.n4) return;
  const float4* w4 = reinterpret_cast<const float4*>(t.ws);
  float4 s = w4[e];
  for (int k = 1; k < t.psplits; ++k) {
    const float4 v = w4[(size_t)k * t.n4 + e];
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  const int n4_per_row = t.N >> 2;
  const int row = e / n4_per_row, col0 = (e - row * n4_per_row) * 4;
  const int kdim = t.kk * t.cin_p;
  if (row > kdim) return;  // the 3 zero rows after the bias row
  const int tap = row / t.cin_p, ci = row - tap * t.cin_p;
  if (row < kdim && ci >= t.cin) return;
  const float v[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int np = col0 + j;
    const int grp = np / t.cg, c = np - grp * t.cg;
    const int co = c * t.rr + grp;
    if (grp >= t.rr || co >= t.cout) continue;
    if (row == kdim) {
      if (t.db_ref) t.db_ref[co] = v[j];
    } else if (t.dw_ref) {
      t.dw_ref[((size_t)co * t.cin_dst + ci) * t.kk + tap] = v[j];
    }
  }
}

int check_conv_desc(const nq_conv_desc* d);

static int fill_wg_plan(const nq_conv_desc* d, int a_planes, int b_planes, nq_tc_wgrad_plan* pl) {
  int st = check_conv_desc(d);
  if (st) return st;
  if (!pl || a_planes < 1 || a_planes > 2 || b_planes < 1 || b_planes > 2) return NQ_ERR_BAD_ARG;
  const int C = d->cin_p;
  const int N = (d->rh * d->rw * d->cg + 15) / 16 * 16;  // padded columns are zero-filled by the loader
  if (C % 8 || d->ksize > 7) return NQ_ERR_BAD_SHAPE;
  pl->C = C; pl->N = N; pl->a_planes = a_planes; pl->b_planes = b_planes;
  pl->ncg = C / 8;
  pl->G = d->ksize * pl->ncg;
  // Choose (input-channel slices msplit, kernel rows per CTA nkh, columns per CTA NC) with
  // nkh * MB * NC <= 512 TMEM columns, MB = 128-row blocks of one slice's (kw, group) pairs (+ the bias group).
  // Cost model fitted to the B200 measurements in profiles/: the kernel is bound by the larger of
  //   MMA time     ~ max(NC/2, 40 + NC/8) cycles per MMA (the A read from shared memory binds narrow N), and
  //   L2->SM time  ~ 1.8 x algorithmic bytes / 5.5 TB/s, where dZ is read once per (slice, kernel-row group) and
  //                  the input once per (column slice, kernel-row group) with (TR + nkh - 1) / TR row overlap.
  const double P = (double)d->n * d->h * d->w;
  const int passes = 1 + (a_planes == 2) + (b_planes == 2);
  double best_cost = 1e300;
  int best_nkh = 0, best_nc = 0, best_ms = 1;
  for (int ms = 1; ms <= pl->ncg && ms <= 8; ++ms) {
    const int ncg_c = (pl->ncg + ms - 1) / ms;
    if ((ms - 1) * ncg_c >= pl->ncg) continue;  // empty last slice
    const int mb = (d->ksize * ncg_c + 1 + 15) / 16;
    if (mb > 4) continue;
    for (int nkh = 1; nkh <= d->ksize; ++nkh) {
      int nc = (512 / (nkh * mb)) / 16 * 16;
      if (nc > 256) nc = 256;
      if (nc > N) nc = N;
      if (nc < 16) continue;
      const int nsp = (N + nc - 1) / nc;
      const int khg = (d->ksize + nkh - 1) / nkh;
      const double cyc = nc * 0.5 > 40.0 + nc / 8.0 ? nc * 0.5 : 40.0 + nc / 8.0;
      const double mma = P / 16.0 * d->ksize * mb * ms * nsp * passes * cyc / (148.0 * 1.9e9);
      const double bytes = P * 2.0 * (b_planes * (double)N * ms * khg + a_planes * (double)C * nsp * khg * (4.0 + nkh - 1) / 4.0);
      const double traffic = 1.8 * bytes / 5.5e12;
      const double ctas = (double)ms * nsp * khg;
      double cost = (mma > traffic ? mma : traffic) + 2e-6 * ctas;  // mild preference for fewer CTA types
      if (cost < best_cost) { best_cost = cost; best_nkh = nkh; best_nc = nc; best_ms = ms; }
    }
  }
  if (!best_nkh) return NQ_ERR_UNSUPPORTED;  // C * ks too large for one TMEM pass even with 8 slices
  if (const char* e = getenv("NQ_WG_CFG")) {  // tuning override "msplit,nkh"
    int ms = 0, nkh = 0;
    if (sscanf(e, "%d,%d", &ms, &nkh) == 2 && ms >= 1 && ms <= pl->ncg && nkh >= 1 && nkh <= d->ksize) {
      const int ncg_c = (pl->ncg + ms - 1) / ms;
      const int mb = (d->ksize * ncg_c + 1 + 15) / 16;
      int c = mb <= 4 ? (512 / (nkh * mb)) / 16 * 16 : 0;
      if (c > 256) c = 256;
      if (c > N) c = N;
      if (c >= 16 && (ms - 1) * ncg_c < pl->ncg) { best_ms = ms; best_nkh = nkh; best_nc = c; }
    }
  }
  pl->msplit = best_ms;
  pl->ncg_c = (pl->ncg + best_ms - 1) / best_ms;
  pl->MB = (d->ksize * pl->ncg_c + 1 + 15) / 16;
  // shared-memory fit: tile rows as many as leave room for >= 2 pipeline buffers
  auto fit = [&](int nkh, int nc) -> bool {
    pl->nkh = nkh;
    pl->khg = (d->ksize + nkh - 1) / nkh;
    pl->NC = nc;
    pl->nsplits = (N + nc - 1) / nc;
    int nbuf = 0, tr = 4;
    int min_buf = 2;
    if (const char* e = getenv("NQ_WG_MINBUF")) min_buf = atoi(e) < 2 ? 2 : atoi(e);  // tuning override
    for (; tr >= 1; tr >>= 1) {
      pl->AR = tr + nkh - 1;
      pl->CGS_A = pl->AR * WG_TW * 16 + 64;  // +64: lane pairs (same pixel, adjacent groups) store conflict free
      pl->CGS_B = tr * WG_TW * 16 + 64;
      pl->a_plane_bytes = pl->MB * 16 * pl->CGS_A;
      pl->b_plane_bytes = (nc / 8) * pl->CGS_B;
      pl->buf_bytes = a_planes * pl->a_plane_bytes + b_planes * pl->b_plane_bytes;
      nbuf = (227 * 1024 - 128) / pl->buf_bytes;
      if (nbuf >= min_buf || (tr == 1 && nbuf >= 2)) break;
    }
    if (nbuf < 2) return false;
    pl->TR = tr;
    pl->nbuf = nbuf > 4 ? 4 : nbuf;
    pl->smem_bytes = 128 + pl->nbuf * pl->buf_bytes;
    return true;
  };
  if (!fit(best_nkh, best_nc)) {
    // fall back to one kernel row per CTA with the widest column slice
    int c = (512 / pl->MB) / 16 * 16;
    if (c > 256) c = 256;
    if (c > N) c = N;
    if (!fit(1, c)) return NQ_ERR_UNSUPPORTED;
  }
  // dZ planes as one operand (see WgParams::bcat): single 128-row block, one column slice, accumulators still fit
  pl->bcat = (a_planes == 2 && b_planes == 2 && pl->MB == 1 && pl->nsplits == 1 && pl->NC <= 128 && 2 * pl->nkh * pl->NC <= 512) ? 1 : 0;  // one MMA takes N <= 256
  if (const char* e = getenv("NQ_WG_BCAT")) { if (atoi(e) == 0) pl->bcat = 0; }  // tuning override
  pl->tiles_x = (d->w + WG_TW - 1) / WG_TW;
  pl->tiles_y = (d->h + pl->TR - 1) / pl->TR;
  pl->tiles_total = pl->tiles_x * pl->tiles_y * d->n;
  int ps = sm_count() / (pl->khg * pl->nsplits * pl->msplit);
  if (ps < 1) ps = 1;
  if (ps > pl->tiles_total) ps = pl->tiles_total;
  pl->tiles_per_split = (pl->tiles_total + ps - 1) / ps;
  pl->psplits = (pl->tiles_total + pl->tiles_per_split - 1) / pl->tiles_per_split;
  pl->workspace_floats = (int64_t)pl->psplits * (d->ksize * d->ksize * C + 4) * N;
  return NQ_OK;
}

}  // namespace nq

using namespace nq;

extern "C" int nq_tc_plan_wgrad(const nq_conv_desc* d, int a_planes, int b_planes, nq_tc_wgrad_plan* plan) {
  return fill_wg_plan(d, a_planes, b_planes, plan);
}

extern "C" int nq_tc_conv_wgrad(const nq_conv_desc* d, const nq_tc_wgrad_plan* pl, const void* x_split, const void* dz_split,
                                float* dwk, float* workspace, int64_t workspace_floats, void* stream) {
  const uint8_t* x = reinterpret_cast<const uint8_t*>(x_split);
  const uint8_t* dz = reinterpret_cast<const uint8_t*>(dz_split);
  int st = check_conv_desc(d);
  if (st) return st;
  if (!pl || !x || !dz || !workspace) return NQ_ERR_BAD_ARG;  // dwk NULL: the caller finishes with nq_tc_wgrad_finish_multi
  if (workspace_floats < pl->workspace_floats) return NQ_ERR_WORKSPACE;
  WgParams p{};
  p.x = x; p.dz = dz; p.ws = workspace;
  p.n = d->n; p.h = d->h; p.w = d->w; p.C = pl->C; p.N = pl->N; p.ks = d->ksize; p.pad = d->ksize / 2;
  p.ncg = pl->ncg; p.G = pl->G; p.MB = pl->MB; p.NC = pl->NC; p.nsplits = pl->nsplits; p.TR = pl->TR;
  p.nkh = pl->nkh; p.khg = pl->khg; p.AR = pl->AR; p.msplit = pl->msplit; p.ncg_c = pl->ncg_c; p.bcat = pl->bcat;
  if (p.bcat && (pl->MB != 1 || pl->nsplits != 1 || pl->a_planes != 2 || pl->b_planes != 2 || 2 * pl->nkh * pl->NC > 512 || pl->NC > 128 ||
                 pl->b_plane_bytes != (pl->NC / 8) * pl->CGS_B))
    return NQ_ERR_BAD_ARG;
  p.tiles_x = pl->tiles_x; p.tiles_y = pl->tiles_y; p.tiles_total = pl->tiles_total; p.psplits = pl->psplits;
  p.tiles_per_split = pl->tiles_per_split; p.a_planes = pl->a_planes; p.b_planes = pl->b_planes;
  p.CGS_A = pl->CGS_A; p.CGS_B = pl->CGS_B; p.a_plane_bytes = pl->a_plane_bytes; p.b_plane_bytes = pl->b_plane_bytes;
  p.buf_bytes = pl->buf_bytes; p.nbuf = pl->nbuf;
  p.dz_stride = (d->rh * d->rw * d->cg + 7) / 8 * 8;  // channels per pixel as stored (the head's 4 are stored as 8)
  p.n_valid = p.dz_stride;
  p.x_plane_bytes = (size_t)d->n * d->h * d->w * pl->C * 2;
  p.dz_plane_bytes = (size_t)d->n * d->h * d->w * p.dz_stride * 2;
  cudaStream_t s = as_stream(stream);
  const int passes = (p.a_planes == 2 ? 1 : 0) | (p.b_planes == 2 ? 2 : 0);
  if (p.bcat && passes != 3) return NQ_ERR_BAD_ARG;
  void (*kern)(const WgParams) = p.bcat      ? wgrad_tc_kernel<1, 3>
                                 : passes == 3 ? wgrad_tc_kernel<0, 3>
                                 : passes == 2 ? wgrad_tc_kernel<0, 2>
                                 : passes == 1 ? wgrad_tc_kernel<0, 1>
                                               : wgrad_tc_kernel<0, 0>;
  NQ_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int grid = pl->psplits * pl->msplit * pl->nsplits * pl->khg;
  static const int sts_env = getenv("NQ_WG_STS") ? atoi(getenv("NQ_WG_STS")) : 1;  // tuning override (0: cp.async copies)
  p.a_sts = sts_env != 0 && p.ks > 1 && p.ks <= 5;
  static const int skip_flags = getenv("NQ_WG_SKIP") ? atoi(getenv("NQ_WG_SKIP")) : 0;
  p.skip = skip_flags & 3;  // bit 2 (no MMAs) leaves the commits out as well and never finishes: not accepted
  static const bool dbg_on = getenv("NQ_TC_DBG") != nullptr;
  static long long* dbg_buf = nullptr;
  if (dbg_on) {  // debugging aid: synchronises after the launch and prints the SM cycles / clock the kernel saw
    if (!dbg_buf) NQ_CUDA_CHECK(cudaMalloc(&dbg_buf, 2 * sizeof(long long)));
    p.dbg = dbg_buf;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(WG_THREADS);
  cfg.dynamicSmemBytes = (size_t)pl->smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  NQ_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
  if (dbg_on) {
    long long h[2] = {0, 0};
    NQ_CUDA_CHECK(cudaStreamSynchronize(s));
    NQ_CUDA_CHECK(cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[NQ_TC_DBG] wgrad hw=%dx%d C=%d N=%d ks=%d msplit=%d nkh=%d NC=%d TR=%d nbuf=%d: %lld cycles, %lld ns, SM clock %.0f MHz\n", p.h, p.w,
            p.C, p.N, p.ks, p.msplit, p.nkh, p.NC, p.TR, p.nbuf, h[0], h[1], h[1] > 0 ? (double)h[0] / (double)h[1] * 1e3 : 0.0);
  }
  NQ_LAUNCH_CHECK();
  if (!dwk) return NQ_OK;
  const int64_t n4 = (int64_t)(d->ksize * d->ksize * pl->C + 4) * pl->N / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  wg_reduce_kernel<<<(unsigned)blocks, 256, 0, s>>>(workspace, n4, pl->psplits, dwk);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

extern "C" int nq_tc_wgrad_finish_multi(const nq_wgrad_finish_task* tasks, int n_tasks, void* stream) {
  if (!tasks || n_tasks <= 0) return NQ_ERR_BAD_ARG;
  for (int i0 = 0; i0 < n_tasks; i0 += NQ_MULTI_MAX) {
    WgFinishMulti m{};
    m.n = n_tasks - i0 < NQ_MULTI_MAX ? n_tasks - i0 : NQ_MULTI_MAX;
    int blocks = 0;
    for (int i = 0; i < m.n; ++i) {
      const nq_wgrad_finish_task& t = tasks[i0 + i];
      if (!t.d || !t.workspace || t.psplits <= 0 || t.n_cols <= 0 || t.n_cols % 4) return NQ_ERR_BAD_ARG;
      const int st = check_conv_desc(t.d);
      if (st) return st;
      const nq_conv_desc* d = t.d;
      if (t.cin_dst < d->cin || d->rh * d->rw * d->cg > t.n_cols) return NQ_ERR_BAD_ARG;
      WgFinishItem& w = m.t[i];
      w.ws = t.workspace; w.dw_ref = t.dw_ref; w.db_ref = t.db_ref;
      w.psplits = t.psplits; w.N = t.n_cols;
      w.kk = d->ksize * d->ksize; w.cin = d->cin; w.cin_p = d->cin_p; w.cin_dst = t.cin_dst; w.cout = d->cout;
      w.rr = d->rh * d->rw; w.cg = d->cg;
      w.rows_total = w.kk * d->cin_p + 4;
      const long long n4 = (long long)w.rows_total * t.n_cols / 4;
      const long long pad = t.dw_ref ? (long long)d->cout * (t.cin_dst - d->cin) * w.kk : 0;
      if (n4 >= (1LL << 31) || pad >= (1LL << 31)) return NQ_ERR_BAD_SHAPE;
      w.n4 = (int)n4; w.pad_elems = (int)pad;
      m.blk_start[i] = blocks;
      const long long work = n4 > pad ? n4 : pad;
      blocks += (int)((work + 255) / 256);
    }
    m.blk_start[m.n] = blocks;
    wg_finish_multi_kernel<<<blocks, 256, 0, as_stream(stream)>>>(m);
    NQ_LAUNCH_CHECK();
  }
  return NQ_OK;
}
