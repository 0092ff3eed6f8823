// tcgen05 / TMEM implicit-GEMM convolution for sm_100a: forward (bias + per-channel scale + up-shuffle +
// exact-erf GELU fused in the epilogue) and data gradient (GELU' + un-shuffle fused), one kernel template.
// Replaces F.conv2d (quant_layer.py:80) + nn.PixelShuffle + nn.GELU (quant_block.py:31-35) and the
// dgrad half of their autograd.
//
// Formulation ("halo tile" implicit GEMM, no im2col buffer, no per-tap re-load):
//   out[m][n] = sum_{tap, c} In[pix(m) + tap][c] * B[tap][c][n]          m = pixel, n = output channel
//   * a CTA owns a 16x8-pixel output tile (GEMM-M = 128) and one N tile (<= 256 columns)
//   * the (16+k-1) x (8+k-1) input halo of the tile is staged ONCE in shared memory as bf16, in the
//     canonical no-swizzle K-major core-matrix order  [channel-group of 8][halo pixel][8 channels]
//     (16 B per pixel and group).  For a fixed tap the 128 A rows are then a strided view of that
//     buffer -- 8 consecutive x are 8 consecutive 16-byte rows (one core matrix), tile rows are
//     PW*16 bytes apart (the descriptor's SBO), the next 8 channels are CGS bytes away (its LBO) --
//     so every tap is just a different descriptor start address: k*k-fold reuse out of shared memory.
//   * B (weights) streams through a ring of stages with cp.async.bulk (TMA bulk copy) + mbarrier
//     complete_tx; it is pre-packed by nq_tc_pack_weight in exactly the order the MMA consumes it.
//   * accumulators live in TMEM (2 x 256 fp32 columns, double buffered so the epilogue of tile i
//     overlaps the MMAs of tile i+1); one elected thread issues tcgen05.mma.kind::f16.
//   * fp32 accuracy from bf16 tensor cores: activations are split x = hi + lo (two bf16 planes, 16
//     mantissa bits), weights either are exact in one bf16 plane (integer codes - zero_point, |v| <= 255,
//     per-channel scale applied in the epilogue) or are split as well; the product is accumulated as
//     hi*hi + lo*hi + hi*lo in fp32.
//   * activations and gradients LIVE in HBM in that split form ("split-bf16": plane 0 = hi, plane 1 = lo,
//     each NHWC bf16; same bytes as fp32): the producing epilogue converts once, every consumer (forward
//     and wgrad of the next stage, dgrad and wgrad of this one) copies 16-byte chunks with cp.async --
//     no conversion, no register staging, the copies of a whole halo tile are in flight at once and
//     complete on an mbarrier (cp.async.mbarrier.arrive.noinc).
//
// Warp roles (640 threads): warps 0..n_epi-1 epilogue (TMEM -> registers -> global; n_epi = 8, or 12 for short-K
// stages whose epilogue is the longest role), warps n_epi..15 activation loaders
// (fp32 -> bf16 hi/lo), warp 16 weight-stage producer, warp 17 MMA issuer, warp 18 TMEM allocator.
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>

#include "nq_common.cuh"

namespace nq {

constexpr int TC_THREADS = 640;
constexpr int TC_WORK_WARPS = 16;  // warps 0-15: n_epi epilogue warps (8 or 12), the rest load activations
constexpr int TILE_H = 16, TILE_W = 8;  // 128 output pixels per tile (GEMM M)
constexpr int TC_MAX_BSTAGES = 16;
constexpr int TC_MAX_RING = 8;        // activation buffers / TMEM accumulator slots
constexpr int TC_HDR_BYTES = 1024;    // barriers (4 * TC_MAX_RING + 2 * TC_MAX_BSTAGES) + TMEM pointer
constexpr int EPI_ROW = 20;            // floats per staged epilogue row (16 + 4 pad: conflict-free 16-byte accesses)
constexpr int EPI_STAGE_BYTES = 12 * 32 * EPI_ROW * 4;  // up to 12 epilogue warps

// Division by a kernel-invariant divisor as one multiply-high: m = floor(2^32 / d) + 1 gives the exact quotient
// for every x with x * d < 2^32 (all uses below: tile and pixel indices, d <= a few hundred).
struct FastDiv {
  uint32_t m, d;
};
__host__ __device__ __forceinline__ FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = (uint32_t)d;
  f.m = d > 1 ? (uint32_t)((1ull << 32) / (uint32_t)d) + 1u : 0u;
  return f;
}
__device__ __forceinline__ int fdiv(int x, FastDiv f) { return f.d == 1u ? x : (int)__umulhi((uint32_t)x, f.m); }

struct TcParams {
  const uint8_t* in;     // split-bf16 input: plane 0 (hi) then plane 1 (lo), each (n, h, w, in_stride) bf16
  size_t in_plane_bytes;
  const uint8_t* wpk;    // packed bf16 weight stages
  const float* scale;    // [N] per-column scale (fwd) or null
  const float* bias;     // [N] (fwd) or null
  const float* zprev;    // dgrad: (n, h, w, n_store) fp32 pre-activation of the previous stage, or null
  float* out_z;          // fwd: fp32 pre-activation (may be null)
  uint8_t* out_y;        // split-bf16 output (fwd: activated output, may be null; dgrad: dz_prev)
  size_t out_plane_bytes;
  int n, h, w, C;        // input grid and GEMM-K channels (C % 16 == 0)
  int ks, pad;
  int N, NT;             // GEMM N (multiple of 16) and its tile
  int KC, SBC;           // channels per A unit / per B stage (multiples of 16, SBC | KC, SBC | C)
  int a_planes, b_planes;
  int epi;               // 0 forward, 1 dgrad, 2 head (forward + OutImg + loss + dL/dz)
  // head epilogue (epi == 2)
  const float* head_target; float* head_img; float* head_loss; uint8_t* head_dz;
  float head_p, head_inv_mean; int head_out_bias;
  int rh, rw, cg, act;   // fwd: up-shuffle of this stage's output; dgrad: previous stage's (un-shuffle)
  int tiles_x, tiles_y, tiles_n, total_tiles;
  int PW, PH, CGS;       // halo width/height (pixels), channel-group stride (bytes)
  int a_plane_bytes, a_buf_bytes, b_stage_bytes, n_bstages;  // n_bstages ring slots of gst weight stages each
  int gst, b_slot_bytes;  // weight stages per ring slot (one bulk copy, one barrier round trip), bytes per slot
  int in_stride, c_valid;  // channels per pixel stored in `in` (<= C) : channels >= c_valid are read as zero
  int n_store;           // dgrad: columns actually stored / row stride of zprev and out (<= N, N padded to 16)
  int epi_stage_off;     // byte offset of the epilogue staging tiles in shared memory
  int cs;                // CTAs per cluster sharing every weight stage by TMA multicast (1, 2 or 4)
  int tiles_m, tiles_m_pad, total_groups;  // pixel tiles per N tile, padded to a multiple of cs; tile groups
  int n_abuf, n_acc;     // activation buffers in shared memory, accumulator slots in TMEM (rings, 2 .. TC_MAX_RING)
  int acc_stride, sub_stride;  // TMEM columns per accumulator slot / between the pixel tiles of a slot
  FastDiv fd_tiles_x, fd_tiles_y, fd_rh, fd_rw, fd_PW, fd_npix, fd_cg;
  int ksplit, cb_per_split, ncb;  // split-K over activation units (channel blocks): units per split, units in total
  float* part;           // epi 3: fp32 partial sums (ksplit, n, h, w, n_store); the finish kernel applies the epilogue
  size_t part_stride;    // floats per partial
  int nsb_last;          // weight stages of the last (possibly partial) activation unit
  int out_h, out_w, quad_stride;  // output grid (fwd: h*rh, w*rw; dgrad: h/rh, w/rw) and dgrad's floats per output pixel
  int resident;          // 1: all weight stages of a tile fit the ring and stay there: loaded once per CTA, never released
  int n_epi;             // epilogue warps (8, or 12 for short-K stages whose epilogue binds); loaders = 16 - n_epi warps
  int bcat;              // weight stage stores the planes side by side per k-group ([k-group][plane][n][8]): A_hi x [B_hi | B_lo]
                         // is ONE MMA of 2 * nt columns (hi*hi in columns [0, nt), hi*lo in [nt, 2 nt)) + A_lo x B_hi
  int mt;                // 16x8 pixel tiles (side by side in x) per CTA step: they share every weight stage (NT <= 256 / mt)
  int n_prod;            // lanes of the producer warp issuing weight-stage copies (stages round robin)
  int skip;              // NQ_TC_SKIP (debug): bit 0 epilogue body, bit 2 activation copies
  long long* dbg;        // NQ_TC_DBG: {SM cycles, nanoseconds} of CTA 0 (the SM clock this launch really ran at), or null
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
// elect.sync: ptxas knows that exactly one lane runs the guarded region and keeps its values on the
// uniform datapath (UTCHMMA takes uniform-register operands; a plain `lane == 0` guard makes it emit a
// broadcast loop of R2UR moves around every MMA)
__device__ __forceinline__ bool elect_one() {
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
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16-byte asynchronous copy global -> shared; src_bytes = 0 zero-fills (halo outside the image)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// arrive on `bar` once all of this thread's prior cp.async have landed (does not bump the pending count)
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
// same, descriptors passed as 32-bit halves (the high halves are loop invariant)
__device__ __forceinline__ void umma_bf16_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---- CTA pair (cta_group::2): one MMA of M = 256 spans the two CTAs of a cluster; each holds its 128 A rows and HALF of
// the B rows at the same shared-memory offsets; the leader CTA issues, tcgen05.commit signals both CTAs' barriers.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (release at CTA scope), as cutlass::arch::ClusterBarrier::arrive(cta_id): what this arrival publishes was
  // written by the async proxy (TMA) or already ordered by fence.proxy.async; a cluster-scope release costs a full fence per stage
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {  // a remote CTA arrives on this barrier
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
#define NQ_UMMA2(NAME, QUAL)                                                                                              \
  __device__ __forceinline__ void NAME(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,          \
                                       uint32_t idesc, uint32_t accum) {                                                    \
    asm volatile(                                                                                                           \
        "{\n\t"                                                                                                             \
        ".reg .pred p;\n\t"                                                                                                 \
        ".reg .b64 da, db;\n\t"                                                                                             \
        "setp.ne.b32 p, %6, 0;\n\t"                                                                                         \
        "mov.b64 da, {%1, %2};\n\t"                                                                                         \
        "mov.b64 db, {%3, %4};\n\t"                                                                                         \
        "tcgen05.mma.cta_group::2.kind::f16" QUAL " [%0], da, db, %5, p;\n\t"                                              \
        "}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)                               \
        : "memory");                                                                                                        \
  }
NQ_UMMA2(umma2_w, "")
// the A operand stays in the collector for the next MMA (same A, other B): its 4 KB are read from shared memory once
NQ_UMMA2(umma2_w_keep, ".collector::a::fill")
NQ_UMMA2(umma2_w_reuse, ".collector::a::lastuse")
#undef NQ_UMMA2
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start >> 4, [16,30) leading-dim byte offset >> 4 (between the two 16-byte K chunks of one MMA),
// [32,46) stride-dim byte offset >> 4 (between 8-row groups), [46,48) version = 1, [61,64) layout 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = n.
__device__ __forceinline__ uint32_t make_idesc(int n, int m = 128) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits in two bf16 planes
__device__ __forceinline__ void split8(const float4& a, const float4& b, uint4& hi, uint4& lo) {
  const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  float r[8];
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * i]), h1 = __float2bfloat16_rn(x[2 * i + 1]);
    r[2 * i] = x[2 * i] - __bfloat162float(h0);
    r[2 * i + 1] = x[2 * i + 1] - __bfloat162float(h1);
    h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[i] = pack_bf16x2(r[2 * i], r[2 * i + 1]);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// four consecutive channels of element index `o` -> hi and lo bf16 planes (8 bytes each)
__device__ __forceinline__ void store_split4(uint8_t* base, size_t plane_bytes, size_t o, const float4& r) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(r.x), h1 = __float2bfloat16_rn(r.y), h2 = __float2bfloat16_rn(r.z),
                      h3 = __float2bfloat16_rn(r.w);
  uint2 hi, lo;
  hi.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
  hi.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
  lo.x = pack_bf16x2(r.x - __bfloat162float(h0), r.y - __bfloat162float(h1));
  lo.y = pack_bf16x2(r.z - __bfloat162float(h2), r.w - __bfloat162float(h3));
  *reinterpret_cast<uint2*>(base + o * 2) = hi;
  *reinterpret_cast<uint2*>(base + plane_bytes + o * 2) = lo;
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
struct TileCoord {
  int img, y0, x0, n0, nt;  // image, tile origin (pixels), first column, columns in this N tile
  bool real;                // false: padding slot of a cluster group (runs the pipeline, touches no pixels)
  int cb0, cb1, ksp;        // activation units [cb0, cb1) of this slot and its split index (split-K)
};
// Tile slots are ordered pixel-tile fastest within an N tile, padded so that the cs CTAs of a cluster
// always work on the same N tile (they share its weight stages by multicast).
__device__ __forceinline__ TileCoord tile_coord(const TcParams& p, int group, int rank) {
  TileCoord c;
  const int slot = group * p.cs + rank;
  int tn, tm;
  if (p.ksplit == 1) {
    tn = p.tiles_n == 1 ? 0 : slot / p.tiles_m_pad;
    tm = slot - tn * p.tiles_m_pad;
    c.cb0 = 0; c.cb1 = p.ncb; c.ksp = 0;
  } else {  // slots: pixel tile fastest, then K split, then N tile
    const int tnk = slot / p.tiles_m_pad;
    tm = slot - tnk * p.tiles_m_pad;
    tn = tnk / p.ksplit;
    c.ksp = tnk - tn * p.ksplit;
    c.cb0 = c.ksp * p.cb_per_split;
    c.cb1 = min(p.ncb, c.cb0 + p.cb_per_split);
  }
  c.real = tm < p.tiles_m;
  if (!c.real) tm = 0;
  const int tmx = fdiv(tm, p.fd_tiles_x);
  const int tx = tm - tmx * p.tiles_x;
  c.img = fdiv(tmx, p.fd_tiles_y);
  const int ty = tmx - c.img * p.tiles_y;
  c.y0 = ty * TILE_H;
  c.x0 = tx * TILE_W * p.mt;
  c.n0 = tn * p.NT;
  c.nt = min(p.NT, p.N - c.n0);
  return c;
}

// MMAs of one weight stage (k16 K steps of 16 channels), fully unrolled.  PASSES bit 0: activation lo plane, bit 1: weight lo
// plane; BCAT: both weight planes as one operand of 2 nt columns (idesc2).  Only the elected lane issues; the block is one
// reconvergence scope, the descriptor low words are independent adds.
// tcgen05.commit of a weight stage's B_EMPTY barrier: pair / single CTA / every CTA of a multicast cluster
template <int CG>
__device__ __forceinline__ void commit_stage(uint32_t bar, int cs, uint16_t mc_mask) {
  if (CG == 2) umma2_commit_mc(bar, 3);
  else if (cs == 1) umma_commit(bar);
  else umma_commit_mc(bar, mc_mask);
}

// MT / BCAT / RES are compile-time copies of TcParams::mt / bcat / resident: the common (1, 0, 0) variant carries
// none of their code.
// CG = 2: CTA-pair MMAs (cta_group::2, M = 256).  The two CTAs of a cluster work on neighbouring pixel tiles of the same N
// tile; each stages its own halo tile and HALF of every weight stage (N / 2 rows of B); the leader (rank 0) issues every MMA
// for both.  What the leader must know about its peer -- activation tile landed, weight half landed, accumulator drained --
// reaches it as one extra arrival on ITS OWN barriers (remote mbarrier.arrive from the peer's relay warps 17 / 19 and
// epilogue warps); what the peer must know -- operands consumed, accumulator ready -- is the multicast tcgen05.commit.
// Every weight byte is read from L2 once per pair, written to shared memory once per pair, and each tensor core fetches
// half of B from its neighbour: operand reads per MMA drop from 128 x 32 + N x 32 bytes to 128 x 32 + N x 16 per SM
// (tools/umma_bench2.cu: the operand fetch runs at 128 B/clk/SM and binds below N = 128).  Measured on HNeRV-3M
// (profiles/r02h_conv_analysis.md): conv_fwd[5] 537k -> 476k SM cycles, conv_dgrad[5] 914k -> 843k, stage 4 likewise;
// the deep split-K stages (few tiles per SM) keep the multicast form, where the relay hop is not exposed.
template <int MT, int BCAT, int RES, int CG>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  // [0, TC_HDR_BYTES): barriers + tmem pointer; then A buffers, then B stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar0 = smem_u32(bars);
  // barrier indices
  const uint32_t A_FULL = bar0, A_EMPTY = bar0 + TC_MAX_RING * 8, T_FULL = bar0 + 2 * TC_MAX_RING * 8,
                 T_EMPTY = bar0 + 3 * TC_MAX_RING * 8;
  const uint32_t B_FULL = bar0 + 4 * TC_MAX_RING * 8, B_EMPTY = B_FULL + TC_MAX_BSTAGES * 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (4 * TC_MAX_RING + 2 * TC_MAX_BSTAGES) * 8);
  const uint32_t a_base = smem_u32(smem + TC_HDR_BYTES);
  const uint32_t b_base = a_base + p.n_abuf * p.a_buf_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = p.cs > 1 ? (int)cluster_rank() : 0;
  const int cluster_id = blockIdx.x / p.cs, n_clusters = gridDim.x / p.cs;
  const uint16_t mc_mask = (uint16_t)((1u << p.cs) - 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_MAX_RING; ++i) {
      // one deferred arrival per loader thread (+ the peer's relay on the pair leader)
      mbar_init(A_FULL + i * 8, (TC_WORK_WARPS - p.n_epi) * 32 + (CG == 2 && rank == 0 ? 1 : 0));
      mbar_init(A_EMPTY + i * 8, 1);
      mbar_init(T_FULL + i * 8, 1);
      mbar_init(T_EMPTY + i * 8, CG == 2 ? 2 * p.n_epi : p.n_epi);  // pair: both CTAs' epilogue warps arrive on the leader's
    }
    for (int i = 0; i < p.n_bstages; ++i) {
      mbar_init(B_FULL + i * 8, CG == 2 && rank == 0 ? 2 : 1);  // pair leader: own producer + the peer's relay
      mbar_init(B_EMPTY + i * 8, CG == 2 ? 1 : p.cs);  // multicast: every CTA of the cluster must have consumed the slot
    }
    fence_barrier_init();
  }
  if (warp == 18) {
    if (CG == 2) {  // both CTAs of the pair, the same warp
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.cs > 1) cluster_sync_all();  // peers' barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  pdl_wait();     // everything above ran under the tail of the kernel before this one (nq_common.cuh)
  pdl_trigger();
  const uint32_t tmem_base = *tmem_slot;
  long long dbg_c0 = 0, dbg_t0 = 0;
  if (p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    dbg_c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
  }

  const int taps = p.ks * p.ks;
  const int ncb = (p.C + p.KC - 1) / p.KC;

  if (warp == 16) {
    // ===================== weight-stage producer (TMA bulk copies) =====================
    // n_prod lanes take the ring slots round robin.  One thread sustains only about one bulk copy per 650-1200 SM cycles
    // however small the copy (tools/tma_bulk_bench.cu: 15 KB copies land at 22 B/clk/SM from one thread, 49 from four
    // lanes, 3 KB copies at 2.6 / 8.8), which would bind plans with small slots; on HNeRV-3M one lane or four measure the
    // same (profiles/r02h_conv_analysis.md: the stage time is set by the issuing thread, not by the stream).
    if (lane < p.n_prod) {
      int turn = 0;
      const int nsb_full = p.KC / p.SBC;
      const int stages_per_ntile = (p.C / p.SBC) * taps;
      // pair plans with side-by-side planes (CG == 2 && BCAT) store 3 nt SBC bytes per stage and CTA, see pack
      const size_t ntile_stride = (size_t)stages_per_ntile * p.NT * p.SBC * (CG == 2 && BCAT ? 6 : 2 * p.b_planes);
      uint32_t bs = 0, bph = 0;
      for (int t = cluster_id; t < p.total_groups; t += n_clusters) {
        const TileCoord tc = tile_coord(p, t, rank);
        const uint32_t stage_bytes = (uint32_t)tc.nt * p.SBC * (CG == 2 && BCAT ? 6 : 2 * p.b_planes);
        const uint32_t part = stage_bytes / p.cs;  // pair (CG == 2): this CTA's half of the stage = its N / 2 rows of B
        // the tile's stages are contiguous in issue order (pair: each half of the columns is, [N tile][half][stage]);
        // a ring slot takes up to gst of them in ONE copy
        int rem = taps * (nsb_full * (tc.cb1 - tc.cb0) - (tc.cb1 == ncb ? nsb_full - p.nsb_last : 0));
        const size_t s0 = (size_t)tc.cb0 * taps * nsb_full;
        const uint32_t unit = CG == 2 ? part : stage_bytes;  // bytes per stage in this CTA's stream
        const uint8_t* src = p.wpk + (size_t)(tc.n0 / p.NT) * ntile_stride + s0 * unit +
                             (CG == 2 ? (size_t)rank * stages_per_ntile * part : (size_t)0);
        while (rem > 0) {
          const int g = rem < p.gst ? rem : p.gst;
          const uint32_t s = bs, ph = bph, bytes = (uint32_t)g * unit;
          if (++bs == (uint32_t)p.n_bstages) { bs = 0; bph ^= 1; }
          if (turn == lane) {
            mbar_wait(B_EMPTY + s * 8, ph ^ 1);
            mbar_arrive_expect_tx(B_FULL + s * 8, bytes);
            if (CG == 2 || p.cs == 1) {
              bulk_g2s(b_base + s * p.b_slot_bytes, src, bytes, B_FULL + s * 8);
            } else {  // this CTA fetches its 1/cs of the slot's bytes and multicasts it to every CTA of the cluster
              const uint32_t gp = bytes / p.cs;
              bulk_g2s_mc(b_base + s * p.b_slot_bytes + rank * gp, src + rank * gp, gp, B_FULL + s * 8, mc_mask);
            }
          }
          if (++turn == p.n_prod) turn = 0;
          src += bytes;
          rem -= g;
        }
        if (RES) break;  // the ring now holds every stage of the (single) N tile for the rest of the kernel
      }
    }
  } else if (CG == 2 && warp == 17 && rank != 0) {
    // ===================== pair, peer CTA: weight-stage relay =====================
    // follows the leader's issue order; each landed half-stage of this CTA becomes one arrival on the LEADER's B_FULL
    if (lane == 0) {
      const int nsb_full = p.KC / p.SBC;
      const uint32_t lead_full = mapa_u32(B_FULL, 0);
      uint32_t bs = 0, bph = 0;
      for (int t = cluster_id; t < p.total_groups; t += n_clusters) {
        const TileCoord tc = tile_coord(p, t, rank);
        int rem = taps * (nsb_full * (tc.cb1 - tc.cb0) - (tc.cb1 == ncb ? nsb_full - p.nsb_last : 0));
        for (; rem > 0; rem -= p.gst) {  // one arrival per ring slot
          mbar_wait(B_FULL + bs * 8, bph);
          mbar_arrive_cluster(lead_full + bs * 8);
          if (++bs == (uint32_t)p.n_bstages) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else if (CG == 2 && warp == 19 && rank != 0) {
    // ===================== pair, peer CTA: activation-tile relay =====================
    if (lane == 0) {
      const uint32_t lead_full = mapa_u32(A_FULL, 0);
      uint32_t abuf = 0, aph = 0;
      for (int t = cluster_id; t < p.total_groups; t += n_clusters) {
        const TileCoord tc = tile_coord(p, t, rank);
        for (int cb = tc.cb0; cb < tc.cb1; ++cb) {
          mbar_wait(A_FULL + abuf * 8, aph);
          fence_proxy_async();  // this CTA's cp.async writes (generic proxy) -> the pair's MMA reads (async proxy)
          mbar_arrive_cluster(lead_full + abuf * 8);
          if (++abuf == (uint32_t)p.n_abuf) { abuf = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp == 17) {
    // ===================== MMA issuer =====================
    // The issue loop is the critical path of the kernel: one MMA has to leave every ~nt/2 cycles.  All 32
    // lanes run the (warp-uniform) loop so that the address arithmetic stays on the uniform datapath; lane 0
    // alone issues.  Descriptors differ only in their 14-bit start-address field: 32-bit adds per MMA.
    // This warp has the highest id of its scheduler partition (the arbiter favours high ids).
    const bool leader = elect_one();
    uint32_t abuf = 0, aph = 0, acc = 0, tph = 0;  // activation-buffer / accumulator ring positions and phases
    uint32_t bs = 0, bph = 0;  // weight-stage ring position and phase
    const int nsb_full = p.KC / p.SBC;
    const uint32_t a_hi32 = ((uint32_t)(p.PW * 16) >> 4) | (1u << 14);  // SBO, descriptor version 1
    const uint32_t b_hi32 = (128u >> 4) | (1u << 14);
    const uint32_t a_lbo16 = (uint32_t)p.CGS >> 4;
    const uint32_t a_plane16 = (uint32_t)p.a_plane_bytes >> 4;
    const uint32_t a_step16 = 2 * a_lbo16;  // two 8-channel groups per k16 step
    const int k16_per_stage = p.SBC / 16;
    const uint32_t row_skip16 = (uint32_t)(p.PW - p.ks + 1);
    const int passes = (p.a_planes == 2 ? 1 : 0) | (p.b_planes == 2 ? 2 : 0);
    for (int t = cluster_id; t < p.total_groups; t += n_clusters) {
      const TileCoord tc = tile_coord(p, t, rank);
      if (CG == 2) mbar_wait_cluster(T_EMPTY + acc * 8, tph ^ 1); else mbar_wait(T_EMPTY + acc * 8, tph ^ 1);
      tc_fence_after();
      if (RES) bs = 0;  // stage i of the tile lives in ring slot i
      const bool b_wait = !RES || t == cluster_id;
      const uint32_t d_tmem = tmem_base + acc * p.acc_stride;
      const uint32_t d_tmem1 = d_tmem + p.sub_stride;  // second pixel tile (mt == 2): 8 pixels = 8 16-byte rows further in the halo
      constexpr bool two = MT == 2;
      const uint32_t idesc = make_idesc(tc.nt, 128 * CG);
      const uint32_t idesc2 = make_idesc(2 * tc.nt, 128 * CG);  // bcat: both weight planes as one operand
      // k-group stride: nt (bcat: 2 nt; pair: this CTA's nt / 2, pair + bcat: nt, see below) rows * 16 bytes >> 4
      const uint32_t b_lbo16 = CG == 2 ? (uint32_t)tc.nt >> (BCAT ? 0 : 1) : (uint32_t)tc.nt << BCAT;
      const uint32_t b_plane16 = (uint32_t)((tc.nt / CG) * p.SBC * 2) >> 4;
      const uint32_t b_step16 = 2 * b_lbo16;
      uint32_t accum = 0;
      if (RES) {
        // Resident weights (ks <= 3, one activation unit, one stage per tap): the whole tile is a fully unrolled
        // sequence of MMAs whose descriptors are INDEPENDENT adds off two bases.  The generic loops below carry
        // their descriptors through dependent uniform-datapath adds, ~10 cycles per instruction for this lone
        // warp: ~250 cycles per MMA on the head's 27-MMA tiles, far above the MMAs themselves.
        mbar_wait(A_FULL + abuf * 8, aph);
        fence_proxy_async();
        tc_fence_after();
        if (b_wait) {
          for (int s2 = 0; s2 < taps; ++s2) mbar_wait(B_FULL + s2 * 8, 0);
          tc_fence_after();
        }
        const uint32_t a0 = (((a_base + abuf * p.a_buf_bytes) & 0x3FFFFu) >> 4) | (a_lbo16 << 16);
        const uint32_t b0 = ((b_base & 0x3FFFFu) >> 4) | (b_lbo16 << 16);
        const uint32_t stage16 = (uint32_t)p.b_stage_bytes >> 4;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          if (kh >= p.ks) break;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            if (kw >= p.ks) break;
            const uint32_t at = a0 + (uint32_t)(kh * p.PW + kw);
            const uint32_t bt = b0 + (uint32_t)(kh * p.ks + kw) * stage16;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              if (j >= k16_per_stage) break;
              const uint32_t aj = at + j * a_step16, bj = bt + j * b_step16;
              if (leader) {
                umma_bf16_w(d_tmem, aj, a_hi32, bj, b_hi32, idesc, (kh | kw | j) ? 1u : 0u);
                if (passes & 1) umma_bf16_w(d_tmem, aj + a_plane16, a_hi32, bj, b_hi32, idesc, 1);
                if (passes & 2) umma_bf16_w(d_tmem, aj, a_hi32, bj + b_plane16, b_hi32, idesc, 1);
              }
            }
          }
        }
        if (leader) umma_commit(A_EMPTY + abuf * 8);
        if (++abuf == (uint32_t)p.n_abuf) { abuf = 0; aph ^= 1; }
      } else {
      // weight stages left in this tile / in the current ring slot, position inside the slot
      int rem = taps * (nsb_full * (tc.cb1 - tc.cb0) - (tc.cb1 == ncb ? nsb_full - p.nsb_last : 0));
      int glen = 0, gi = 0;
      const uint32_t stage16 = CG == 2 && BCAT ? (uint32_t)(tc.nt * p.SBC * 3) >> 4
                                               : (uint32_t)((tc.nt / CG) * p.SBC * 2 * p.b_planes) >> 4;  // this tile's bytes per stage / 16
      uint32_t b_grp = 0;
      for (int cb = tc.cb0; cb < tc.cb1; ++cb) {
        if (CG == 2) mbar_wait_cluster(A_FULL + abuf * 8, aph); else mbar_wait(A_FULL + abuf * 8, aph);
        fence_proxy_async();  // cp.async wrote the tile through the generic proxy; the MMA reads it through the async proxy
        tc_fence_after();
        const uint32_t a_buf16 = (((a_base + abuf * p.a_buf_bytes) & 0x3FFFFu) >> 4) | (a_lbo16 << 16);
        const int nsb = cb == ncb - 1 ? p.nsb_last : nsb_full;
        uint32_t a_tap = a_buf16;  // + (kh * PW + kw) 16-byte rows
        int kw = 0;
        for (int tap = 0; tap < taps; ++tap) {
          uint32_t a_lo = a_tap;
          for (int sb = 0; sb < nsb; ++sb) {
            if (gi == 0) {  // a new ring slot: up to gst stages behind one barrier
              glen = rem < p.gst ? rem : p.gst;
              if (CG == 2) mbar_wait_cluster(B_FULL + bs * 8, bph); else mbar_wait(B_FULL + bs * 8, bph);
              tc_fence_after();
              b_grp = (((b_base + bs * p.b_slot_bytes) & 0x3FFFFu) >> 4) | (b_lbo16 << 16);
            }
            uint32_t b_lo = b_grp;
            b_grp += stage16;
            // uniform loops; only the MMA itself is predicated on the leader lane, so that ptxas keeps the
            // descriptors in uniform registers instead of broadcasting them per instruction
            if (CG == 2 && BCAT) {
              // Pair MMAs with the weight planes side by side: A_hi x [B_hi | B_lo] is ONE MMA of 2 nt columns -- the leader's
              // shared memory holds B_hi (nt rows per k-group), the peer's B_lo at the same offsets -- then A_lo x B_hi with
              // B_hi split in halves between the CTAs, stored a second time behind the first region (k-group stride nt / 2).
              // 2 MMA instructions per K step and pixel tile instead of 3 (narrow N: 41 cycles per N = 48 pair MMA against a
              // pipe time of 24, tools/umma_bench2.cu) for 1.5x the weight bytes.
              const uint32_t y_off16 = (uint32_t)(tc.nt * p.SBC * 2) >> 4;  // region X: SBC / 8 k-groups x nt rows x 16 bytes
              uint32_t by = ((b_lo + y_off16) & 0xFFFFu) | ((uint32_t)(tc.nt >> 1) << 16);
#pragma unroll 1
              for (int j = 0; j < k16_per_stage; ++j) {
                if (leader) {
#pragma unroll
                  for (int m = 0; m < MT; ++m) {
                    const uint32_t d = m == 0 ? d_tmem : d_tmem1, am = a_lo + 8 * m;
                    umma2_w(d, am, a_hi32, b_lo, b_hi32, idesc2, accum);
                    umma2_w(d, am + a_plane16, a_hi32, by, b_hi32, idesc, 1);
                  }
                }
                accum = 1;
                a_lo += a_step16;
                b_lo += b_step16;
                by += (uint32_t)tc.nt;  // two k-groups of nt / 2 rows
              }
            } else if (CG == 2) {
              // pair MMAs; where one A tile meets both weight planes it is read from shared memory once (A collector)
#pragma unroll 1
              for (int j = 0; j < k16_per_stage; ++j) {
                if (leader) {
#pragma unroll
                  for (int m = 0; m < MT; ++m) {
                    const uint32_t d = m == 0 ? d_tmem : d_tmem1, am = a_lo + 8 * m;
                    if (passes & 2) {
                      umma2_w_keep(d, am, a_hi32, b_lo, b_hi32, idesc, accum);
                      umma2_w_reuse(d, am, a_hi32, b_lo + b_plane16, b_hi32, idesc, 1);
                    } else {
                      umma2_w(d, am, a_hi32, b_lo, b_hi32, idesc, accum);
                    }
                    if (passes & 1) umma2_w(d, am + a_plane16, a_hi32, b_lo, b_hi32, idesc, 1);
                  }
                }
                accum = 1;
                a_lo += a_step16;
                b_lo += b_step16;
              }
            } else if (BCAT) {
              // 2 MMAs per k16 instead of 3: the A tile (4 KB of shared-memory reads per MMA, the binding resource
              // at narrow N) is fetched twice, not three times
#pragma unroll 1
              for (int j = 0; j < k16_per_stage; ++j) {
                if (leader) {
                  umma_bf16_w(d_tmem, a_lo, a_hi32, b_lo, b_hi32, idesc2, accum);
                  umma_bf16_w(d_tmem, a_lo + a_plane16, a_hi32, b_lo, b_hi32, idesc, 1);
                  if (two) {
                    umma_bf16_w(d_tmem1, a_lo + 8, a_hi32, b_lo, b_hi32, idesc2, accum);
                    umma_bf16_w(d_tmem1, a_lo + 8 + a_plane16, a_hi32, b_lo, b_hi32, idesc, 1);
                  }
                }
                accum = 1;
                a_lo += a_step16;
                b_lo += b_step16;
              }
            } else if (passes == 3) {
#pragma unroll 1
              for (int j = 0; j < k16_per_stage; ++j) {
                if (leader) {
                  umma_bf16_w(d_tmem, a_lo, a_hi32, b_lo, b_hi32, idesc, accum);
                  umma_bf16_w(d_tmem, a_lo + a_plane16, a_hi32, b_lo, b_hi32, idesc, 1);
                  umma_bf16_w(d_tmem, a_lo, a_hi32, b_lo + b_plane16, b_hi32, idesc, 1);
                  if (two) {
                    umma_bf16_w(d_tmem1, a_lo + 8, a_hi32, b_lo, b_hi32, idesc, accum);
                    umma_bf16_w(d_tmem1, a_lo + 8 + a_plane16, a_hi32, b_lo, b_hi32, idesc, 1);
                    umma_bf16_w(d_tmem1, a_lo + 8, a_hi32, b_lo + b_plane16, b_hi32, idesc, 1);
                  }
                }
                accum = 1;
                a_lo += a_step16;
                b_lo += b_step16;
              }
            } else if (passes == 1) {
#pragma unroll 1
              for (int j = 0; j < k16_per_stage; ++j) {
                if (leader) {
                  umma_bf16_w(d_tmem, a_lo, a_hi32, b_lo, b_hi32, idesc, accum);
                  umma_bf16_w(d_tmem, a_lo + a_plane16, a_hi32, b_lo, b_hi32, idesc, 1);
                  if (two) {
                    umma_bf16_w(d_tmem1, a_lo + 8, a_hi32, b_lo, b_hi32, idesc, accum);
                    umma_bf16_w(d_tmem1, a_lo + 8 + a_plane16, a_hi32, b_lo, b_hi32, idesc, 1);
                  }
                }
                accum = 1;
                a_lo += a_step16;
                b_lo += b_step16;
              }
            } else {
#pragma unroll 1
              for (int j = 0; j < k16_per_stage; ++j) {
                if (leader) {
                  umma_bf16_w(d_tmem, a_lo, a_hi32, b_lo, b_hi32, idesc, accum);
                  if (passes == 2) umma_bf16_w(d_tmem, a_lo, a_hi32, b_lo + b_plane16, b_hi32, idesc, 1);
                  if (two) {
                    umma_bf16_w(d_tmem1, a_lo + 8, a_hi32, b_lo, b_hi32, idesc, accum);
                    if (passes == 2) umma_bf16_w(d_tmem1, a_lo + 8, a_hi32, b_lo + b_plane16, b_hi32, idesc, 1);
                  }
                }
                accum = 1;
                a_lo += a_step16;
                b_lo += b_step16;
              }
            }
            // slot free once these MMAs have read it -- signalled to every CTA that multicasts into it
            if (++gi == glen) {
              if (leader) commit_stage<CG>(B_EMPTY + bs * 8, p.cs, mc_mask);
              if (++bs == (uint32_t)p.n_bstages) { bs = 0; bph ^= 1; }
              rem -= glen;
              gi = 0;
            }
          }
          if (++kw == p.ks) { kw = 0; a_tap += row_skip16; } else { ++a_tap; }
        }
        if (leader) { if (CG == 2) umma2_commit_mc(A_EMPTY + abuf * 8, 3); else umma_commit(A_EMPTY + abuf * 8); }
        if (++abuf == (uint32_t)p.n_abuf) { abuf = 0; aph ^= 1; }
      }
      }
      if (leader) { if (CG == 2) umma2_commit_mc(T_FULL + acc * 8, 3); else umma_commit(T_FULL + acc * 8); }
      if (++acc == (uint32_t)p.n_acc) { acc = 0; tph ^= 1; }
    }
  } else if (warp >= p.n_epi && warp < TC_WORK_WARPS) {
    // ===================== activation loaders: split-bf16 NHWC -> halo tile, 16-byte cp.async =====================
    const int ltid = threadIdx.x - p.n_epi * 32;
    const int nload = (TC_WORK_WARPS - p.n_epi) * 32;
    const int npix = p.PW * p.PH;
    uint32_t abuf = 0, aph = 0;
    for (int t = cluster_id; t < p.total_groups; t += n_clusters) {
      const TileCoord tc = tile_coord(p, t, rank);
      const uint8_t* img = p.in + (size_t)tc.img * p.h * p.w * p.in_stride * 2;
      for (int cb = tc.cb0; cb < tc.cb1; ++cb) {
        const int c0 = cb * p.KC;
        const int ncg = min(p.KC, p.C - c0) >> 3;
        mbar_wait(A_EMPTY + abuf * 8, aph ^ 1);
        const uint32_t dst = a_base + abuf * p.a_buf_bytes;
        // lanes 2k / 2k+1 copy the two 16-byte halves (adjacent channel groups) of one 32-byte sector of the
        // same pixel; consecutive lane pairs take consecutive pixels
        const int npair = (ncg + 1) >> 1;
        const int tasks = npix * npair;
        const int cgp = ltid & 1;
        int j = ltid >> 1;
        int cpi = fdiv(j, p.fd_npix), pix = j - cpi * npix;
        for (; j < tasks; j += nload / 2) {
          const int cgi = 2 * cpi + cgp;
          if (cgi < ncg) {
            const int py = fdiv(pix, p.fd_PW), px = pix - py * p.PW;
            const int gy = tc.y0 + py - p.pad, gx = tc.x0 + px - p.pad;
            const int ch = c0 + cgi * 8;
            const bool ok = tc.real && (unsigned)gy < (unsigned)p.h && (unsigned)gx < (unsigned)p.w && ch < p.c_valid;
            const uint8_t* src = ok ? img + ((size_t)(gy * p.w + gx) * p.in_stride + ch) * 2 : p.in;
            const uint32_t d = dst + cgi * p.CGS + pix * 16;
            if (!(p.skip & 4)) {
            cp_async16(d, src, ok ? 16u : 0u);
            if (p.a_planes == 2) cp_async16(d + p.a_plane_bytes, src + p.in_plane_bytes, ok ? 16u : 0u);  // plane 1 of the dummy address is valid
            }
          }
          pix += nload / 2;
          while (pix >= npix) { pix -= npix; ++cpi; }
        }
        cp_async_arrive(A_FULL + abuf * 8);
        if (++abuf == (uint32_t)p.n_abuf) { abuf = 0; aph ^= 1; }
      }
    }
  } else if (warp < p.n_epi) {
    // ===================== epilogue: TMEM -> registers -> smem transpose -> global (8 warps) =====================
    // Two warps per TMEM lane quarter; they take alternate 16-column chunks.  tcgen05.ld hands every lane one
    // ROW (pixel) of the chunk; storing from that layout touches 32 different sectors per instruction with 8-16
    // useful bytes each.  The chunk is therefore transposed through a per-warp staging tile so that 4 lanes hold
    // the 16 consecutive channels of one pixel: every store (and the z / scale / bias load) then moves whole
    // 32-byte sectors.
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int half = warp >> 2;
    float* stg = reinterpret_cast<float*>(smem + p.epi_stage_off) + warp * (32 * EPI_ROW);
    const int qd = lane & 3;         // 4-channel quad of the chunk this lane stores
    const int rsub = lane >> 2;      // row (pixel) within each group of 8 rows
    uint32_t acc_next = 0, tph_next = 0;
    float head_loss_acc = 0.f;
    for (int t = cluster_id; t < p.total_groups; t += n_clusters) {
      const TileCoord tc = tile_coord(p, t, rank);
      const uint32_t acc = acc_next, tph = tph_next;
      if (++acc_next == (uint32_t)p.n_acc) { acc_next = 0; tph_next ^= 1; }
      // the four pixels (one per 8-row group) this lane stores: m = q*32 + it*8 + rsub, tile row m>>3 = q*4 + it
      size_t row_base[4], zrow[4];
      bool valid[4];
      int sub = 0;  // pixel tile of the step (0 .. mt-1)
    next_sub:
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int y = tc.y0 + q * 4 + it, x = tc.x0 + sub * TILE_W + rsub;
        valid[it] = tc.real && y < p.h && x < p.w;
        // pixel indices fit 32 bits (checked at launch); one 64-bit multiply per offset
        if (p.epi == 0) {
          row_base[it] = (size_t)((tc.img * p.out_h + y * p.rh) * p.out_w + x * p.rw) * (size_t)p.cg;
          zrow[it] = 0;
        } else {
          const int qh = fdiv(y, p.fd_rh), si = y - qh * p.rh, qw = fdiv(x, p.fd_rw), sj = x - qw * p.rw;
          row_base[it] = (size_t)((tc.img * p.out_h + qh) * p.out_w + qw) * (size_t)p.quad_stride + (size_t)((si * p.rw + sj) * p.n_store);
          zrow[it] = (size_t)((tc.img * p.h + y) * p.w + x) * (size_t)p.n_store;
        }
      }
      if (p.epi == 2) {
        // head: every lane keeps its own accumulator row = one pixel, 3 real columns: OutImg (models/_layers.py:10-16)
        // + lp_loss partial sum (quantizer.py:66-73) + dL/dz.  The target is requested before the accumulator wait.
        const int m = q * 32 + lane;
        const int y = tc.y0 + (m >> 3), x = tc.x0 + (m & 7);
        const bool ok = tc.real && y < p.h && x < p.w && half == 0;
        const size_t plane = (size_t)p.h * p.w;
        const size_t o = (size_t)tc.img * 3 * plane + (size_t)y * p.w + x;
        float tg[3] = {0.f, 0.f, 0.f};
        if (ok && p.head_target) {
#pragma unroll
          for (int c = 0; c < 3; ++c) tg[c] = __ldg(p.head_target + o + c * plane);
        }
        mbar_wait(T_FULL + acc * 8, tph);
        tc_fence_after();
        if (half == 0) {
          uint32_t v[16];
          tmem_ld16(tmem_base + acc * p.acc_stride + ((uint32_t)(q * 32) << 16), v);
          tmem_ld_wait();
          if (ok) {
            float gr[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              float val = __uint_as_float(v[c]);
              if (p.scale) val *= __ldg(p.scale + c);
              val += __ldg(p.bias + c);
              float outv, dout;
              if (p.head_out_bias == 0) {
                const float th = tanhf(val);
                outv = th * 0.5f + 0.5f;
                dout = 0.5f * (1.0f - th * th);
              } else {
                outv = sigmoid_f(val);
                dout = outv * (1.0f - outv);
              }
              if (p.head_img) p.head_img[o + c * plane] = outv;
              if (p.head_target) {
                const float dlt = outv - tg[c];
                const float a = fabsf(dlt);
                if (p.head_p == 2.0f) {
                  head_loss_acc += dlt * dlt;
                  gr[c] = 2.0f * dlt * p.head_inv_mean * dout;
                } else {
                  head_loss_acc += powf(a, p.head_p);
                  const float sgn = dlt > 0.f ? 1.f : (dlt < 0.f ? -1.f : 0.f);
                  gr[c] = p.head_p * powf(a, p.head_p - 1.0f) * sgn * p.head_inv_mean * dout;
                }
              }
            }
            if (p.head_dz) {  // (n, h, w, 8) split-bf16 planes: 3 real channels + zeros
              uint32_t hb[3], lb[3];
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const __nv_bfloat16 hv = __float2bfloat16_rn(gr[c]);
                hb[c] = __bfloat16_as_ushort(hv);
                lb[c] = __bfloat16_as_ushort(__float2bfloat16_rn(gr[c] - __bfloat162float(hv)));
              }
              const size_t px = (((size_t)tc.img * p.h + y) * p.w + x) * 16;  // 8 channels * 2 bytes
              *reinterpret_cast<uint4*>(p.head_dz + px) = make_uint4(hb[0] | (hb[1] << 16), hb[2], 0u, 0u);
              *reinterpret_cast<uint4*>(p.head_dz + p.out_plane_bytes + px) = make_uint4(lb[0] | (lb[1] << 16), lb[2], 0u, 0u);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(T_EMPTY + acc * 8);
        continue;
      }
      mbar_wait(T_FULL + acc * 8, tph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * p.acc_stride + sub * p.sub_stride + ((uint32_t)(q * 32) << 16);
      for (int c0 = half * 16; c0 < tc.nt && !(p.skip & 1); c0 += (p.n_epi >> 2) * 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        if (BCAT) {  // hi*lo partial sums live nt columns further
          uint32_t v2[16];
          tmem_ld16(taddr + tc.nt + c0, v2);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) + __uint_as_float(v2[k]));
        }
        const int n = tc.n0 + c0 + qd * 4;  // first of this lane's 4 columns
        const bool col_ok = n < p.n_store;
        float4 g0 = make_float4(1.f, 1.f, 1.f, 1.f), g1 = make_float4(0.f, 0.f, 0.f, 0.f);
        size_t col_off = (size_t)n;  // dgrad: plain column; fwd: (group, channel) of the shuffle
        if (p.epi != 1 && col_ok) {
          if (p.scale) g0 = __ldg(reinterpret_cast<const float4*>(p.scale + n));
          if (p.bias) g1 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
          const int grp = fdiv(n, p.fd_cg), c = n - grp * p.cg;
          const int si = fdiv(grp, p.fd_rw), sj = grp - si * p.rw;
          col_off = ((size_t)si * (p.w * p.rw) + sj) * p.cg + c;
        }
        float4 zv[4];  // dgrad: z of the previous stage for this lane's 4 pixels, requested before the TMEM wait
        if (p.epi == 1) {
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            zv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.zprev && p.act != 0 && valid[it] && col_ok)
              zv[it] = __ldg(reinterpret_cast<const float4*>(p.zprev + zrow[it] + n));
          }
        }
        tmem_ld_wait();
        {  // own row -> staging tile (row stride EPI_ROW floats keeps the 16-byte stores conflict free)
          float4* rowp = reinterpret_cast<float4*>(stg + lane * EPI_ROW);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            rowp[k] = make_float4(__uint_as_float(v[4 * k]), __uint_as_float(v[4 * k + 1]), __uint_as_float(v[4 * k + 2]),
                                  __uint_as_float(v[4 * k + 3]));
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          float4 r = *reinterpret_cast<const float4*>(stg + (it * 8 + rsub) * EPI_ROW + qd * 4);
          if (!(valid[it] && col_ok)) continue;
          if (p.epi == 0) {
            r.x = fmaf(r.x, g0.x, g1.x); r.y = fmaf(r.y, g0.y, g1.y);
            r.z = fmaf(r.z, g0.z, g1.z); r.w = fmaf(r.w, g0.w, g1.w);
            const size_t o = row_base[it] + col_off;
            if (p.act == 2) {  // GELU, keeping GELU'(z) for the backward pass in place of z (one exponential for both)
              float4 g;
              gelu_both_fast(r.x, r.x, g.x); gelu_both_fast(r.y, r.y, g.y);
              gelu_both_fast(r.z, r.z, g.z); gelu_both_fast(r.w, r.w, g.w);
              if (p.out_z) *reinterpret_cast<float4*>(p.out_z + o) = g;
              if (p.out_y) store_split4(p.out_y, p.out_plane_bytes, o, r);
            } else {
              if (p.out_z) *reinterpret_cast<float4*>(p.out_z + o) = r;
              if (p.out_y) {
                if (p.act == 1) { r.x = gelu_fast(r.x); r.y = gelu_fast(r.y); r.z = gelu_fast(r.z); r.w = gelu_fast(r.w); }
                store_split4(p.out_y, p.out_plane_bytes, o, r);
              }
            }
          } else if (p.epi == 1) {
            if (p.zprev && p.act == 1) {
              r.x *= gelu_grad_fast(zv[it].x); r.y *= gelu_grad_fast(zv[it].y);
              r.z *= gelu_grad_fast(zv[it].z); r.w *= gelu_grad_fast(zv[it].w);
            } else if (p.zprev && p.act == 2) {  // z_prev holds GELU'(z) already
              r.x *= zv[it].x; r.y *= zv[it].y; r.z *= zv[it].z; r.w *= zv[it].w;
            }
            store_split4(p.out_y, p.out_plane_bytes, row_base[it] + n, r);
          } else if (p.epi == 3) {  // split-K partial: raw sums in the input grid's own order, finished by dgrad_finish_kernel
            *reinterpret_cast<float4*>(p.part + (size_t)tc.ksp * p.part_stride + zrow[it] + n) = r;
          }
        }
        __syncwarp();
      }
      if (MT > 1 && ++sub < MT) goto next_sub;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(mapa_u32(T_EMPTY + acc * 8, 0));  // the pair leader issues for both CTAs
        else mbar_arrive(T_EMPTY + acc * 8);
      }
    }
    if (p.epi == 2 && p.head_loss != nullptr) {
      head_loss_acc = warp_sum(head_loss_acc);
      if (lane == 0 && head_loss_acc != 0.f) atomicAdd(p.head_loss, head_loss_acc);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    p.dbg[0] = clock64() - dbg_c0;
    p.dbg[1] = t1 - dbg_t0;
  }
  if (p.cs > 1) cluster_sync_all();  // no CTA exits while a peer may still multicast into it / signal its barriers
  if (warp == 18) {
    tc_fence_after();
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// weight packing into the stage order above
//   stage layout: [plane][k-group of 8][n in tile][8 k-channels] bf16
//   n tile t: stages cb-major, then tap, then sub-block; every stage of tile t has nt(t) columns
// ------------------------------------------------------------------------------------------------
struct TcPackParams {
  const float* w;     // reference layout (cout, cin_src, k, k): codes, or de-quantised weights
  const float* zp;    // per-cout zero point subtracted from w (integer-domain forward), or null
  int zp_stride;      // 1: per output channel, 0: one value
  uint8_t* out;
  int cout, cin, cin_src, ks;
  int rh, rw, c_grp, cg;  // packed output-channel order (up-shuffle groups)
  int dir;                // 0 forward (K = input channels, N = packed output channels), 1 dgrad (swapped, flipped)
  int C, N, NT, KC, SBC, b_planes, bcat;
  int cg2;                // CTA-pair plan: every stage stored per half of its columns, [half][plane][k-group][n / 2][8]
};

__device__ __forceinline__ int unpack_cout(int np, int rh, int rw, int c_grp, int cg) {
  // packed n' = (i*rw + j)*cg + c  ->  reference channel c*rh*rw + i*rw + j   (-1 for a pad column)
  const int grp = np / cg, c = np - grp * cg;
  if (c >= c_grp || grp >= rh * rw) return -1;
  return c * rh * rw + grp;
}

__device__ __forceinline__ void tc_pack_body(const TcPackParams& q, long long e_first, long long e_step) {
  const int taps = q.ks * q.ks;
  const int tiles_n = (q.N + q.NT - 1) / q.NT;
  const int stages_per_ntile = (q.C / q.SBC) * taps;
  const int nsb_full = q.KC / q.SBC;
  const int g_per_stage = q.SBC / 8;
  // one thread per (n tile, stage, k-group, column): both planes
  const long long per_tile_full = (long long)stages_per_ntile * g_per_stage * q.NT;
  const long long total = per_tile_full * tiles_n;
  for (long long e = e_first; e < total; e += e_step) {
    const int tn = (int)(e / per_tile_full);
    long long r = e - (long long)tn * per_tile_full;
    const int nt = min(q.NT, q.N - tn * q.NT);
    const int nn = (int)(r % q.NT);
    r /= q.NT;
    if (nn >= nt) continue;
    const int g = (int)(r % g_per_stage);
    const int s = (int)(r / g_per_stage);
    // stage -> (cb, tap, sb)
    const int full_block_stages = taps * nsb_full;
    int cb = s / full_block_stages;
    int rem = s - cb * full_block_stages;
    const int ncb_full = q.C / q.KC;
    int nsb = nsb_full;
    if (cb >= ncb_full) {
      cb = ncb_full;
      rem = s - ncb_full * full_block_stages;
      nsb = (q.C - ncb_full * q.KC) / q.SBC;
    }
    const int tap = rem / nsb, sb = rem - tap * nsb;
    const int k0 = cb * q.KC + sb * q.SBC + g * 8;  // first K channel of this chunk
    const int n = tn * q.NT + nn;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = k0 + i;
      int co, ci, t;
      if (q.dir == 0) {
        co = unpack_cout(n, q.rh, q.rw, q.c_grp, q.cg);
        ci = k;
        t = tap;
      } else {
        co = unpack_cout(k, q.rh, q.rw, q.c_grp, q.cg);
        ci = n;
        t = taps - 1 - tap;
      }
      float x = 0.f;
      if (co >= 0 && co < q.cout && ci < q.cin) {
        x = q.w[((size_t)co * q.cin_src + ci) * taps + t];
        if (q.zp) x -= q.zp[co * q.zp_stride];
      }
      v[i] = x;
    }
    uint4 hi, lo;
    split8(make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]), hi, lo);
    const size_t stage_bytes = (size_t)nt * q.SBC * 2 * q.b_planes;
    const size_t ntile_stride = (size_t)stages_per_ntile * q.NT * q.SBC * 2 * q.b_planes;
    uint8_t* st = q.out + (size_t)tn * ntile_stride + (size_t)s * stage_bytes;
    if (q.cg2 && q.bcat) {
      // pair plan with side-by-side planes, per CTA and stage: region X [k-group][nt][8] (leader: hi plane, peer: lo plane),
      // then region Y [k-group][nt / 2][8] = this CTA's half of the hi plane's columns
      const int nh = nt >> 1, h = nn >= nh ? 1 : 0, nl = nn - h * nh;
      const size_t part = (size_t)nt * q.SBC * 3, xbytes = (size_t)nt * q.SBC * 2;
      const size_t nts = (size_t)stages_per_ntile * q.NT * q.SBC * 6;
      uint8_t* r0 = q.out + (size_t)tn * nts + (size_t)s * part;
      uint8_t* r1 = r0 + (size_t)stages_per_ntile * part;
      *reinterpret_cast<uint4*>(r0 + ((size_t)g * nt + nn) * 16) = hi;
      *reinterpret_cast<uint4*>(r1 + ((size_t)g * nt + nn) * 16) = lo;
      *reinterpret_cast<uint4*>((h ? r1 : r0) + xbytes + ((size_t)g * nh + nl) * 16) = hi;
    } else if (q.cg2) {  // [half of the columns][stage][plane][k-group][n in half][8]: each CTA of the pair streams ONE contiguous region
      const int nh = nt >> 1, h = nn >= nh ? 1 : 0, nl = nn - h * nh;
      uint8_t* hb = q.out + (size_t)tn * ntile_stride + ((size_t)h * stages_per_ntile + s) * (stage_bytes >> 1);
      const size_t o = ((size_t)g * nh + nl) * 16;
      *reinterpret_cast<uint4*>(hb + o) = hi;
      if (q.b_planes == 2) *reinterpret_cast<uint4*>(hb + (size_t)nh * q.SBC * 2 + o) = lo;
    } else if (q.bcat) {  // [k-group][plane][n][8]
      const size_t o = ((size_t)g * 2 * nt + nn) * 16;
      *reinterpret_cast<uint4*>(st + o) = hi;
      *reinterpret_cast<uint4*>(st + o + (size_t)nt * 16) = lo;
    } else {  // [plane][k-group][n][8]
      const size_t o = ((size_t)g * nt + nn) * 16;
      *reinterpret_cast<uint4*>(st + o) = hi;
      if (q.b_planes == 2) *reinterpret_cast<uint4*>(st + (size_t)nt * q.SBC * 2 + o) = lo;
    }
  }
}

__global__ void __launch_bounds__(256) tc_pack_kernel(const TcPackParams q) {
  tc_pack_body(q, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

// Several packs (every stage's forward and data-gradient operand, plus the per-column epilogue vectors) in one
// launch: a block belongs to one task and strides over that task's elements with the task's own block count.
struct TcPackVec {
  const float* delta; const float* bias; float* scale_out; float* bias_out;
  int d_stride, cout, rh, rw, c_grp, cg, N;
};
struct TcPackMulti {
  TcPackParams t[NQ_MULTI_MAX];
  TcPackVec v[NQ_MULTI_MAX];   // v[i].N == 0: task i packs no vectors
  int blk_start[NQ_MULTI_MAX + 1];
  int n;
};
__device__ __forceinline__ void tc_pack_vec_body(const TcPackVec& v, int first, int step);

// per-column epilogue vectors in packed order: scale[n] = delta[co] (or 1), bias[n] = b[co] (or 0)
__device__ __forceinline__ void tc_pack_vec_body(const TcPackVec& v, int first, int step) {
  for (int n = first; n < v.N; n += step) {
    const int co = unpack_cout(n, v.rh, v.rw, v.c_grp, v.cg);
    const bool ok = co >= 0 && co < v.cout;
    if (v.scale_out) v.scale_out[n] = (ok && v.delta) ? v.delta[co * v.d_stride] : 1.0f;
    if (v.bias_out) v.bias_out[n] = (ok && v.bias) ? v.bias[co] : 0.0f;
  }
}
__global__ void __launch_bounds__(256) tc_pack_vec_kernel(const TcPackVec v) {
  tc_pack_vec_body(v, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
__global__ void __launch_bounds__(256) tc_pack_multi_kernel(const __grid_constant__ TcPackMulti m) {
  int ti = 0;
  while (ti + 1 < m.n && (int)blockIdx.x >= m.blk_start[ti + 1]) ++ti;
  const int lb = blockIdx.x - m.blk_start[ti], nb = m.blk_start[ti + 1] - m.blk_start[ti];
  if (m.t[ti].out != nullptr) tc_pack_body(m.t[ti], (long long)lb * blockDim.x + threadIdx.x, (long long)nb * blockDim.x);
  if (m.v[ti].N > 0 && lb == 0) tc_pack_vec_body(m.v[ti], threadIdx.x, blockDim.x);
}

int check_conv_desc(const nq_conv_desc* d);

static int fill_plan(const nq_conv_desc* d, int dir, int a_planes, int b_planes, nq_tc_plan* pl) {
  int st = check_conv_desc(d);
  if (st) return st;
  if (!pl || (dir != 0 && dir != 1) || a_planes < 1 || a_planes > 2 || b_planes < 1 || b_planes > 2) return NQ_ERR_BAD_ARG;
  const int nout_p = d->rh * d->rw * d->cg;
  int C = dir == 0 ? d->cin_p : nout_p;
  int N = dir == 0 ? nout_p : d->cin_p;
  // both GEMM dims may be padded to 16: channels >= the stored count are read as zero (c_valid), columns
  // >= the stored count are not written (n_store)
  C = (C + 15) / 16 * 16;
  N = (N + 15) / 16 * 16;
  if (C % 16 || N % 16 || d->ksize > 7) return NQ_ERR_BAD_SHAPE;
  pl->dir = dir;
  pl->C = C;
  pl->N = N;
  pl->a_planes = a_planes;
  pl->b_planes = b_planes;
  pl->NT = N < 256 ? N : 256;
  // Two pixel tiles per CTA step (side by side in x, one 16x16 halo) when the weight stream would otherwise bind:
  // every weight stage then feeds 2x the MMAs, halving the L2 -> SM weight traffic per pixel (conv_dgrad[5] of
  // HNeRV-3M re-streams 768 KB of weights per 128-pixel tile).  Needs both accumulators in one 256-column TMEM half
  // and enough tile pairs to fill the machine.
  {
    const long long w_tile_bytes = (long long)d->ksize * d->ksize * C * pl->NT * 2 * b_planes;
    const long long pairs = (long long)((d->w + 2 * TILE_W - 1) / (2 * TILE_W)) * ((d->h + TILE_H - 1) / TILE_H) * d->n;
    pl->mt = (pl->NT <= 128 && w_tile_bytes >= 64 * 1024 && pairs >= 2LL * sm_count()) ? 2 : 1;
    if (const char* e = getenv("NQ_TC_MT")) {  // tuning override
      const int v = atoi(e);
      if (v == 1 || (v == 2 && pl->NT <= 128)) pl->mt = v;
    }
  }
  // Planes side by side (see TcParams::bcat) when both operands are split, the doubled tile fits the accumulator
  // slot (256 columns, 128 per pixel tile when mt == 2) and K is long enough to amortise the second TMEM read of
  // the epilogue.
  // CTA pairs (cta_group::2, see conv_tc_kernel): whenever the weights of an N tile are too large to stay resident in the
  // ring (the small stages keep the single-CTA resident / multicast forms) and there are pixel-tile pairs to form.
  {
    const long long w_tile_bytes = (long long)d->ksize * d->ksize * C * pl->NT * 2 * b_planes;
    const long long tiles_m = (long long)((d->w + TILE_W * pl->mt - 1) / (TILE_W * pl->mt)) * ((d->h + TILE_H - 1) / TILE_H) * d->n;
    // ... and there are enough pixel tiles for the relay hop between the two CTAs to stay hidden (measured, HNeRV-3M stage 3 with
    // 50 tiles: forward 110 k -> 102 k cycles, split-K data gradient 162 k -> 135 k; stage 2 with 6 tiles: no difference)
    pl->cg2 = (w_tile_bytes > 96 * 1024 && tiles_m >= 32) ? 1 : 0;
    if (const char* e = getenv("NQ_TC_CG2")) {  // tuning override
      if (atoi(e) == 0) pl->cg2 = 0;
    }
  }
  pl->bcat = (a_planes == 2 && b_planes == 2 && 2 * pl->NT * pl->mt <= 256 && d->ksize * d->ksize * C >= 512) ? 1 : 0;
  if (const char* e = getenv("NQ_TC_BCAT")) {  // tuning override
    if (atoi(e) == 0) pl->bcat = 0;
    else if (a_planes == 2 && b_planes == 2 && 2 * pl->NT * pl->mt <= 256) pl->bcat = 1;
  }
  const int tile_w = TILE_W * pl->mt;
  pl->PW = tile_w + d->ksize - 1;
  pl->PH = TILE_H + d->ksize - 1;
  int npix = pl->PW * pl->PH;
  int cgs16 = npix;  // channel-group stride in 16-byte units, forced to 4 mod 8 (conflict-free lane-pair stores)
  while (cgs16 % 8 != 4) ++cgs16;
  pl->CGS = cgs16 * 16;
  // Weight stage = SBC channels of one tap; activation unit = KC channels of the halo tile.  Big stages
  // amortise the per-stage barrier round trips of the producer and issuer threads (a 6 KB stage is only
  // ~270 MMA cycles), so take the largest SBC (multiple of 16 dividing C) whose stage is <= 32 KB and that
  // leaves room for two activation buffers and >= 3 weight stages in the 227 KB of shared memory.
  int best = 0, best_kc = 0;
  for (int sbc = (C < 128 ? C : 128) / 16 * 16; sbc >= 16 && !best; sbc -= 16) {
    if (C % sbc) continue;
    const int stage = pl->cg2 && pl->bcat ? pl->NT * sbc * 3 : pl->NT * sbc * 2 * b_planes >> pl->cg2;  // pair: each CTA stages half of the columns (+ half a plane again when side by side)
    if (stage > 32 * 1024 && sbc > 16) continue;
    // activation unit: ~64 channels (a multiple of the stage), fewer when the (16x16-pixel) halo is large
    for (int kc = sbc >= 64 ? sbc : sbc * (64 / sbc); kc >= sbc; kc -= sbc) {
      const int kcc = kc > C ? C : kc;
      const int a_buf = pl->CGS * (kcc / 8) * a_planes;
      const int budget = 227 * 1024 - TC_HDR_BYTES - EPI_STAGE_BYTES - 2 * a_buf;
      if (budget < 3 * stage) continue;
      best = sbc;
      best_kc = kcc;
      break;
    }
  }
  if (!best) return NQ_ERR_UNSUPPORTED;
  const int sbc = best;
  pl->SBC = sbc;
  pl->KC = best_kc;
  pl->a_plane_bytes = pl->CGS * (pl->KC / 8);
  pl->a_buf_bytes = pl->a_plane_bytes * a_planes;
  pl->b_stage_bytes = pl->cg2 && pl->bcat ? pl->NT * sbc * 3 : pl->NT * sbc * 2 * b_planes >> pl->cg2;
  // Ring depths: as many activation buffers (<= 8) as fit next to ~48 KB of weight stages, and as many accumulator
  // slots as the 512 TMEM columns hold, so that several short-K tiles can be in flight.
  const int total = 227 * 1024 - TC_HDR_BYTES - EPI_STAGE_BYTES;
  int nst_min = 48 * 1024 / pl->b_stage_bytes;
  if (nst_min < 3) nst_min = 3;
  if (nst_min > TC_MAX_BSTAGES) nst_min = TC_MAX_BSTAGES;
  int n_abuf = (total - nst_min * pl->b_stage_bytes) / pl->a_buf_bytes;
  if (n_abuf > TC_MAX_RING) n_abuf = TC_MAX_RING;
  if (n_abuf < 2) n_abuf = 2;
  if (const char* e = getenv("NQ_TC_ABUF")) {  // tuning override
    const int v = atoi(e);
    if (v >= 2 && v <= TC_MAX_RING && total - v * pl->a_buf_bytes >= 2 * pl->b_stage_bytes) n_abuf = v;
  }
  // Weight stages per ring slot.  Every slot costs the issuing thread one barrier wait, one tcgen05.commit and the ring
  // arithmetic, ~250 cycles that the one-or-two-deep MMA queue hides only in part (HNeRV-3M stage 5: 720 cycles of MMAs per
  // stage, 970 per stage measured); slots of several stages amortise them, and one bulk copy moves the whole slot (a copy costs
  // its issuing lane ~1000 cycles whatever its size, tools/tma_bulk_bench.cu).  Largest slot <= 32 KB that leaves >= 3 slots,
  // trading activation buffers beyond the second for it.
  int gst = 1;
  if ((long long)d->ksize * d->ksize * C * pl->NT * 2 * b_planes > 96 * 1024) {  // not a candidate for resident weights
    const int stages_tile = (C / sbc) * d->ksize * d->ksize;
    for (int na = n_abuf; na >= 2; --na) {
      const int bud = total - na * pl->a_buf_bytes;
      int g = bud / 3 / pl->b_stage_bytes;
      if (g * pl->b_stage_bytes > 32 * 1024) g = 32 * 1024 / pl->b_stage_bytes;
      if (g > 8) g = 8;
      if (g > stages_tile) g = stages_tile;
      if (g > gst) { gst = g; n_abuf = na; }
    }
  }
  if (const char* e = getenv("NQ_TC_GST")) {  // tuning override
    const int v = atoi(e);
    if (v >= 1 && v <= 8 && (total - n_abuf * pl->a_buf_bytes) / (v * pl->b_stage_bytes) >= 2) gst = v;
  }
  pl->gst = gst;
  pl->n_abuf = n_abuf;
  const int budget = total - n_abuf * pl->a_buf_bytes;
  int nst = budget / (gst * pl->b_stage_bytes);
  if (nst > TC_MAX_BSTAGES) nst = TC_MAX_BSTAGES;
  if (nst < 2) return NQ_ERR_UNSUPPORTED;
  pl->n_bstages = nst;
  pl->smem_bytes = TC_HDR_BYTES + n_abuf * pl->a_buf_bytes + nst * gst * pl->b_stage_bytes + EPI_STAGE_BYTES;
  int sub_cols = pl->NT * (pl->bcat ? 2 : 1), sub_stride = 32;
  while (sub_stride < sub_cols) sub_stride *= 2;
  pl->acc_stride = sub_stride * pl->mt;
  pl->n_acc = 512 / pl->acc_stride;
  if (pl->n_acc > TC_MAX_RING) pl->n_acc = TC_MAX_RING;
  if (const char* e = getenv("NQ_TC_NACC")) {  // tuning override
    const int v = atoi(e);
    if (v >= 2 && v <= pl->n_acc) pl->n_acc = v;
  }
  if (pl->n_acc < 2) return NQ_ERR_UNSUPPORTED;
  // Short K (the head's dgrad: 9 taps x 16 channels): ~1000 MMA cycles per tile against an epilogue of ~3 chunks
  // x 150 dependent instructions per warp -- give the epilogue 12 of the 16 worker warps, the loaders 4.
  pl->n_epi = (d->ksize * d->ksize * C <= 1024 && pl->NT >= 48) ? 12 : 8;
  if (const char* e = getenv("NQ_TC_NEPI")) {  // tuning override
    const int v = atoi(e);
    if (v == 8 || v == 12) pl->n_epi = v;
  }
  // Weights that fit the ring whole (the head: 9 stages of 3 KB) are loaded once per CTA and stay resident.
  pl->resident = (!pl->cg2 && N <= pl->NT && pl->mt == 1 && !pl->bcat && d->ksize <= 3 && pl->KC == C && sbc == C && sbc <= 96 && (C / sbc) * d->ksize * d->ksize <= pl->n_bstages) ? 1 : 0;
  pl->tiles_x = (d->w + tile_w - 1) / tile_w;
  pl->tiles_y = (d->h + TILE_H - 1) / TILE_H;
  pl->tiles_n = (N + pl->NT - 1) / pl->NT;
  pl->ksplit = 1;
  pl->workspace_floats = 0;
  pl->total_tiles = pl->tiles_x * pl->tiles_y * d->n * pl->tiles_n;
  const int taps = d->ksize * d->ksize;
  // bytes of the packed weight buffer: full-width stages for all but the last N tile
  const long long stages = (long long)(C / sbc) * taps;
  const int last_nt = N - (pl->tiles_n - 1) * pl->NT;
  pl->wpk_bytes = stages * sbc * (pl->cg2 && pl->bcat ? 6 : 2 * b_planes) * ((long long)(pl->tiles_n - 1) * pl->NT + last_nt);
  // weight-stream sharing: 2 CTAs per cluster pack all 148 SMs (74 TPCs); every stage splits evenly
  // (nt * SBC * 2 * planes is a multiple of 512 bytes)
  pl->cluster = 2;
  // split-K (data gradient only): when the tiles cannot fill half the SMs, deal the activation units to several CTAs
  if (dir == 1 && !pl->resident) {
    const int ncb = (C + pl->KC - 1) / pl->KC;
    const int slots = pl->total_tiles;
    int ks = slots * 2 <= sm_count() ? sm_count() / slots : 1;
    if (ks > ncb) ks = ncb;
    if (const char* e = getenv("NQ_TC_KSPLIT")) {  // tuning override (0: off)
      const int v = atoi(e);
      if (v <= 1) ks = 1; else if (v <= ncb) ks = v;
    }
    if (ks > 1) {
      const int per = (ncb + ks - 1) / ks;
      pl->ksplit = (ncb + per - 1) / per;
      pl->workspace_floats = (long long)pl->ksplit * d->n * d->h * d->w * d->cin_p;
    }
  }
  return NQ_OK;
}

}  // namespace nq

using namespace nq;

extern "C" int nq_tc_plan_conv(const nq_conv_desc* d, int dir, int a_planes, int b_planes, nq_tc_plan* plan) {
  return fill_plan(d, dir, a_planes, b_planes, plan);
}

static int fill_pack(const nq_conv_desc* d, const nq_tc_plan* pl, const float* w_ref, int cin_src, const float* zero_point,
                     int zp_stride, void* wpk, TcPackParams& q, long long& blocks) {
  int st = check_conv_desc(d);
  if (st) return st;
  if (!pl || !w_ref || !wpk || cin_src < d->cin) return NQ_ERR_BAD_ARG;
  q.w = w_ref; q.zp = zero_point; q.zp_stride = zp_stride; q.out = reinterpret_cast<uint8_t*>(wpk);
  q.cout = d->cout; q.cin = d->cin; q.cin_src = cin_src; q.ks = d->ksize;
  q.rh = d->rh; q.rw = d->rw; q.c_grp = d->c_grp; q.cg = d->cg;
  q.dir = pl->dir; q.C = pl->C; q.N = pl->N; q.NT = pl->NT; q.KC = pl->KC; q.SBC = pl->SBC; q.b_planes = pl->b_planes; q.bcat = pl->bcat; q.cg2 = pl->cg2;
  const long long total = (long long)(pl->C / pl->SBC) * d->ksize * d->ksize * (pl->SBC / 8) * pl->NT * pl->tiles_n;
  blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  return NQ_OK;
}

extern "C" int nq_tc_pack_weight(const nq_conv_desc* d, const nq_tc_plan* pl, const float* w_ref, int cin_src,
                                 const float* zero_point, int zp_stride, void* wpk, void* stream) {
  TcPackParams q{};
  long long blocks = 0;
  const int st = fill_pack(d, pl, w_ref, cin_src, zero_point, zp_stride, wpk, q, blocks);
  if (st) return st;
  tc_pack_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(q);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

extern "C" int nq_tc_pack_multi(const nq_tc_pack_task* tasks, int n_tasks, void* stream) {
  if (!tasks || n_tasks <= 0) return NQ_ERR_BAD_ARG;
  for (int i0 = 0; i0 < n_tasks; i0 += NQ_MULTI_MAX) {
    TcPackMulti m{};
    m.n = n_tasks - i0 < NQ_MULTI_MAX ? n_tasks - i0 : NQ_MULTI_MAX;
    int blocks = 0;
    for (int i = 0; i < m.n; ++i) {
      const nq_tc_pack_task& t = tasks[i0 + i];
      if (!t.d) return NQ_ERR_BAD_ARG;
      long long nb = 0;
      if (t.wpk) {
        const int st = fill_pack(t.d, t.plan, t.w_ref, t.cin_src, t.zero_point, t.zp_stride, t.wpk, m.t[i], nb);
        if (st) return st;
      } else {
        const int st = check_conv_desc(t.d);
        if (st) return st;
      }
      if (t.scale_packed || t.bias_packed) {
        const nq_conv_desc* d = t.d;
        m.v[i] = TcPackVec{t.delta, t.bias_ref, t.scale_packed, t.bias_packed, t.d_stride, d->cout, d->rh, d->rw, d->c_grp, d->cg,
                           d->rh * d->rw * d->cg};
        if (nb < 1) nb = 1;
      }
      if (nb < 1) return NQ_ERR_BAD_ARG;  // a task with nothing to do
      // cap the share of one task so that 16 tasks stay within a few waves
      const long long cap = (long long)sm_count() * 4;
      if (nb > cap) nb = cap;
      m.blk_start[i] = blocks;
      blocks += (int)nb;
    }
    m.blk_start[m.n] = blocks;
    tc_pack_multi_kernel<<<blocks, 256, 0, as_stream(stream)>>>(m);
    NQ_LAUNCH_CHECK();
  }
  return NQ_OK;
}

extern "C" int nq_tc_pack_epilogue(const nq_conv_desc* d, const float* delta, int d_stride, const float* bias_ref,
                                   float* scale_packed, float* bias_packed, void* stream) {
  int st = check_conv_desc(d);
  if (st) return st;
  const int N = d->rh * d->rw * d->cg;
  TcPackVec v{delta, bias_ref, scale_packed, bias_packed, d_stride, d->cout, d->rh, d->rw, d->c_grp, d->cg, N};
  tc_pack_vec_kernel<<<(N + 255) / 256, 256, 0, as_stream(stream)>>>(v);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}

static int launch_tc(const nq_conv_desc* d, const nq_tc_plan* pl, TcParams& p, cudaStream_t s) {
  p.n = d->n; p.h = d->h; p.w = d->w; p.C = pl->C; p.ks = d->ksize; p.pad = d->ksize / 2;
  if (p.in_stride == 0) { p.in_stride = pl->C; p.c_valid = pl->C; }
  if (p.n_store == 0) p.n_store = pl->N;
  p.N = pl->N; p.NT = pl->NT; p.KC = pl->KC; p.SBC = pl->SBC; p.a_planes = pl->a_planes; p.b_planes = pl->b_planes;
  p.tiles_x = pl->tiles_x; p.tiles_y = pl->tiles_y; p.tiles_n = pl->tiles_n; p.total_tiles = pl->total_tiles;
  p.PW = pl->PW; p.PH = pl->PH; p.CGS = pl->CGS; p.mt = pl->mt; p.bcat = pl->bcat; p.n_epi = pl->n_epi;
  if (p.n_epi != 8 && p.n_epi != 12) return NQ_ERR_BAD_ARG;
  p.n_abuf = pl->n_abuf; p.n_acc = pl->n_acc; p.acc_stride = pl->acc_stride; p.sub_stride = pl->acc_stride / pl->mt;
  if (p.n_abuf < 2 || p.n_abuf > TC_MAX_RING || p.n_acc < 2 || p.n_acc > TC_MAX_RING || p.n_acc * p.acc_stride > 512 ||
      p.sub_stride < pl->NT * (pl->bcat ? 2 : 1))
    return NQ_ERR_BAD_ARG;
  if (p.mt < 1 || p.mt > 2 || (p.mt == 2 && (pl->NT > 128 || p.epi == 2))) return NQ_ERR_BAD_ARG;
  if (p.bcat && (p.epi == 2 || pl->a_planes != 2 || pl->b_planes != 2 || 2 * pl->NT * pl->mt > 256)) return NQ_ERR_BAD_ARG;
  if (pl->cg2 && (pl->resident || p.epi == 2 || (pl->NT & 15))) return NQ_ERR_BAD_ARG;
  p.a_plane_bytes = pl->a_plane_bytes; p.a_buf_bytes = pl->a_buf_bytes; p.b_stage_bytes = pl->b_stage_bytes;
  p.n_bstages = pl->n_bstages;
  p.gst = pl->gst < 1 ? 1 : pl->gst;
  if (p.gst > 1 && pl->resident) return NQ_ERR_BAD_ARG;
  p.b_slot_bytes = p.gst * pl->b_stage_bytes;
  p.epi_stage_off = TC_HDR_BYTES + pl->n_abuf * pl->a_buf_bytes + pl->n_bstages * p.b_slot_bytes;
  if (p.epi == 1 || p.epi == 3) { p.out_h = p.h / p.rh; p.out_w = p.w / p.rw; p.quad_stride = p.rh * p.rw * p.n_store; }
  else { p.out_h = p.h * p.rh; p.out_w = p.w * p.rw; p.quad_stride = 0; }
  if ((long long)p.n * p.out_h * p.out_w >= (1LL << 31) || (long long)p.n * p.h * p.w >= (1LL << 31)) return NQ_ERR_BAD_SHAPE;
  p.nsb_last = (pl->C - ((pl->C + pl->KC - 1) / pl->KC - 1) * pl->KC) / pl->SBC;
  p.resident = pl->resident;
  p.fd_tiles_x = make_fastdiv(pl->tiles_x); p.fd_tiles_y = make_fastdiv(pl->tiles_y);
  p.fd_rh = make_fastdiv(p.rh); p.fd_rw = make_fastdiv(p.rw); p.fd_PW = make_fastdiv(pl->PW);
  p.fd_npix = make_fastdiv(pl->PW * pl->PH); p.fd_cg = make_fastdiv(p.cg);
  void (*kern)(const TcParams) = pl->cg2   ? (p.bcat ? (p.mt == 2 ? conv_tc_kernel<2, 1, 0, 2> : conv_tc_kernel<1, 1, 0, 2>)
                                                      : (p.mt == 2 ? conv_tc_kernel<2, 0, 0, 2> : conv_tc_kernel<1, 0, 0, 2>))
                                 : p.mt == 2 ? (p.bcat ? conv_tc_kernel<2, 1, 0, 1> : conv_tc_kernel<2, 0, 0, 1>)
                                 : p.bcat    ? conv_tc_kernel<1, 1, 0, 1>
                                             : (p.resident ? conv_tc_kernel<1, 0, 1, 1> : conv_tc_kernel<1, 0, 0, 1>);
  NQ_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  // cluster size: CTAs working on neighbouring pixel tiles of the same N tile share the weight stream
  int cs = pl->cluster;
  if (p.resident) {
    if (pl->tiles_n != 1 || pl->mt != 1 || pl->bcat || d->ksize > 3 || pl->KC != pl->C || pl->SBC != pl->C || pl->SBC > 96 ||
        d->ksize * d->ksize > pl->n_bstages)
      return NQ_ERR_BAD_ARG;
    cs = 1;  // nothing left to share: every CTA loads its own copy once
  }
  p.tiles_m = pl->tiles_x * pl->tiles_y * d->n;
  if (cs < 1 || p.tiles_m < 2 * cs) cs = 1;
  if (pl->cg2) {  // the pair IS the cluster
    if (p.tiles_m < 4) return NQ_ERR_BAD_ARG;
    cs = 2;
  }
  p.cs = cs;
  p.tiles_m_pad = (p.tiles_m + cs - 1) / cs * cs;
  p.ncb = (pl->C + pl->KC - 1) / pl->KC;
  if (p.ksplit < 1) { p.ksplit = 1; p.cb_per_split = p.ncb; }
  p.total_groups = p.tiles_m_pad / cs * pl->tiles_n * p.ksplit;
  int n_clusters = sm_count() / cs;
  if (n_clusters > p.total_groups) n_clusters = p.total_groups;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(n_clusters * cs));
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = (size_t)pl->smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  static const int n_prod_env = getenv("NQ_TC_PROD") ? atoi(getenv("NQ_TC_PROD")) : 0;  // tuning override
  p.n_prod = n_prod_env >= 1 && n_prod_env <= 32 ? n_prod_env : 4;
  if (p.n_prod > pl->n_bstages) p.n_prod = pl->n_bstages;
  static const int skip_flags = getenv("NQ_TC_SKIP") ? atoi(getenv("NQ_TC_SKIP")) : 0;
  p.skip = skip_flags;
  static const bool dbg_on = getenv("NQ_TC_DBG") != nullptr;
  static long long* dbg_buf = nullptr;
  if (dbg_on) {  // debugging aid: synchronises after every launch and prints the SM clock the kernel saw
    if (!dbg_buf) NQ_CUDA_CHECK(cudaMalloc(&dbg_buf, 8 * sizeof(long long)));
    NQ_CUDA_CHECK(cudaMemsetAsync(dbg_buf, 0, 8 * sizeof(long long), s));
    p.dbg = dbg_buf;
  }
  NQ_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
  if (dbg_on) {
    long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    NQ_CUDA_CHECK(cudaStreamSynchronize(s));
    NQ_CUDA_CHECK(cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[NQ_TC_DBG] epi=%d hw=%dx%d C=%d N=%d ks=%d mt=%d bcat=%d: %lld cycles, %lld ns, SM clock %.0f MHz\n", p.epi, p.h, p.w,
            p.C, p.N, p.ks, p.mt, p.bcat, h[0], h[1], h[1] > 0 ? (double)h[0] / (double)h[1] * 1e3 : 0.0);
  }
  return NQ_OK;
}

extern "C" int nq_tc_conv_fwd(const nq_conv_desc* d, const nq_tc_plan* pl, const void* x_split, const void* wpk,
                              const float* scale_packed, const float* bias_packed, float* z, void* y_split,
                              void* stream) {
  int st = check_conv_desc(d);
  if (st) return st;
  if (!pl || pl->dir != 0 || !x_split || !wpk || (!y_split && !z)) return NQ_ERR_BAD_ARG;
  TcParams p{};
  p.in = reinterpret_cast<const uint8_t*>(x_split);
  p.in_plane_bytes = (size_t)d->n * d->h * d->w * d->cin_p * 2;
  p.in_stride = d->cin_p; p.c_valid = d->cin_p;
  p.wpk = reinterpret_cast<const uint8_t*>(wpk); p.scale = scale_packed; p.bias = bias_packed;
  p.zprev = nullptr; p.out_z = z; p.out_y = reinterpret_cast<uint8_t*>(y_split); p.epi = 0;
  p.out_plane_bytes = (size_t)d->n * d->h * d->rh * d->w * d->rw * d->cg * 2;
  p.n_store = d->rh * d->rw * d->cg;
  p.rh = d->rh; p.rw = d->rw; p.cg = d->cg; p.act = d->act;
  return launch_tc(d, pl, p, as_stream(stream));
}

extern "C" int nq_tc_head_fwd_loss(const nq_conv_desc* d, const nq_tc_plan* pl, const void* x_split, const void* wpk,
                                   const float* scale_packed, const float* bias_packed, int out_bias,
                                   const float* target, float p_norm, float mean_pixels, float* img, float* loss_sum,
                                   void* dz_head_split, void* stream) {
  int st = check_conv_desc(d);
  if (st) return st;
  if (!pl || pl->dir != 0 || !x_split || !wpk || !bias_packed) return NQ_ERR_BAD_ARG;
  if (d->ksize != 3 || d->rh != 1 || d->rw != 1 || d->cout != 3 || d->cg != 4) return NQ_ERR_BAD_SHAPE;
  if (out_bias != 0 && out_bias != 1) return NQ_ERR_UNSUPPORTED;
  if (target && (!(p_norm > 0.f) || !(mean_pixels > 0.f))) return NQ_ERR_BAD_ARG;
  if (!target && !img) return NQ_ERR_BAD_ARG;
  TcParams p{};
  p.in = reinterpret_cast<const uint8_t*>(x_split);
  p.in_plane_bytes = (size_t)d->n * d->h * d->w * d->cin_p * 2;
  p.in_stride = d->cin_p; p.c_valid = d->cin_p;
  p.wpk = reinterpret_cast<const uint8_t*>(wpk); p.scale = scale_packed; p.bias = bias_packed;
  p.epi = 2; p.n_store = 4;
  p.rh = 1; p.rw = 1; p.cg = 4; p.act = 0;
  p.head_target = target; p.head_img = img; p.head_loss = target ? loss_sum : nullptr;
  p.head_dz = target ? reinterpret_cast<uint8_t*>(dz_head_split) : nullptr;
  p.out_plane_bytes = (size_t)d->n * d->h * d->w * 8 * 2;  // plane stride of dz_head_split
  p.head_p = p_norm; p.head_inv_mean = target ? 1.0f / mean_pixels : 0.f; p.head_out_bias = out_bias;
  return launch_tc(d, pl, p, as_stream(stream));
}

// ------------------------------------------------------------------------------------------------
// Split-K data gradient.  The deep stages have few pixels (HNeRV-3M stage 3: 40 x 80 x 2 = 50 tiles for 148 SMs,
// stage 2: 6 tiles) and a long K (9-25 taps x 1024 channels): every CTA would stream the whole weight tensor.
// With plan->ksplit > 1 the activation units (channel blocks) of a tile are dealt to ksplit CTAs that write raw fp32
// partial sums; this kernel adds them in a fixed order and applies the epilogue (activation derivative, un-shuffle,
// split-bf16 store).
// ------------------------------------------------------------------------------------------------
namespace nq {
__global__ void __launch_bounds__(256) dgrad_finish_kernel(const float* __restrict__ part, int ksplit, size_t part_stride,
                                                           const float* __restrict__ zprev, int act, int n, int h, int w,
                                                           int n_store, int rh, int rw, uint8_t* __restrict__ out,
                                                           size_t out_plane_bytes) {
  const int c4n = n_store >> 2;
  const int64_t total4 = (int64_t)n * h * w * c4n;
  const int hq = h / rh, wq = w / rw;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total4; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = e / c4n;
    const int c = (int)(e - pix * c4n) * 4;
    const float4* p4 = reinterpret_cast<const float4*>(part + pix * n_store + c);
    float4 r = *p4;
    for (int k = 1; k < ksplit; ++k) {
      const float4 v = *reinterpret_cast<const float4*>(part + (size_t)k * part_stride + pix * n_store + c);
      r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w;
    }
    if (zprev != nullptr && act != 0) {
      const float4 z = __ldg(reinterpret_cast<const float4*>(zprev + pix * n_store + c));
      if (act == 1) {
        r.x *= gelu_grad_fast(z.x); r.y *= gelu_grad_fast(z.y); r.z *= gelu_grad_fast(z.z); r.w *= gelu_grad_fast(z.w);
      } else {
        r.x *= z.x; r.y *= z.y; r.z *= z.z; r.w *= z.w;
      }
    }
    const int x = (int)(pix % w);
    const int64_t t = pix / w;
    const int y = (int)(t % h), img = (int)(t / h);
    const int qh = y / rh, si = y - qh * rh, qw = x / rw, sj = x - qw * rw;
    const size_t o = ((size_t)(img * hq + qh) * wq + qw) * ((size_t)rh * rw * n_store) + (size_t)(si * rw + sj) * n_store + c;
    store_split4(out, out_plane_bytes, o, r);
  }
}
}  // namespace nq

extern "C" int nq_tc_conv_dgrad(const nq_conv_desc* d, const nq_tc_plan* pl, const void* dz_split, const void* wpk_t,
                                const float* z_prev, int prev_rh, int prev_rw, int prev_act, void* dz_prev_split,
                                float* workspace, int64_t workspace_floats, void* stream) {
  int st = check_conv_desc(d);
  if (st) return st;
  if (!pl || pl->dir != 1 || !dz_split || !wpk_t || !dz_prev_split || prev_rh <= 0 || prev_rw <= 0) return NQ_ERR_BAD_ARG;
  if (d->h % prev_rh || d->w % prev_rw) return NQ_ERR_BAD_SHAPE;
  if (prev_act < 0 || prev_act > 2) return NQ_ERR_BAD_ARG;
  TcParams p{};
  const int nout_p = d->rh * d->rw * d->cg;
  const int dz_ch = (nout_p + 7) / 8 * 8;  // channels per pixel as stored (the head's 4 are stored as 8)
  p.in = reinterpret_cast<const uint8_t*>(dz_split);
  p.in_plane_bytes = (size_t)d->n * d->h * d->w * dz_ch * 2;
  p.in_stride = dz_ch; p.c_valid = dz_ch;
  p.wpk = reinterpret_cast<const uint8_t*>(wpk_t); p.scale = nullptr; p.bias = nullptr;
  p.zprev = z_prev; p.out_z = nullptr; p.out_y = reinterpret_cast<uint8_t*>(dz_prev_split); p.epi = 1;
  p.out_plane_bytes = (size_t)d->n * d->h * d->w * d->cin_p * 2;
  p.rh = prev_rh; p.rw = prev_rw; p.cg = d->cin_p; p.act = prev_act;
  p.n_store = d->cin_p;
  if (pl->ksplit > 1) {  // split-K: raw partial sums, then one finishing pass
    if (!workspace || workspace_floats < pl->workspace_floats) return NQ_ERR_WORKSPACE;
    const size_t per = (size_t)d->n * d->h * d->w * d->cin_p;
    p.epi = 3; p.zprev = nullptr; p.out_y = nullptr;
    p.part = workspace; p.part_stride = per;
    p.ksplit = pl->ksplit;
    p.cb_per_split = ((pl->C + pl->KC - 1) / pl->KC + pl->ksplit - 1) / pl->ksplit;
    st = launch_tc(d, pl, p, as_stream(stream));
    if (st) return st;
    const int64_t total4 = (int64_t)per / 4;
    int64_t blocks = (total4 + 255) / 256;
    if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
    dgrad_finish_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(workspace, pl->ksplit, per, z_prev, prev_act, d->n, d->h, d->w,
                                                                        d->cin_p, prev_rh, prev_rw,
                                                                        reinterpret_cast<uint8_t*>(dz_prev_split),
                                                                        (size_t)d->n * d->h * d->w * d->cin_p * 2);
    NQ_LAUNCH_CHECK();
    return NQ_OK;
  }
  return launch_tc(d, pl, p, as_stream(stream));
}

// ------------------------------------------------------------------------------------------------
// split-bf16 edges: fp32 <-> (hi, lo) planes
// ------------------------------------------------------------------------------------------------
namespace nq {
__global__ void __launch_bounds__(256) nchw_to_split_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst,
                                                            int n, int c, int h, int w, int c_p) {
  const int64_t total = (int64_t)n * h * w * c_p;
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(dst);
  __nv_bfloat16* lo = hi + total;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(e % c_p);
    const int64_t pix = e / c_p;
    const int x = (int)(pix % w), y = (int)((pix / w) % h), b = (int)(pix / ((int64_t)w * h));
    const float v = ch < c ? src[(((int64_t)b * c + ch) * h + y) * w + x] : 0.f;
    const __nv_bfloat16 hv = __float2bfloat16_rn(v);
    hi[e] = hv;
    lo[e] = __float2bfloat16_rn(v - __bfloat162float(hv));
  }
}
__global__ void __launch_bounds__(256) split_to_nchw_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int n,
                                                            int c, int h, int w, int c_p) {
  const int64_t total = (int64_t)n * c * h * w;
  const __nv_bfloat16* hi = reinterpret_cast<const __nv_bfloat16*>(src);
  const __nv_bfloat16* lo = hi + (int64_t)n * h * w * c_p;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(e % w), y = (int)((e / w) % h);
    const int ch = (int)((e / ((int64_t)w * h)) % c), b = (int)(e / ((int64_t)w * h * c));
    const int64_t o = (((int64_t)b * h + y) * w + x) * c_p + ch;
    dst[e] = __bfloat162float(hi[o]) + __bfloat162float(lo[o]);
  }
}
__global__ void __launch_bounds__(256) f32_to_split_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int64_t numel) {
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(dst);
  __nv_bfloat16* lo = hi + numel;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < numel; e += (int64_t)gridDim.x * blockDim.x) {
    const float v = src[e];
    const __nv_bfloat16 hv = __float2bfloat16_rn(v);
    hi[e] = hv;
    lo[e] = __float2bfloat16_rn(v - __bfloat162float(hv));
  }
}
__global__ void __launch_bounds__(256) split_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int64_t numel) {
  const __nv_bfloat16* hi = reinterpret_cast<const __nv_bfloat16*>(src);
  const __nv_bfloat16* lo = hi + numel;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < numel; e += (int64_t)gridDim.x * blockDim.x)
    dst[e] = __bfloat162float(hi[e]) + __bfloat162float(lo[e]);
}
static inline unsigned split_grid(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}
}  // namespace nq

extern "C" int nq_nchw_to_split(const float* src, void* dst, int n, int c, int h, int w, int c_p, void* stream) {
  if (!src || !dst || n <= 0 || c <= 0 || h <= 0 || w <= 0 || c_p < c) return NQ_ERR_BAD_ARG;
  nchw_to_split_kernel<<<split_grid((int64_t)n * h * w * c_p), 256, 0, as_stream(stream)>>>(
      src, reinterpret_cast<uint8_t*>(dst), n, c, h, w, c_p);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}
extern "C" int nq_split_to_nchw(const void* src, float* dst, int n, int c, int h, int w, int c_p, void* stream) {
  if (!src || !dst || n <= 0 || c <= 0 || h <= 0 || w <= 0 || c_p < c) return NQ_ERR_BAD_ARG;
  split_to_nchw_kernel<<<split_grid((int64_t)n * c * h * w), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint8_t*>(src), dst, n, c, h, w, c_p);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}
extern "C" int nq_f32_to_split(const float* src, void* dst, int64_t numel, void* stream) {
  if (!src || !dst || numel <= 0) return NQ_ERR_BAD_ARG;
  f32_to_split_kernel<<<split_grid(numel), 256, 0, as_stream(stream)>>>(src, reinterpret_cast<uint8_t*>(dst), numel);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}
extern "C" int nq_split_to_f32(const void* src, float* dst, int64_t numel, void* stream) {
  if (!src || !dst || numel <= 0) return NQ_ERR_BAD_ARG;
  split_to_f32_kernel<<<split_grid(numel), 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint8_t*>(src), dst, numel);
  NQ_LAUNCH_CHECK();
  return NQ_OK;
}
