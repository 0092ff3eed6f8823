// Micro-benchmark 2: what bounds tcgen05.mma (SS mode) at the shapes the convolutions issue.
//   * cta_group::1, M = 128, N in {48 .. 256}: pipe floor N/2 cycles against the shared-memory operand reads
//     (128 x 32 B of A + N x 32 B of B per instruction)
//   * cta_group::2, M = 256 (two CTAs, each holding its 128 A rows and HALF of B)
//   * kind::tf32 (K = 8 per instruction)
//   * the SM clock under this load: clock64 against globaltimer
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench2 umma_bench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct Cfg { int N, cg, tf32 /* MODE */, reps, a_bytes_step, b_bytes_step; int data; /* 0: zeros, 1: random bf16 in [-2, 2) */ };

__device__ __forceinline__ uint64_t mkdesc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint64_t gtimer() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// MODE 0: bf16, every MMA its own A and B   1: tf32   2: bf16, pairs of MMAs sharing A through the A collector (fill, lastuse)
//      3: bf16 weight-stationary (.ws), pairs sharing B through collector b0 (fill, lastuse)
template <int CG, int MODE>
__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += 128) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (c.data) {  // pseudo-random bf16 pairs: sign, exponent 126..128, random mantissa
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
      uint32_t w[4];
      for (int k = 0; k < 4; ++k) {
        h = h * 1664525u + 1013904223u;
        const uint32_t lo = (h & 0x807Fu) | ((126u + ((h >> 8) & 1u)) << 7);
        const uint32_t hi = ((h >> 16) & 0x807Fu) | ((126u + ((h >> 24) & 1u)) << 7);
        w[k] = lo | (hi << 16);
      }
      v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    reinterpret_cast<uint4*>(smem)[i] = v;
  }
  const uint32_t barA = (uint32_t)__cvta_generic_to_shared(&bar);
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barA));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tslot)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tslot)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;");
    asm volatile("barrier.cluster.wait.acquire.aligned;");
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tslot;
  if (threadIdx.x < 32 && rank == 0) {
    uint32_t leader = 0;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(leader));
    const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(smem), b0 = a0 + 100 * 1024;
    const int M = 128 * CG;
    const uint32_t fmt = MODE == 1 ? ((2u << 7) | (2u << 10)) : ((1u << 7) | (1u << 10));
    const uint32_t idesc = (1u << 4) | fmt | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t nb = CG == 2 ? c.N / 2 : c.N;  // B rows held by THIS CTA
    uint64_t ad[4], bd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // MODE 2: MMAs 2j and 2j+1 share A;  MODE 3: they share B
      ad[k] = mkdesc(a0 + (MODE == 2 ? (k >> 1) : k) * c.a_bytes_step, 2048, 128);
      bd[k] = mkdesc(b0 + (MODE == 3 ? (k >> 1) : k) * c.b_bytes_step, nb * 16, 128);
    }
    const long long t0 = clock64();
    const uint64_t g0 = gtimer();
    for (int r = 0; r < c.reps; r += 4) {
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t d = tm + (MODE >= 2 ? (k & 1) * c.N : 0);
#define MMA(txt) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n" txt " [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(ad[k]), "l"(bd[k]), "r"(idesc), "r"(1u))
          if (MODE == 0 && CG == 1) MMA("tcgen05.mma.cta_group::1.kind::f16");
          if (MODE == 0 && CG == 2) MMA("tcgen05.mma.cta_group::2.kind::f16");
          if (MODE == 1 && CG == 1) MMA("tcgen05.mma.cta_group::1.kind::tf32");
          if (MODE == 1 && CG == 2) MMA("tcgen05.mma.cta_group::2.kind::tf32");
          if (MODE == 2 && CG == 1 && !(k & 1)) MMA("tcgen05.mma.cta_group::1.kind::f16.collector::a::fill");
          if (MODE == 2 && CG == 1 && (k & 1)) MMA("tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse");
          if (MODE == 2 && CG == 2 && !(k & 1)) MMA("tcgen05.mma.cta_group::2.kind::f16.collector::a::fill");
          if (MODE == 2 && CG == 2 && (k & 1)) MMA("tcgen05.mma.cta_group::2.kind::f16.collector::a::lastuse");
          if (MODE == 3 && !(k & 1)) MMA("tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill");
          if (MODE == 3 && (k & 1)) MMA("tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse");
        }
      }
    }
    if (leader) {
      if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barA));
      else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barA));
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"(barA));
    const long long t1 = clock64();
    const uint64_t g1 = gtimer();
    if (leader) {
      out[2 * (blockIdx.x / CG)] = t1 - t0;
      out[2 * (blockIdx.x / CG) + 1] = (long long)(g1 - g0);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;");
    asm volatile("barrier.cluster.wait.acquire.aligned;");
  }
  if (threadIdx.x < 32) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm));
  }
}

typedef void (*kern_t)(Cfg, long long*);
static kern_t pick(int cg, int mode) {
  if (cg == 1) return mode == 0 ? bench<1, 0> : mode == 1 ? bench<1, 1> : mode == 2 ? bench<1, 2> : bench<1, 3>;
  return mode == 0 ? bench<2, 0> : mode == 1 ? bench<2, 1> : bench<2, 2>;
}

int main(int argc, char** argv) {
  long long* d;
  cudaMalloc(&d, 2 * 148 * sizeof(long long));
  const int R = argc > 1 ? atoi(argv[1]) : 100000;
  struct Named { const char* name; Cfg c; };
  // Cfg: N, cta_group, mode, reps, A bytes between the 4 operand tiles, B bytes between them
  Named cfgs[] = {
      {"cg1 bf16 N48", {48, 1, 0, R, 4096, 1536, 0}},   {"cg1 bf16 N64", {64, 1, 0, R, 4096, 2048, 0}},
      {"cg1 bf16 N96", {96, 1, 0, R, 4096, 3072, 0}},   {"cg1 bf16 N128", {128, 1, 0, R, 4096, 4096, 0}},
      {"cg1 bf16 N160", {160, 1, 0, R, 4096, 5120, 0}}, {"cg1 bf16 N192", {192, 1, 0, R, 4096, 6144, 0}},
      {"cg1 bf16 N256", {256, 1, 0, R, 4096, 8192, 0}}, {"cg1 tf32 N48", {48, 1, 1, R, 4096, 1536, 0}},
      {"cg1 tf32 N160", {160, 1, 1, R, 4096, 5120, 0}}, {"cg1 tf32 N256", {256, 1, 1, R, 4096, 8192, 0}},
      {"cg1 bf16 N48 A-collector pairs", {48, 1, 2, R, 4096, 1536, 0}},
      {"cg1 bf16 N96 A-collector pairs", {96, 1, 2, R, 4096, 3072, 0}},
      {"cg1 bf16 N160 A-collector pairs", {160, 1, 2, R, 4096, 5120, 0}},
      {"cg1 ws bf16 N64 B-collector pairs", {64, 1, 3, R, 4096, 2048, 0}},
      {"cg1 ws bf16 N128 B-collector pairs", {128, 1, 3, R, 4096, 4096, 0}},
      {"cg1 ws bf16 N256 B-collector pairs", {256, 1, 3, R, 4096, 8192, 0}},
      {"cg2 bf16 N48 (M256)", {48, 2, 0, R, 4096, 768, 0}},    {"cg2 bf16 N96 (M256)", {96, 2, 0, R, 4096, 1536, 0}},
      {"cg2 bf16 N128 (M256)", {128, 2, 0, R, 4096, 2048, 0}}, {"cg2 bf16 N160 (M256)", {160, 2, 0, R, 4096, 2560, 0}},
      {"cg2 bf16 N192 (M256)", {192, 2, 0, R, 4096, 3072, 0}}, {"cg2 bf16 N256 (M256)", {256, 2, 0, R, 4096, 4096, 0}},
      {"cg2 tf32 N160 (M256)", {160, 2, 1, R, 4096, 2560, 0}},
      {"cg2 bf16 N48 A-collector pairs", {48, 2, 2, R, 4096, 768, 0}},
      {"cg2 bf16 N96 A-collector pairs", {96, 2, 2, R, 4096, 1536, 0}},
  };
  for (int data = 0; data < 2; ++data)
  for (auto& nc : cfgs) {
    nc.c.data = data;
    kern_t k = pick(nc.c.cg, nc.c.tf32);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaMemset(d, 0, 2 * 148 * sizeof(long long));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nc.c.cg; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, nc.c, d);
    if (e != cudaSuccess) { printf("%s: launch %s\n", nc.name, cudaGetErrorString(e)); return 1; }
    e = cudaDeviceSynchronize();
    long long h[2 * 148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0, ns = 1;
    const int n = 148 / nc.c.cg;
    for (int i = 0; i < n; ++i) if (h[2 * i] > mx) { mx = h[2 * i]; ns = h[2 * i + 1]; }
    const double cyc = (double)mx / nc.c.reps;
    const double flops = 2.0 * 128 * nc.c.cg * nc.c.N * (nc.c.tf32 == 1 ? 8 : 16) * (double)nc.c.reps * n;
    printf("%s %-36s %7.1f cyc/MMA (pipe floor %5.1f)  SM clock %6.0f MHz  %7.1f TFLOP/s  %s\n", data ? "random" : "zeros ", nc.name, cyc, 128.0 * nc.c.N / 256.0,
           (double)mx / (double)ns * 1e3, flops / ((double)ns * 1e-9) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
