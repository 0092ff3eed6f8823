// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16, M=128) for different shared-memory operand
// layouts.  Data are zeros; only the issue/execute rate matters.   nvcc -arch=sm_100a -o umma_bench umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct Cfg { int N; int a_layout, b_layout; int a_lbo, a_sbo, b_lbo, b_sbo; int a_major, b_major; int a_step, b_step; int reps; int a_off; };

__device__ __forceinline__ uint64_t mkdesc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
         ((uint64_t)((addr >> 7) & 7) << 49) * (layout != 0 ? 1 : 0) | ((uint64_t)layout << 61);
}

__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  const uint32_t barA = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barA));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tslot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tslot;
  if (threadIdx.x == 0) {
    const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(smem) + c.a_off, b0 = (uint32_t)__cvta_generic_to_shared(smem) + 100 * 1024;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)c.a_major << 15) | ((uint32_t)c.b_major << 16) |
                           ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    long long t0 = clock64();
    for (int r = 0; r < c.reps; ++r) {
      const uint32_t k = r & 3;
      const uint64_t ad = mkdesc(a0 + k * c.a_step, c.a_lbo, c.a_sbo, c.a_layout);
      const uint64_t bd = mkdesc(b0 + k * c.b_step, c.b_lbo, c.b_sbo, c.b_layout);
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tm),
                   "l"(ad), "l"(bd), "r"(idesc), "r"(1u));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barA));
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"(barA));
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Named { const char* name; Cfg c; };
  const int R = 4000;
  Named cfgs[] = {
      // name, {N, a_layout, b_layout, a_lbo, a_sbo, b_lbo, b_sbo, a_major, b_major, a_step, b_step, reps, a_off}
      {"K-major none  N160 halo(sbo192,lbo3856)", {160, 0, 0, 3856, 192, 2560, 128, 0, 0, 7712, 5120, R, 0}},
      {"K-major none  N160 dense(sbo128,lbo2048)", {160, 0, 0, 2048, 128, 2560, 128, 0, 0, 4096, 5120, R, 0}},
      {"K-major none  N256 dense", {256, 0, 0, 2048, 128, 4096, 128, 0, 0, 4096, 8192, R, 0}},
      {"K-major none  N64  dense", {64, 0, 0, 2048, 128, 1024, 128, 0, 0, 4096, 2048, R, 0}},
      {"K-major none  N48  dense", {48, 0, 0, 2048, 128, 768, 128, 0, 0, 4096, 1536, R, 0}},
      {"K-major none  N160 halo +16B offset", {160, 0, 0, 3856, 192, 2560, 128, 0, 0, 7712, 5120, R, 16}},
      {"K-major SW128 N160", {160, 2, 2, 16, 1024, 16, 1024, 0, 0, 32, 32, R, 0}},
      {"K-major SW128 N256", {256, 2, 2, 16, 1024, 16, 1024, 0, 0, 32, 32, R, 0}},
      {"K-major SW128 N64", {64, 2, 2, 16, 1024, 16, 1024, 0, 0, 32, 32, R, 0}},
      {"K-major SW128 N160 A rows shifted 128B (base_offset)", {160, 2, 2, 16, 2048, 16, 1024, 0, 0, 32, 32, R, 128}},
      {"K-major SW64  N160", {160, 4, 4, 16, 512, 16, 512, 0, 0, 32, 32, R, 0}},
      {"K-major SW32  N160", {160, 6, 6, 16, 256, 16, 256, 0, 0, 4096, 5120, R, 0}},
      {"A none / B SW128 N160", {160, 0, 2, 3856, 192, 16, 1024, 0, 0, 7712, 32, R, 0}},
      {"A SW128 / B none N160", {160, 2, 0, 16, 1024, 2560, 128, 0, 0, 32, 5120, R, 0}},
      {"MN-major none N160 (wgrad)", {160, 0, 0, 128, 1040, 128, 1040, 1, 1, 256, 256, R, 0}},
      {"MN-major SW128 N160", {160, 2, 2, 1024, 2048, 1024, 2048, 1, 1, 4096, 4096, R, 0}},
  };
  for (auto& nc : cfgs) {
    bench<<<148, 128, 200 * 1024>>>(nc.c, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-58s  %7.1f cyc/MMA  (ideal %5.1f)  %s\n", nc.name, (double)mx / nc.c.reps, 128.0 * nc.c.N / 256.0,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
