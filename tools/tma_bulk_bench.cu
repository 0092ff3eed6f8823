// Micro-benchmark 4: how fast one CTA per SM can stream an L2-resident buffer into shared memory with cp.async.bulk
// (the weight-stage stream of conv_tc_kernel): bytes per SM clock per SM for different copy sizes, ring depths, issuing
// threads, and with 2-CTA multicast.  All 148 SMs stream the same `src_bytes` buffer round and round.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bulk_bench tma_bulk_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>

struct Cfg { int copy_bytes, ring, n_copies, src_bytes, producers, mc; int wait; /* 0 try_wait, 1 test_wait spin, 2 try_wait with a 32 ns hint */
             int lanes; /* 1: the producers are lanes of ONE warp */ int tensor; /* 1: cp.async.bulk.tensor.2d through a tensor map (rows of 128 bytes) */ };

__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar),
               "r"(parity)
               : "memory");
}

__device__ __forceinline__ void bar_spin(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar),
               "r"(parity)
               : "memory");
}
__device__ __forceinline__ void bar_wait_hint(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar),
               "r"(parity), "r"(32u)
               : "memory");
}
__device__ __forceinline__ void bar_any(int mode, uint32_t bar, uint32_t parity) {
  if (mode == 1) bar_spin(bar, parity); else if (mode == 2) bar_wait_hint(bar, parity); else bar_wait(bar, parity);
}

__global__ void __launch_bounds__(128, 1) bench(Cfg c, const uint8_t* src, long long* out, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[32];
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
  const uint32_t buf0 = (uint32_t)__cvta_generic_to_shared(smem);
  uint32_t rank = 0;
  if (c.mc > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    for (int i = 0; i < c.ring; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + i * 8));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (c.mc > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;");
    asm volatile("barrier.cluster.wait.acquire.aligned;");
  }
  // producer threads: thread p * 32 handles copies p, p + producers, ... (each its own slots)
  const int pid = c.lanes ? (int)threadIdx.x : (int)(threadIdx.x >> 5);
  if ((c.lanes ? threadIdx.x < 32 : (threadIdx.x & 31) == 0) && pid < c.producers) {
    const long long t0 = clock64();
    uint32_t off = (uint32_t)((blockIdx.x * 7919u * 1024u + pid * c.copy_bytes) % (uint32_t)c.src_bytes);
    for (int i = pid; i < c.n_copies; i += c.producers) {
      const int slot = i % c.ring;
      const uint32_t round = (uint32_t)(i / c.ring);
      if (round > 0) bar_any(c.wait, bar0 + slot * 8, (round - 1) & 1);  // the slot's previous copy has landed ("consumed" at once)
      const uint32_t bar = bar0 + slot * 8, dst = buf0 + slot * c.copy_bytes;
      if (off + (uint32_t)c.copy_bytes > (uint32_t)c.src_bytes) off = 0;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(c.copy_bytes) : "memory");
      if (c.tensor) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                     "l"(&tmap), "r"(0), "r"((int)(off >> 7)), "r"(bar)
                     : "memory");
      } else if (c.mc == 1) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(src + off), "r"(c.copy_bytes), "r"(bar)
                     : "memory");
      } else {  // each CTA fetches half and multicasts it to both
        const uint32_t part = c.copy_bytes / 2;
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                dst + rank * part),
            "l"(src + off + rank * part), "r"(part), "r"(bar), "h"((uint16_t)3)
            : "memory");
      }
      off += c.copy_bytes * c.producers;
    }
    // drain: every slot's last copy
    for (int s = 0; s < c.ring; ++s) {
      int last = -1;
      for (int i = pid; i < c.n_copies; i += c.producers) if (i % c.ring == s) last = i;
      if (last >= 0) bar_any(c.wait, bar0 + s * 8, (uint32_t)(last / c.ring) & 1);
    }
    if (pid == 0) out[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (c.mc > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;");
    asm volatile("barrier.cluster.wait.acquire.aligned;");
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  uint8_t* src;
  cudaMalloc(&src, 8 << 20);
  cudaMemset(src, 1, 8 << 20);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Named { const char* name; Cfg c; };
  // copy bytes, ring, copies, source bytes (L2-resident), producer threads, multicast cluster size
  Named cfgs[] = {
      {"3 KB x16, 1 thread", {3072, 16, 4000, 768 * 1024, 1, 1, 0, 0, 0}},
      {"6 KB x16, 1 thread", {6144, 16, 4000, 768 * 1024, 1, 1, 0, 0, 0}},
      {"15 KB x5, 1 thread", {15360, 5, 2000, 384 * 1024, 1, 1, 0, 0, 0}},
      {"30 KB x4, 1 thread", {30720, 4, 1000, 768 * 1024, 1, 1, 0, 0, 0}},
      {"6 KB x16, 4 warps", {6144, 16, 4000, 768 * 1024, 4, 1, 0, 0, 0}},
      {"15 KB x8, 4 warps", {15360, 8, 2000, 384 * 1024, 4, 1, 0, 0, 0}},
      {"3 KB x16, 4 lanes of a warp", {3072, 16, 4000, 768 * 1024, 4, 1, 0, 1, 0}},
      {"6 KB x16, 4 lanes of a warp", {6144, 16, 4000, 768 * 1024, 4, 1, 0, 1, 0}},
      {"6 KB x16, 16 lanes of a warp", {6144, 16, 4000, 768 * 1024, 16, 1, 0, 1, 0}},
      {"15 KB x8, 4 lanes of a warp", {15360, 8, 2000, 384 * 1024, 4, 1, 0, 1, 0}},
      {"15 KB x8, 8 lanes of a warp", {15360, 8, 2000, 384 * 1024, 8, 1, 0, 1, 0}},
      {"3 KB x16, tensor map, 1 thread", {3072, 16, 4000, 768 * 1024, 1, 1, 0, 0, 1}},
      {"6 KB x16, tensor map, 1 thread", {6144, 16, 4000, 768 * 1024, 1, 1, 0, 0, 1}},
      {"15 KB x5, tensor map, 1 thread", {15360, 5, 2000, 384 * 1024, 1, 1, 0, 0, 1}},
      {"15 KB x8, tensor map, 1 thread", {15360, 8, 2000, 384 * 1024, 1, 1, 0, 0, 1}},
      {"30 KB x4, tensor map, 1 thread", {30720, 4, 1000, 768 * 1024, 1, 1, 0, 0, 1}},
      {"15 KB x8, tensor map, 4 lanes", {15360, 8, 2000, 384 * 1024, 4, 1, 0, 1, 1}},
  };
  for (auto& nc : cfgs) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nc.c.mc; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CUtensorMap tmap;
    {
      cuuint64_t gdim[2] = {128, (cuuint64_t)(8 << 20) / 128};
      cuuint64_t gstride[1] = {128};
      cuuint32_t box[2] = {128, (cuuint32_t)nc.c.copy_bytes / 128};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = cuTensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, src, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, bench, nc.c, (const uint8_t*)src, d, tmap);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) if (h[i] > mx) mx = h[i];
    printf("%-34s %6.1f B/clk/SM landed  (%7.0f cycles per copy)  %s\n", nc.name, (double)nc.c.copy_bytes * nc.c.n_copies / (double)mx,
           (double)mx / nc.c.n_copies, e == cudaSuccess ? "" : cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
