// Micro-benchmark 3: what one issuing thread can sustain.  The MMA itself is made cheap (M128 N32: shared-memory
// operand floor 40 cycles) so that the instruction stream around it shows.  Styles:
//   0  descriptors loop-invariant, 4 MMAs per iteration (the rate the hardware takes MMAs from one thread)
//   1  rolled loop, 1 MMA per iteration, descriptor low words advanced by dependent adds, mov.b64 {lo, hi} per MMA
//   2  rolled loop, 2 MMAs per iteration (second one on another A plane), as conv_tc_kernel's forward loop
//   3  as 1, descriptors kept as 64-bit values advanced by 64-bit adds
//   4  style 0 + every 8 MMAs one try_wait on a completed mbarrier and one tcgen05.commit (a weight stage's bookkeeping)
//   5  style 2 + the same bookkeeping every 3 iterations (the forward stage loop as shipped in round 1)
//   6  as 2, but all descriptors of 4 iterations computed first (independent adds), then 8 MMAs back to back
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_issue_bench umma_issue_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_w(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}" ::"r"(d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_l(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d),
               "l"(da), "l"(db), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar),
               "r"(parity)
               : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

template <int STYLE>
__global__ void __launch_bounds__(128, 1) bench(int reps, int n_ring, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  const uint32_t bar_done = (uint32_t)__cvta_generic_to_shared(&bars[0]);
  const uint32_t bar_ready = (uint32_t)__cvta_generic_to_shared(&bars[1]);   // never armed: waiting for parity 1 succeeds at once
  const uint32_t bar_sink = (uint32_t)__cvta_generic_to_shared(&bars[2]);    // takes the per-stage commits
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_done));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_ready));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(bar_sink));
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
  if (threadIdx.x < 32) {
    uint32_t leader = 0;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(leader));
    const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(smem), b0 = a0 + 100 * 1024;
    constexpr int N = 32;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_hi = (128u >> 4) | (1u << 14), b_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_base = ((a0 & 0x3FFFFu) >> 4) | ((2048u >> 4) << 16), b_base = ((b0 & 0x3FFFFu) >> 4) | (((uint32_t)N * 16u >> 4) << 16);
    const uint32_t a_step = (uint32_t)n_ring * 0 + 256, b_step = 128, a_plane = 1024;  // runtime-looking constants
    const long long t0 = clock64();
    if (STYLE == 0 || STYLE == 4) {
      uint64_t ad[4], bd[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ad[k] = ((uint64_t)a_hi << 32) | (a_base + k * a_step);
        bd[k] = ((uint64_t)b_hi << 32) | (b_base + k * b_step);
      }
      for (int r = 0; r < reps; r += 8) {
        if (STYLE == 4) {
          bar_wait(bar_ready, 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        if (leader) {
#pragma unroll
          for (int k = 0; k < 8; ++k) mma_l(tm, ad[k & 3], bd[k & 3], idesc, 1);
          if (STYLE == 4) commit(bar_sink);
        }
      }
    } else if (STYLE == 1) {
      uint32_t a_lo = a_base, b_lo = b_base;
#pragma unroll 1
      for (int r = 0; r < reps; ++r) {
        if (leader) mma_w(tm, a_lo, a_hi, b_lo, b_hi, idesc, 1);
        a_lo += a_step; b_lo += b_step;
        if ((r & 3) == 3) { a_lo = a_base; b_lo = b_base; }
      }
    } else if (STYLE == 2 || STYLE == 5) {
      uint32_t a_lo = a_base, b_lo = b_base;
      int rr = 0;
      for (int r = 0; r < reps; r += 6) {
        if (STYLE == 5) {
          bar_wait(bar_ready, 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
#pragma unroll 1
        for (int j = 0; j < 3; ++j) {
          if (leader) {
            mma_w(tm, a_lo, a_hi, b_lo, b_hi, idesc, 1);
            mma_w(tm, a_lo + a_plane, a_hi, b_lo, b_hi, idesc, 1);
          }
          a_lo += a_step; b_lo += b_step;
        }
        if (STYLE == 5 && leader) commit(bar_sink);
        if (++rr == n_ring) { rr = 0; a_lo = a_base; b_lo = b_base; }
      }
    } else if (STYLE == 3) {
      uint64_t ad = ((uint64_t)a_hi << 32) | a_base, bd = ((uint64_t)b_hi << 32) | b_base;
#pragma unroll 1
      for (int r = 0; r < reps; ++r) {
        if (leader) mma_l(tm, ad, bd, idesc, 1);
        ad += a_step; bd += b_step;
        if ((r & 3) == 3) { ad = ((uint64_t)a_hi << 32) | a_base; bd = ((uint64_t)b_hi << 32) | b_base; }
      }
    } else if (STYLE == 6) {
      uint32_t a_lo = a_base, b_lo = b_base;
      int rr = 0;
      for (int r = 0; r < reps; r += 8) {
        uint64_t ah[4], al[4], bb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ah[j] = ((uint64_t)a_hi << 32) | (a_lo + j * a_step);
          al[j] = ((uint64_t)a_hi << 32) | (a_lo + j * a_step + a_plane);
          bb[j] = ((uint64_t)b_hi << 32) | (b_lo + j * b_step);
        }
        if (leader) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            mma_l(tm, ah[j], bb[j], idesc, 1);
            mma_l(tm, al[j], bb[j], idesc, 1);
          }
        }
        a_lo += 4 * a_step; b_lo += 4 * b_step;
        if (++rr == n_ring) { rr = 0; a_lo = a_base; b_lo = b_base; }
      }
    }
    const long long t_issue = clock64();
    if (leader) commit(bar_done);
    bar_wait(bar_done, 0);
    const long long t1 = clock64();
    if (leader) {
      out[2 * blockIdx.x] = t_issue - t0;
      out[2 * blockIdx.x + 1] = t1 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

typedef void (*kern_t)(int, int, long long*);
int main() {
  long long* d;
  cudaMalloc(&d, 2 * 148 * sizeof(long long));
  const int reps = 48000;
  kern_t ks[] = {bench<0>, bench<1>, bench<2>, bench<3>, bench<4>, bench<5>, bench<6>};
  const char* names[] = {"0 invariant descriptors, 8 MMAs/iter", "1 rolled, 1 MMA/iter, dependent adds", "2 rolled, 2 MMAs/iter (fwd loop)",
                         "3 rolled, 64-bit descriptor adds", "4 style 0 + wait/commit per 8 MMAs", "5 style 2 + wait/commit per 6 MMAs (round-1 stage loop)",
                         "6 descriptors of 8 MMAs first, then the MMAs"};
  for (int s = 0; s < 7; ++s) {
    cudaFuncSetAttribute(ks[s], cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    ks[s]<<<148, 128, 200 * 1024>>>(reps, 4, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2 * 148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mi = 0, mt = 0;
    for (int i = 0; i < 148; ++i) { if (h[2 * i] > mi) mi = h[2 * i]; if (h[2 * i + 1] > mt) mt = h[2 * i + 1]; }
    printf("%-58s issue %6.1f cyc/MMA   complete %6.1f cyc/MMA  %s\n", names[s], (double)mi / reps, (double)mt / reps,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
