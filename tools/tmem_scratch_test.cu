// TMEM as a per-thread scratchpad: each thread of a 256-thread CTA parks NCOL 32-bit words in tensor memory with
// tcgen05.st and reads them back with tcgen05.ld (32x32b shape: thread t of warp w owns TMEM lane 32*(w%4) + t%32;
// warps w and w+4 share a lane quarter and use disjoint column ranges).  Two CTAs per SM, 256 columns each.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tmem_scratch_test tools/tmem_scratch_test.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void tmem_st2(uint32_t taddr, double v) {
  const uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ double tmem_ld2(uint32_t taddr) {
  uint32_t lo, hi;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return __hiloint2double((int)hi, (int)lo);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, double (&v)[8]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}

constexpr int KD = 24;   // doubles per (thread, spin)

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#ifndef NPRESS
#define NPRESS 0
#endif
template <int NCOLS>
__global__ void __launch_bounds__(256, 2) tmem_test(unsigned long long* bad, long long* cycles, int iters, unsigned long long* span) {
  double press[NPRESS + 1];
#pragma unroll
  for (int q = 0; q <= NPRESS; ++q) press[q] = threadIdx.x * 0.5 + q;
  extern __shared__ unsigned char dyn[];   // sized to force 2 CTAs/SM like the sweep kernel
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned long long ts0 = gtime();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base_s;
  const unsigned long long ts1 = gtime();
  // this thread's private window: lane quarter of the warp, columns [col0, col0 + 4 KD)
  const uint32_t col0 = (warp >= 4) ? 4 * KD : 0;
  const uint32_t my = base + ((uint32_t)(32 * (warp & 3)) << 16) + col0;
  unsigned long long nbad = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int spin = 0; spin < 2; ++spin)
      for (int m = 0; m < KD; ++m) tmem_st2(my + 2 * (spin * KD + m), (double)(blockIdx.x * 1000003 + tid * 131 + spin * 29 + m + it) * 1.25);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    for (int spin = 0; spin < 2; ++spin) {
      for (int m0 = 0; m0 < KD; m0 += 8) {
        double v[8];
        tmem_ld16(my + 2 * (spin * KD + m0), v);
        for (int q = 0; q < 8; ++q)
          if (v[q] != (double)(blockIdx.x * 1000003 + tid * 131 + spin * 29 + m0 + q + it) * 1.25) ++nbad;
      }
      const double one = tmem_ld2(my + 2 * (spin * KD + (lane % KD)));     // per-warp uniform address, dynamic column
      (void)one;
    }
#pragma unroll
    for (int q = 0; q <= NPRESS; ++q) press[q] = fma(press[q], 1.0000001, press[(q + 1) % (NPRESS + 1)]);
    __syncwarp();
  }
  { double sacc = 0; 
#pragma unroll
    for (int q = 0; q <= NPRESS; ++q) sacc += press[q];
    if (sacc == 1.2345) ++nbad; }
  long long t1 = clock64();
  if (nbad) atomicAdd(bad, nbad);
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(NCOLS) : "memory");
  if (tid == 0) { unsigned smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid)); span[4 * blockIdx.x] = smid; span[4 * blockIdx.x + 1] = ts0; span[4 * blockIdx.x + 2] = ts1; span[4 * blockIdx.x + 3] = gtime(); }
  (void)dyn;
}

template <int NCOLS>
void run(int blocks) {
  unsigned long long* bad; long long* cyc; unsigned long long* span;
  cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8);
  cudaMalloc(&cyc, blocks * 8); cudaMalloc(&span, blocks * 32);
  cudaFuncSetAttribute(tmem_test<NCOLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  tmem_test<NCOLS><<<blocks, 256, 100 * 1024>>>(bad, cyc, 2000, span);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h = 0;
  cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost);
  unsigned long long* hs = new unsigned long long[4 * blocks];
  cudaMemcpy(hs, span, blocks * 32, cudaMemcpyDeviceToHost);
  // blocks sharing an SM: do their [alloc done, end] intervals overlap?  how long did alloc wait?
  int overlap = 0, pairs = 0; double wait_max = 0, run_avg = 0;
  for (int a = 0; a < blocks; ++a) {
    wait_max = (hs[4*a+2]-hs[4*a+1]) > wait_max ? (double)(hs[4*a+2]-hs[4*a+1]) : wait_max;
    run_avg += (double)(hs[4*a+3]-hs[4*a+2]) / blocks;
    for (int b = a + 1; b < blocks; ++b)
      if (hs[4*a] == hs[4*b]) { ++pairs; if (hs[4*a+2] < hs[4*b+3] && hs[4*b+2] < hs[4*a+3]) ++overlap; }
  }
  printf("cols %d blocks %d: status %s mismatches %llu; same-SM pairs %d, overlapping in time %d; max alloc wait %.0f ns, mean run %.0f ns\n",
         NCOLS, blocks, cudaGetErrorString(e), h, pairs, overlap, wait_max, run_avg);
}

int main() {
  for (int smem : {100 * 1024, 109808}) {
    int n = -1;
    cudaFuncSetAttribute(tmem_test<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, tmem_test<256>, 256, smem);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, tmem_test<256>);
    printf("occupancy API: %d CTAs/SM at %d B dynamic smem (regs %d, static smem %zu)\n", n, smem, fa.numRegs, fa.sharedSizeBytes);
  }
  for (int blocks : {148, 149, 296}) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    unsigned long long* bad; long long* cyc; unsigned long long* span;
    cudaMalloc(&bad, 8); cudaMalloc(&cyc, blocks * 8); cudaMalloc(&span, blocks * 32);
    cudaEventRecord(a);
    tmem_test<256><<<blocks, 256, 100 * 1024>>>(bad, cyc, 2000, span);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("blocks %d: %.3f ms\n", blocks, ms);
  }
  run<256>(296);
  return 0;
}
