// What does tcgen05.ld.16x256b deliver to each thread?  Every thread of a 256-thread CTA stores, with the 32x32b shape, the word
// value  1000 * (its TMEM lane) + column  into columns [0, 16) of its own lane (warps 0-3 only: window 0); then each warp reads with
// 16x256b.x1 at lane offsets 0 and 16 of its quarter and prints what lanes 0..7 of warp 0 and warp 2 received.
// Expected (PTX ISA matrix-fragment figure, same as the m16n8 accumulator layout): thread t gets words
//   r0, r1 = [lane base + t/4][col0 + 2 (t%4) + {0,1}],  r2, r3 = [lane base + t/4 + 8][same columns].
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tmem_frag_test tools/tmem_frag_test.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__global__ void __launch_bounds__(256, 1) k(uint32_t* out) {
  __shared__ uint32_t base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&base_s)), "n"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = base_s;
  const int tl = 32 * (warp & 3) + lane;                    // this thread's TMEM lane
  if (warp < 4) {
    for (int c = 0; c < 16; ++c)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(base + ((uint32_t)(32 * (warp & 3)) << 16) + c), "r"((uint32_t)(1000 * tl + c)) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int half = 0; half < 2; ++half)
    for (int col0 = 0; col0 < 16; col0 += 8) {
      uint32_t r0, r1, r2, r3;
      const uint32_t ta = base + ((uint32_t)(32 * (warp & 3) + 16 * half) << 16) + col0;
      asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(ta) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      uint32_t* o = out + (((warp * 2 + half) * 2 + col0 / 8) * 32 + lane) * 4;
      o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3;
    }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(64) : "memory");
}
int main() {
  uint32_t* d; cudaMalloc(&d, 8 * 2 * 2 * 32 * 4 * 4);
  k<<<1, 256>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  static uint32_t h[8 * 2 * 2 * 32 * 4];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int warp = 0; warp < 8; ++warp)
    for (int half = 0; half < 2; ++half)
      for (int cb = 0; cb < 2; ++cb)
        for (int t = 0; t < 32; ++t) {
          const uint32_t* o = h + (((warp * 2 + half) * 2 + cb) * 32 + t) * 4;
          const int lb = 32 * (warp & 3) + 16 * half, c = 8 * cb + 2 * (t % 4);
          const uint32_t want[4] = {(uint32_t)(1000 * (lb + t / 4) + c), (uint32_t)(1000 * (lb + t / 4) + c + 1),
                                    (uint32_t)(1000 * (lb + t / 4 + 8) + c), (uint32_t)(1000 * (lb + t / 4 + 8) + c + 1)};
          for (int i = 0; i < 4; ++i) if (o[i] != want[i]) { if (bad < 12) printf("warp %d half %d cb %d t %d r%d = %u want %u\n", warp, half, cb, t, i, o[i], want[i]); ++bad; }
        }
  printf("16x256b.x1 fragment layout: %d mismatches against  r0,r1 = [lane t/4][2(t%%4)+{0,1}], r2,r3 = [lane t/4+8][...]\n", bad);
  for (int t = 0; t < 8; ++t) { const uint32_t* o = h + t * 4; printf("warp0 half0 cols0-7 thread %d: %u %u %u %u\n", t, o[0], o[1], o[2], o[3]); }
  return bad != 0;
}
