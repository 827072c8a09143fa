// FP64 pipe microbenchmark for B200 (sm_100a).
//
// MEASURED_PEAKS.json carries HBM and bf16 numbers only; the sweep engine is
// FP64-bound, so the roofline denominator for every kernel in this repo is
// measured here: DFMA (vector FP64 pipe), DMUL+DADD (the no-contraction pair
// the exact-arithmetic path uses), DMMA (mma.sync .f64 shapes), and both pipes
// at once (are they one physical pipe on this part?).  Also the latency
// constants the serial Metropolis loop is bound by: FP64 divide, bar.sync,
// shared-memory round trip.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/fp64_peak tools/fp64_peak.cu
// Run  : tools/fp64_peak [out.json]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <string>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

// ---- DFMA: 16 independent accumulators per thread ------------------------------------------
__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- DMUL then DSUB, no contraction: acc = acc - x*y --------------------------------------
__global__ void __launch_bounds__(256) k_dmulsub(double* out, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = __dsub_rn(b, __dmul_rn(acc[i], a));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- DMMA shapes ---------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void __launch_bounds__(256) k_dmma884(double* out, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_dmma1684(double* out, double a, double b) {
  double c[8][4]; double av[2] = {a, a + 1};
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma1684(c[i], av, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_dmma1688(double* out, double a, double b) {
  double c[8][4]; double av[4] = {a, a + 1, a + 2, a + 3}; double bv[2] = {b, b + 1};
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma1688(c[i], av, bv);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_dmma16816(double* out, double a, double b) {
  double c[8][4]; double av[8]; double bv[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) av[i] = a + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) bv[i] = b + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma16816(c[i], av, bv);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- both pipes at once: even warps DFMA, odd warps DMMA m8n8k4 ------------------------------
__global__ void __launch_bounds__(256) k_mixed(double* out, double a, double b) {
  const int warp = threadIdx.x >> 5;
  double s = 0;
  if (warp & 1) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma884(c[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  } else {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- FP64 divide throughput (8 independent per thread) ---------------------------------------
__global__ void __launch_bounds__(256) k_ddiv(double* out, double a, double b) {
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 1.0 + threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = (acc[i] + a) / b;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- latency probes (one CTA, clock64) -------------------------------------------------------
__global__ void k_lat(long long* out, double a, double b, int n) {
  __shared__ double sm[256];
  sm[threadIdx.x] = a + threadIdx.x;
  __syncthreads();
  // dependent DFMA chain
  double x = a;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = fma(x, b, a);
  long long t1 = clock64();
  // dependent divide chain
  double y = a + 3.0;
  for (int i = 0; i < n; ++i) y = (y + a) / b;
  long long t2 = clock64();
  // bar.sync chain
  for (int i = 0; i < n; ++i) __syncthreads();
  long long t3 = clock64();
  // smem store -> bar -> load round trip (what one accepted flip pays)
  double z = x;
  for (int i = 0; i < n; ++i) {
    sm[(threadIdx.x + 1) % blockDim.x] = z;
    __syncthreads();
    z = sm[threadIdx.x] + 1.0;
    __syncthreads();
  }
  long long t4 = clock64();
  // dependent smem load chain
  volatile double* vs = sm;
  int idx = threadIdx.x;
  for (int i = 0; i < n; ++i) idx = (int)vs[idx & 255] & 255;
  long long t5 = clock64();
  if (threadIdx.x == 0) {
    out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = t5 - t4;
  }
  if (x + y + z + idx == 12345.678) out[7] = 1;
}

template <typename K>
static double time_kernel(K kern, int blocks, int threads, double* d_out, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) kern<<<blocks, threads>>>(d_out, 1.0000001, 1e-9);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    kern<<<blocks, threads>>>(d_out, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
  }
  CK(cudaGetLastError());
  return best * 1e-3;
}

// sustained: run back to back for ~secs seconds
template <typename K>
static double sustained(K kern, int blocks, int threads, double* d_out, double flops_per_launch, double secs) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double one = time_kernel(kern, blocks, threads, d_out, 2);
  int n = std::max(8, (int)(secs / one));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < n; ++i) kern<<<blocks, threads>>>(d_out, 1.0000001, 1e-9);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return flops_per_launch * n / (ms * 1e-3) * 1e-12;
}

int main(int argc, char** argv) {
  const char* outpath = argc > 1 ? argv[1] : nullptr;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const int threads = 256, blocks = sms * 8;
  double* d_out; CK(cudaMalloc(&d_out, sizeof(double) * blocks * threads));
  const double nthreads = (double)blocks * threads;
  const double nwarps = nthreads / 32;

  struct Row { std::string name; double tflops; double sustained; };
  std::vector<Row> rows;
  auto add = [&](const char* name, double fl, double t, double sus) {
    rows.push_back({name, fl / t * 1e-12, sus});
    printf("%-14s burst %8.2f TFLOP/s   sustained %8.2f TFLOP/s\n", name, fl / t * 1e-12, sus);
  };
  {
    double fl = nthreads * 16.0 * ITERS * 2;
    add("dfma", fl, time_kernel(k_dfma, blocks, threads, d_out), sustained(k_dfma, blocks, threads, d_out, fl, 2.0));
  }
  {
    double fl = nthreads * 16.0 * ITERS * 2;  // counted as the 2 flops of the mul-sub pair
    add("dmul_dsub", fl, time_kernel(k_dmulsub, blocks, threads, d_out), sustained(k_dmulsub, blocks, threads, d_out, fl, 1.0));
  }
  {
    double fl = nwarps * 8.0 * ITERS * (8 * 8 * 4) * 2;
    add("dmma_m8n8k4", fl, time_kernel(k_dmma884, blocks, threads, d_out), sustained(k_dmma884, blocks, threads, d_out, fl, 2.0));
  }
  {
    double fl = nwarps * 8.0 * ITERS * (16 * 8 * 4) * 2;
    add("dmma_m16n8k4", fl, time_kernel(k_dmma1684, blocks, threads, d_out), 0);
  }
  {
    double fl = nwarps * 8.0 * ITERS * (16 * 8 * 8) * 2;
    add("dmma_m16n8k8", fl, time_kernel(k_dmma1688, blocks, threads, d_out), 0);
  }
  {
    double fl = nwarps * 8.0 * ITERS * (16 * 8 * 16) * 2;
    add("dmma_m16n8k16", fl, time_kernel(k_dmma16816, blocks, threads, d_out), 0);
  }
  {
    double fl = (nthreads / 2) * 16.0 * ITERS * 2 + (nwarps / 2) * 8.0 * ITERS * 256 * 2;
    add("mixed_dfma_dmma", fl, time_kernel(k_mixed, blocks, threads, d_out), 0);
  }
  double ddiv_gops;
  {
    double ops = nthreads * 8.0 * (ITERS / 8);
    double t = time_kernel(k_ddiv, blocks, threads, d_out);
    ddiv_gops = ops / t * 1e-9;
    printf("%-14s %8.2f Gdiv/s (with one dependent DADD each)\n", "ddiv", ddiv_gops);
  }
  long long* d_lat; CK(cudaMalloc(&d_lat, 8 * sizeof(long long)));
  long long lat[3][8];
  const int n = 2048;
  int tcounts[3] = {32, 128, 256};
  for (int c = 0; c < 3; ++c) {
    k_lat<<<1, tcounts[c]>>>(d_lat, 1.0000001, 1.0000002, n);
    CK(cudaDeviceSynchronize());
    k_lat<<<1, tcounts[c]>>>(d_lat, 1.0000001, 1.0000002, n);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(lat[c], d_lat, sizeof(long long) * 8, cudaMemcpyDeviceToHost));
    printf("threads=%3d  dfma_lat %.1f  ddiv_lat %.1f  bar %.1f  sts_bar_lds_bar %.1f  lds_chain %.1f cyc\n", tcounts[c],
           (double)lat[c][0] / n, (double)lat[c][1] / n, (double)lat[c][2] / n, (double)lat[c][3] / n, (double)lat[c][4] / n);
  }
  if (outpath) {
    FILE* f = fopen(outpath, "w");
    fprintf(f, "{\n \"gpu_name\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d,\n", prop.name, sms, prop.clockRate);
    for (auto& r : rows) fprintf(f, " \"%s_tflops\": %.3f, \"%s_tflops_sustained\": %.3f,\n", r.name.c_str(), r.tflops, r.name.c_str(), r.sustained);
    fprintf(f, " \"ddiv_gops\": %.3f,\n", ddiv_gops);
    for (int c = 0; c < 3; ++c)
      fprintf(f, " \"lat_cycles_t%d\": {\"dfma\": %.2f, \"ddiv_plus_dadd\": %.2f, \"bar_sync\": %.2f, \"sts_bar_lds_bar\": %.2f, \"lds_dependent\": %.2f}%s\n",
              tcounts[c], (double)lat[c][0] / n, (double)lat[c][1] / n, (double)lat[c][2] / n, (double)lat[c][3] / n, (double)lat[c][4] / n, c < 2 ? "," : "");
    fprintf(f, "}\n");
    fclose(f);
  }
  return 0;
}
