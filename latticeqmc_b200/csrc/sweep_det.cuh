// det-mode sweep: the reference's slow validation sampler `_update_step_det` (lqmc.py:236-259) on the device.
//
// Per proposal the reference flips h[i,l], rebuilds M_sigma(l) = I + B_{l-1} ... B_0 B_{L-1} ... B_l from the field
// (get_m, lqmc.py:156-185: B = exp_k . diag(exp(-sigma lamb h[:,l])), accumulated left to right), takes
// np.linalg.det of both, and accepts on u <= det(M_up) det(M_dn) / old_det; a rejected flip is undone.  The slice being
// updated is always the LAST factor of that left-to-right product, so the product of the first L-1 factors is the same
// for all N proposals of a slice: it is built once per slice (L-1 GEMMs) and every proposal costs one GEMM and one LU per
// spin - the same floating-point operations in the same association order as rebuilding everything, L times cheaper.
//
// One CTA per chain, 512 threads = 2 spins x 256.  N <= 64: all matrices live in shared memory.  Larger lattices (the
// reference has no size limit, it is just O(L N^3) per proposal): the same code on a per-chain global-memory workspace
// (4 N^2 doubles, L2-resident) and exp(-dtau K) read in place - a validation tool, not a roofline kernel.  det = product of the
// LU pivots with partial pivoting (first entry of largest magnitude, the getrf rule); NumPy forms sign * exp(sum log|u_ii|)
// instead (umath_linalg det), so ratios agree to ~1e-13 relative, not bitwise.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "philox.h"

namespace lqmc {

constexpr int DET_THREADS = 512;
constexpr int DET_TPS = 256;       // threads per spin
constexpr int DET_MAX_N = 64;     // largest N with the matrices in shared memory

struct DetParams {
  int n_sites, n_slices, NPf, ldE;
  const double* E;         // exp(-dtau K), row-major with row stride ldE
  int8_t* field;           // [chain][slice][NPf]
  const double* uniforms;  // device [chain][buf_sweeps][L][N] or nullptr (Philox)
  uint64_t seed;
  long long sweep0, chain0;
  int n_sweeps, buf_sweeps, buf_sweep0;
  double* det_old;         // [chain] carried between launches of one loop
  int init_det;            // 1: old_det = det M_up(0) det M_dn(0) from the field (lqmc.py:264-268, 283-287)
  long long* n_acc;        // [chain]
  double* tr_ratio;        // [chain][buf_sweeps][L][N] or nullptr
  uint8_t* tr_acc;
  double exp_pl, exp_ml;
  double* work;            // nullptr: matrices in shared memory (N <= DET_MAX_N); else [chain][4][N][N] global workspace
};

inline size_t det_smem_bytes(int N, int L) {
  const size_t mats = N <= DET_MAX_N ? (size_t)5 * N * N : 0;
  return (mats + 8) * sizeof(double) + 8 * sizeof(int) + (size_t)N * L + 16;
}

__device__ __forceinline__ void det_bar(int spin) { asm volatile("bar.sync %0, %1;" ::"r"(1 + spin), "n"(DET_TPS) : "memory"); }

// out = in . (E diag(v_l)) for this spin (in == nullptr: the scalar 1 of get_m's first step, lqmc.py:179-183)
__device__ __forceinline__ void det_gemm(double* __restrict__ out, const double* __restrict__ in, const double* __restrict__ Es, int ldE,
                                         const int8_t* __restrict__ hl, int N, int spin, int t, double exp_pl, double exp_ml) {
  for (int e = t; e < N * N; e += DET_TPS) {
    const int r = e / N, c = e - r * N;
    const double v = ((hl[c] > 0) != (spin != 0)) ? exp_ml : exp_pl;
    if (in == nullptr) {
      out[e] = __dmul_rn(Es[r * ldE + c], v);
    } else {
      double acc = 0.0;
      const double* row = in + r * N;
      for (int k = 0; k < N; ++k) acc = fma(row[k], __dmul_rn(Es[k * ldE + c], v), acc);
      out[e] = acc;
    }
  }
}

// det of the N x N matrix A (destroyed) by LU with partial pivoting; every thread of the spin group returns it
__device__ double det_lu(double* __restrict__ A, int N, int spin, int t, int* pivrow) {
  double det = 1.0;
  for (int k = 0; k < N; ++k) {
    if (t < 32) {
      double best = -1.0;
      int idx = k;
      for (int r = k + t; r < N; r += 32) {
        const double a = fabs(A[r * N + k]);
        if (a > best) { best = a; idx = r; }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
        if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
      }
      if (t == 0) pivrow[spin] = idx;
    }
    det_bar(spin);
    const int pr = pivrow[spin];
    if (pr != k) {
      det = -det;
      for (int c = t; c < N; c += DET_TPS) { const double x = A[k * N + c]; A[k * N + c] = A[pr * N + c]; A[pr * N + c] = x; }
    }
    det_bar(spin);
    const double akk = A[k * N + k];
    det = __dmul_rn(det, akk);
    const int n = N - k - 1;
    if (akk != 0.0 && n > 0) {
      for (int r = t; r < n; r += DET_TPS) A[(k + 1 + r) * N + k] = A[(k + 1 + r) * N + k] / akk;
      det_bar(spin);
      for (int e = t; e < n * n; e += DET_TPS) {
        const int rr = e / n, cc = e - rr * n;
        const int r = k + 1 + rr, c = k + 1 + cc;
        A[r * N + c] = fma(-A[r * N + k], A[k * N + c], A[r * N + c]);
      }
    }
    det_bar(spin);
  }
  return det;
}

__global__ void __launch_bounds__(DET_THREADS, 1) sweep_det_kernel(const DetParams p) {
  extern __shared__ __align__(16) unsigned char det_smem_raw[];
  const int N = p.n_sites, L = p.n_slices, NN = N * N;
  const int chain = blockIdx.x, tid = threadIdx.x;
  const int spin = tid / DET_TPS, t = tid % DET_TPS;
  const bool glob = p.work != nullptr;
  double* sm_d = reinterpret_cast<double*>(det_smem_raw);
  const double* Es = glob ? p.E : sm_d;                                          // [N][N], row stride ldE
  const int ldE = glob ? p.ldE : N;
  double* bufA = glob ? p.work + (size_t)chain * 4 * NN : sm_d + NN;             // [2 spin][N][N] prefix product of the first L-1 factors
  double* bufM = bufA + 2 * NN;                                                  // [2 spin][N][N] ping-pong partner / LU workspace
  double* dets = glob ? sm_d : bufM + 2 * NN;                                    // [2]
  int* pivrow = reinterpret_cast<int*>(dets + 8);           // [2]
  int8_t* hf = reinterpret_cast<int8_t*>(pivrow + 8);       // [L][N]
  int8_t* field = p.field + (size_t)chain * L * p.NPf;
  if (!glob)
    for (int e = tid; e < NN; e += DET_THREADS) sm_d[e] = p.E[(size_t)(e / N) * p.ldE + (e % N)];
  for (int e = tid; e < N * L; e += DET_THREADS) hf[e] = field[(size_t)(e / N) * p.NPf + (e % N)];
  __syncthreads();
  double* A = bufA + spin * NN;
  double* M = bufM + spin * NN;

  // prefix(l0): product of the first L-1 factors of get_m(l0): slices l0-1, l0-2, ... (cyclic), ending in A
  auto prefix = [&](int l0) {
    const int nf = L - 1;
    double* cur = (nf & 1) ? A : M;          // the last of nf writes lands in A
    const double* prev = nullptr;
    for (int m = 0; m < nf; ++m) {
      const int l = ((l0 - 1 - m) % L + L) % L;
      det_gemm(cur, prev, Es, ldE, hf + l * N, N, spin, t, p.exp_pl, p.exp_ml);
      det_bar(spin);
      prev = cur;
      cur = (cur == A) ? M : A;
    }
  };
  // finish(l0): M = I + prefix . B_{l0}; returns det M_up * det M_dn to every thread
  auto finish = [&](int l0) -> double {
    det_gemm(M, (L > 1) ? A : nullptr, Es, ldE, hf + l0 * N, N, spin, t, p.exp_pl, p.exp_ml);
    det_bar(spin);
    for (int r = t; r < N; r += DET_TPS) M[r * N + r] = __dadd_rn(M[r * N + r], 1.0);
    det_bar(spin);
    const double d = det_lu(M, N, spin, t, pivrow);
    if (t == 0) dets[spin] = d;
    __syncthreads();
    const double prod = __dmul_rn(dets[0], dets[1]);
    __syncthreads();
    return prod;
  };

  double old_det;
  if (p.init_det) {
    prefix(0);
    old_det = finish(0);
  } else {
    old_det = p.det_old[chain];
  }
  int n_accepted = 0;
  for (int sweep = 0; sweep < p.n_sweeps; ++sweep) {
    for (int step = 0; step < L; ++step) {
      const int l = L - 1 - step;
      prefix(l);
      const long long base = (((long long)chain * p.buf_sweeps + p.buf_sweep0 + sweep) * L + step) * N;
      for (int i = 0; i < N; ++i) {
        __syncthreads();
        if (tid == 0) hf[l * N + i] = (int8_t)(-hf[l * N + i]);     // config.update(i, l), lqmc.py:241
        __syncthreads();
        const double new_det = finish(l);
        const double ratio = new_det / old_det;
        const double u = (p.uniforms != nullptr)
                             ? p.uniforms[base + i]
                             : lqmc_philox_uniform(p.seed, (uint64_t)(p.chain0 + chain), (uint64_t)(p.sweep0 + sweep), (uint32_t)(step * N + i));
        const bool acc = u <= ratio;
        if (acc) { old_det = new_det; ++n_accepted; }
        else if (tid == 0) hf[l * N + i] = (int8_t)(-hf[l * N + i]);      // revert, lqmc.py:256
        if (tid == 0 && p.tr_ratio != nullptr) { p.tr_ratio[base + i] = ratio; p.tr_acc[base + i] = acc ? 1 : 0; }
      }
      __syncthreads();
      for (int c = tid; c < N; c += DET_THREADS) field[(size_t)l * p.NPf + c] = hf[l * N + c];
    }
  }
  if (tid == 0) {
    p.det_old[chain] = old_det;
    if (n_accepted) p.n_acc[chain] += n_accepted;
  }
}

}  // namespace lqmc
