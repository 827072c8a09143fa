// Register-resident sweep kernel: one CTA per Markov chain, N <= 64 sites.
//
// Replaces, for a whole batch of chains at once, LatticeQMC._update_step of the reference
// (/root/reference/lqmc/lqmc.py:301-347) and the helpers it calls (get_m :156-185, get_exp_v :132-154,
// np.linalg.inv :306-307, Configuration.update configuration.py:126-136, measure_loop :356-375).
//
// Layout.  128 threads = 2 spins x (8 x 8) threads.  Each thread owns a TR x TR register tile of its
// spin's Green's function (TR = NP/8; NP = N padded to 16/32/64): rows {2*ty + 16*q + s}, columns
// {2*tx + 16*q + s}.  The interleave makes every shared-memory fragment access of the GEMMs a
// conflict-free LDS.128 and lets the site loop be unrolled so that "the row / column of site i" is a
// compile-time register index.  G never leaves the register file during the N proposals of a slice:
// a flip publishes row i and column i (2*NP doubles per spin) through shared memory, one bar.sync,
// then every thread applies the rank-1 update to its own tile.  Per flip that is ~1 KB of shared
// traffic instead of the 16*N^2 B a shared-memory-resident G would move, so the phase is bound by
// the FP64 pipe (2*N^2 DFMA per flip), not by shared-memory bandwidth.
//
// The wrap G <- B G B^-1 and the sweep-start product are register-tiled DFMA GEMMs: one operand
// staged in shared memory (the chain's own G / running product), the other (exp(-dtau K) or its
// inverse, shared by every chain on the SM) read through L1 with ld.global.nc.  exp(V_l) is
// diagonal and is folded into the epilogue as row / column scales (lqmc.py:339-345 builds it dense).
// FP64 has no tcgen05 kind and DMMA shares the DFMA pipe on sm_100a (profiles/fp64_peaks_r01.json:
// 36.8 vs 37.2 TFLOP/s, 30.9 mixed), so the roofline is the DFMA pipe either way.
//
// Arithmetic modes (template flags):
//   EXACT  - ratio, rank-1 vectors and update use the reference's roundings: separate multiply and
//            subtract, true division (lqmc.py:314-331).  Given the same G, field and uniforms the
//            accept/reject decisions and the updated G of a slice are bit-identical to NumPy's.
//   !EXACT - update contracted to one FMA, e = column * (1/denominator).
//   PHYS   - textbook DQMC instead of the reference recurrence.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "philox.h"

namespace lqmc {

struct SweepParams {
  int n_sites, n_slices, n_chains;
  const double* E;    // exp(-dtau K)            [NP][NP] row-major, identity-padded
  const double* Et;   // its transpose
  const double* Ei;   // exp(+dtau K)
  const double* Eit;  // its transpose
  int8_t* field;      // [chain][slice][NP], pad = +1
  double* G;          // [chain][2][NP][NP]
  const double* uniforms;  // device [chain][n_sweeps][n_steps][N] or nullptr (Philox)
  uint64_t seed;
  long long sweep0;   // global sweep index of the first sweep of this launch
  long long chain0;   // global index of chain 0 of this engine
  double* g_sum;      // [chain][2][N][N]
  double* obs_sum;    // [chain][3][N]
  long long* n_meas;  // [chain]
  long long* n_acc;   // [chain]
  double* tr_ratio;   // trace [chain][n_sweeps][n_steps][N] or nullptr
  uint8_t* tr_acc;
  int n_sweeps, step_lo, step_hi;
  int do_recompute, do_propose, do_wrap, measure, recompute_l0;
  double exp_pl, exp_ml, f_p2, f_m2;  // exp(+lamb), exp(-lamb), exp(+2 lamb)-1, exp(-2 lamb)-1
};

template <int NP>
struct RegCfg {
  static constexpr int TR = NP / 8;       // tile edge per thread
  static constexpr int S = NP + 2;        // shared row stride (doubles): even for 16-B vectors, 2 mod 4 for banks
  static constexpr int THREADS = 128;
  static constexpr size_t stage_bytes = size_t(2) * NP * S * sizeof(double);
  static constexpr size_t smem_bytes = stage_bytes + (2 * NP + 4 * NP * 2 + NP) * sizeof(double) + 2 * NP + 64 * 4 + 2 * NP * sizeof(int);
};

template <int NP>
struct RegSmem {
  double* stage;   // [2][NP][S]
  double* d;       // [2][NP]      diagonal as of the previous accepted flip
  double* e;       // [2 buf][2 spin][NP]
  double* c;       // [2 buf][2 spin][NP]
  double* u;       // [NP]
  int8_t* h;       // [NP] field column of the slice being updated
  int8_t* hn;      // [NP] field column the wrap scales with
  double* red_v;   // [4] pivot search partials (per warp)
  int* red_i;      // [4]
  int* piv;        // [2][NP]
  __device__ explicit RegSmem(unsigned char* base) {
    constexpr int S = RegCfg<NP>::S;
    stage = reinterpret_cast<double*>(base);
    d = stage + 2 * NP * S;
    e = d + 2 * NP;
    c = e + 4 * NP;
    u = c + 4 * NP;
    red_v = u + NP;
    red_i = reinterpret_cast<int*>(red_v + 4);
    piv = red_i + 4;
    h = reinterpret_cast<int8_t*>(piv + 2 * NP);
    hn = h + NP;
  }
};

template <bool EXACT>
__device__ __forceinline__ double rank1(double g, double e, double c) {
  if (EXACT) return __dsub_rn(g, __dmul_rn(e, c));
  return fma(-e, c, g);
}

// ---- register-tiled DFMA GEMMs -----------------------------------------------------------------
// acc[a][b] += sum_k At[k][row_a] * Bs[k][col_b]      (At: global, transposed operand; Bs: shared)
template <int NP>
__device__ __forceinline__ void gemm_gT_s(double (&acc)[RegCfg<NP>::TR][RegCfg<NP>::TR], const double* __restrict__ At,
                                          const double* Bs, int ty, int tx) {
  constexpr int TR = RegCfg<NP>::TR, S = RegCfg<NP>::S, Q = TR / 2;
  const double* ap = At + 2 * ty;
  const double* bp = Bs + 2 * tx;
#pragma unroll 2
  for (int k = 0; k < NP; ++k) {
    double a[TR], b[TR];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(ap + k * NP + 16 * q));
      a[2 * q] = v.x; a[2 * q + 1] = v.y;
      const double2 w = *reinterpret_cast<const double2*>(bp + k * S + 16 * q);
      b[2 * q] = w.x; b[2 * q + 1] = w.y;
    }
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
      for (int j = 0; j < TR; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
  }
}

// acc[a][b] += sum_k As[row_a][k] * Bg[k][col_b]      (As: shared, row-major; Bg: global)
template <int NP>
__device__ __forceinline__ void gemm_s_g(double (&acc)[RegCfg<NP>::TR][RegCfg<NP>::TR], const double* As,
                                         const double* __restrict__ Bg, int ty, int tx) {
  constexpr int TR = RegCfg<NP>::TR, S = RegCfg<NP>::S, Q = TR / 2;
  const double* bp = Bg + 2 * tx;
#pragma unroll 1
  for (int k = 0; k < NP; k += 2) {
    double2 a[TR];
    double b0[TR], b1[TR];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
#pragma unroll
      for (int s = 0; s < 2; ++s) a[2 * q + s] = *reinterpret_cast<const double2*>(As + (2 * ty + 16 * q + s) * S + k);
      const double2 w0 = __ldg(reinterpret_cast<const double2*>(bp + k * NP + 16 * q));
      const double2 w1 = __ldg(reinterpret_cast<const double2*>(bp + (k + 1) * NP + 16 * q));
      b0[2 * q] = w0.x; b0[2 * q + 1] = w0.y;
      b1[2 * q] = w1.x; b1[2 * q + 1] = w1.y;
    }
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
      for (int j = 0; j < TR; ++j) acc[i][j] = fma(a[i].x, b0[j], acc[i][j]);
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
      for (int j = 0; j < TR; ++j) acc[i][j] = fma(a[i].y, b1[j], acc[i][j]);
  }
}

template <int NP>
__device__ __forceinline__ void store_tile(double* M, const double (&g)[RegCfg<NP>::TR][RegCfg<NP>::TR], int ty, int tx) {
  constexpr int TR = RegCfg<NP>::TR, S = RegCfg<NP>::S;
#pragma unroll
  for (int a = 0; a < TR; ++a) {
    const int row = 2 * ty + 16 * (a >> 1) + (a & 1);
#pragma unroll
    for (int q = 0; q < TR / 2; ++q)
      *reinterpret_cast<double2*>(M + row * S + 2 * tx + 16 * q) = make_double2(g[a][2 * q], g[a][2 * q + 1]);
  }
}

template <int NP>
__device__ __forceinline__ void load_tile(const double* M, int stride, double (&g)[RegCfg<NP>::TR][RegCfg<NP>::TR], int ty, int tx) {
  constexpr int TR = RegCfg<NP>::TR;
#pragma unroll
  for (int a = 0; a < TR; ++a) {
    const int row = 2 * ty + 16 * (a >> 1) + (a & 1);
#pragma unroll
    for (int q = 0; q < TR / 2; ++q) {
      const double2 v = *reinterpret_cast<const double2*>(M + row * stride + 2 * tx + 16 * q);
      g[a][2 * q] = v.x; g[a][2 * q + 1] = v.y;
    }
  }
}

// exp(-sigma*lamb*h) for spin index `spin` (0: sigma=+1, 1: sigma=-1)   [get_exp_v, lqmc.py:149-154]
__device__ __forceinline__ double hs_v(int8_t h, int spin, const SweepParams& p) {
  return ((h > 0) != (spin != 0)) ? p.exp_ml : p.exp_pl;
}
__device__ __forceinline__ double hs_vinv(int8_t h, int spin, const SweepParams& p) {
  return ((h > 0) != (spin != 0)) ? p.exp_pl : p.exp_ml;
}

// ---- in-place Gauss-Jordan inverse with partial (row) pivoting, 64 threads per spin -------------------
// Stands in for np.linalg.inv (LAPACK getrf/getri, lqmc.py:306-307): same pivot rule (first entry of
// largest magnitude in the column), different elimination order.
template <int NP>
__device__ void gj_inverse(double* M, RegSmem<NP>& sm, int spin, int t) {
  constexpr int S = RegCfg<NP>::S;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* piv = sm.piv + spin * NP;
  for (int k = 0; k < NP; ++k) {
    double pv = (t >= k && t < NP) ? M[t * S + k] : 0.0;
    double av = (t >= k && t < NP) ? fabs(pv) : -1.0;
    int idx = t;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const double oa = __shfl_down_sync(0xffffffffu, av, off);
      const double op = __shfl_down_sync(0xffffffffu, pv, off);
      const int oi = __shfl_down_sync(0xffffffffu, idx, off);
      if (oa > av || (oa == av && oi < idx)) { av = oa; pv = op; idx = oi; }
    }
    if (lane == 0) { sm.red_v[warp] = pv; sm.red_i[warp] = idx; }
    __syncthreads();
    {
      const double p0 = sm.red_v[2 * spin], p1 = sm.red_v[2 * spin + 1];
      const int i0 = sm.red_i[2 * spin], i1 = sm.red_i[2 * spin + 1];
      const bool second = (i1 >= k && i1 < NP) && (!(i0 >= k && i0 < NP) || fabs(p1) > fabs(p0));
      pv = second ? p1 : p0;
      idx = second ? i1 : i0;
    }
    if (t == 0) piv[k] = idx;
    if (t < NP) {
      const double ak = M[k * S + t];
      const double ap = M[idx * S + t];
      const double x = (t == k) ? 1.0 : ap;
      M[k * S + t] = x / pv;
      if (idx != k) M[idx * S + t] = ak;
    }
    __syncthreads();
    if (t < NP && t != k) {
      double* row = M + t * S;
      const double* rk = M + k * S;
      const double f = row[k];
      row[k] = 0.0;
#pragma unroll 4
      for (int j = 0; j < NP; j += 2) {
        double2 r = *reinterpret_cast<double2*>(row + j);
        const double2 q = *reinterpret_cast<const double2*>(rk + j);
        r.x = fma(-f, q.x, r.x);
        r.y = fma(-f, q.y, r.y);
        *reinterpret_cast<double2*>(row + j) = r;
      }
    }
    __syncthreads();
  }
  if (t < NP) {
    double* row = M + t * S;
    for (int k = NP - 1; k >= 0; --k) {
      const int p = piv[k];
      if (p != k) { const double tmp = row[k]; row[k] = row[p]; row[p] = tmp; }
    }
  }
  __syncthreads();
}

// ---- sweep-start Green's function: G = inv(I + B_{s0} B_{s1} ... ), slice order (l0-1-m) mod L ---------
// get_m + np.linalg.inv (lqmc.py:156-185,303-307).  The first factor is taken as is (the reference starts
// its left-to-right product from the scalar 1); each later one costs one GEMM with the column scale
// exp(V_l) in the epilogue.
template <int NP>
__device__ void recompute_g(double (&g)[RegCfg<NP>::TR][RegCfg<NP>::TR], RegSmem<NP>& sm, const SweepParams& p,
                            const int8_t* field, int l0, int spin, int t, int ty, int tx) {
  constexpr int TR = RegCfg<NP>::TR, S = RegCfg<NP>::S;
  const int L = p.n_slices;
  double* stage = sm.stage + spin * NP * S;
  int l = (l0 - 1 + L) % L;
  {
    const int8_t* hl = field + l * NP;
#pragma unroll
    for (int a = 0; a < TR; ++a) {
      const int row = 2 * ty + 16 * (a >> 1) + (a & 1);
#pragma unroll
      for (int b = 0; b < TR; ++b) {
        const int col = 2 * tx + 16 * (b >> 1) + (b & 1);
        g[a][b] = p.E[row * NP + col] * hs_v(hl[col], spin, p);
      }
    }
  }
  for (int m = 1; m < L; ++m) {
    l = (l0 - 1 - m + 2 * L) % L;
    __syncthreads();                 // everyone is done reading the previous stage contents
    store_tile<NP>(stage, g, ty, tx);
    __syncthreads();
    double acc[TR][TR];
#pragma unroll
    for (int a = 0; a < TR; ++a)
#pragma unroll
      for (int b = 0; b < TR; ++b) acc[a][b] = 0.0;
    gemm_s_g<NP>(acc, stage, p.E, ty, tx);
    const int8_t* hl = field + l * NP;
#pragma unroll
    for (int b = 0; b < TR; ++b) {
      const int col = 2 * tx + 16 * (b >> 1) + (b & 1);
      const double v = hs_v(hl[col], spin, p);
#pragma unroll
      for (int a = 0; a < TR; ++a) g[a][b] = acc[a][b] * v;
    }
  }
  if (ty == tx) {
#pragma unroll
    for (int a = 0; a < TR; ++a) g[a][a] += 1.0;
  }
  __syncthreads();
  store_tile<NP>(stage, g, ty, tx);
  __syncthreads();
  gj_inverse<NP>(stage, sm, spin, t);
  load_tile<NP>(stage, S, g, ty, tx);
}

// ---- wrap from slice l to l-1 ------------------------------------------------------------------------
// parity : G <- diag(v) E G E^-1 diag(1/v),   v = exp(-sigma lamb h[:, l-1])   (lqmc.py:338-345)
// physics: G <- diag(1/v) E^-1 G E diag(v)                                      (Appendix C)
template <int NP, bool PHYS>
__device__ void wrap_g(double (&g)[RegCfg<NP>::TR][RegCfg<NP>::TR], RegSmem<NP>& sm, const SweepParams& p, int spin, int ty, int tx) {
  constexpr int TR = RegCfg<NP>::TR, S = RegCfg<NP>::S;
  double* stage = sm.stage + spin * NP * S;
  store_tile<NP>(stage, g, ty, tx);
  __syncthreads();
  double acc[TR][TR];
#pragma unroll
  for (int a = 0; a < TR; ++a)
#pragma unroll
    for (int b = 0; b < TR; ++b) acc[a][b] = 0.0;
  gemm_gT_s<NP>(acc, PHYS ? p.Eit : p.Et, stage, ty, tx);
  __syncthreads();
  store_tile<NP>(stage, acc, ty, tx);
  __syncthreads();
#pragma unroll
  for (int a = 0; a < TR; ++a)
#pragma unroll
    for (int b = 0; b < TR; ++b) acc[a][b] = 0.0;
  gemm_s_g<NP>(acc, stage, PHYS ? p.E : p.Ei, ty, tx);
  double rs[TR], cs[TR];
#pragma unroll
  for (int a = 0; a < TR; ++a) {
    const int8_t hr = sm.hn[2 * ty + 16 * (a >> 1) + (a & 1)];
    const int8_t hc = sm.hn[2 * tx + 16 * (a >> 1) + (a & 1)];
    rs[a] = PHYS ? hs_vinv(hr, spin, p) : hs_v(hr, spin, p);
    cs[a] = PHYS ? hs_v(hc, spin, p) : hs_vinv(hc, spin, p);
  }
#pragma unroll
  for (int a = 0; a < TR; ++a)
#pragma unroll
    for (int b = 0; b < TR; ++b) g[a][b] = acc[a][b] * rs[a] * cs[b];
}

// ---- the N proposals of one time slice ---------------------------------------------------------------
// lqmc.py:311-335.  All 128 threads evaluate the same ratio from the same shared-memory numbers, so the
// accept decision is CTA-uniform without communication.  The diagonal is kept lazily: d[] holds G_jj as of
// the previous accepted flip and the current value is d[j] - e[j]*c[j] with that flip's vectors, which is
// exactly the arithmetic the tile owner applies.  One bar.sync per accepted flip, none per rejected one.
template <int NP, bool EXACT, bool PHYS>
__device__ void propose_slice(double (&g)[RegCfg<NP>::TR][RegCfg<NP>::TR], RegSmem<NP>& sm, const SweepParams& p,
                              long long trace_base, int spin, int t, int ty, int tx, int& n_accepted) {
  constexpr int TR = RegCfg<NP>::TR;
  const int N = p.n_sites;
  if (ty == tx) {
#pragma unroll
    for (int a = 0; a < TR; ++a) sm.d[spin * NP + 2 * ty + 16 * (a >> 1) + (a & 1)] = g[a][a];
  }
  if (t < NP) {
    sm.e[spin * NP + t] = 0.0; sm.e[2 * NP + spin * NP + t] = 0.0;
    sm.c[spin * NP + t] = 0.0; sm.c[2 * NP + spin * NP + t] = 0.0;
  }
  __syncthreads();
  int cur = 0;
#pragma unroll
  for (int q = 0; q < TR / 2; ++q) {
#pragma unroll 1
    for (int tyi = 0; tyi < 8; ++tyi) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int i = 16 * q + 2 * tyi + s;
        const int li = 2 * q + s;               // local row / column index of site i in its owners' tiles
        if (i < N) {
          const int8_t h = sm.h[i];
          const double* ec = sm.e + cur * 2 * NP;
          const double* cc = sm.c + cur * 2 * NP;
          const double gu = rank1<EXACT>(sm.d[i], ec[i], cc[i]);
          const double gd = rank1<EXACT>(sm.d[NP + i], ec[NP + i], cc[NP + i]);
          // exp(+arg)-1 for spin up and exp(-arg)-1 for spin down, arg = 2*lamb*h   (lqmc.py:313-315)
          const double fu = (h > 0) ? p.f_p2 : p.f_m2;
          const double fd = (h > 0) ? p.f_m2 : p.f_p2;
          const double du = __dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, gu), fu));
          const double dd = __dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, gd), fd));
          const double ratio = __dmul_rn(du, dd);
          const bool acc = sm.u[i] <= ratio;
          if (p.tr_ratio != nullptr && threadIdx.x == 0) {
            p.tr_ratio[trace_base + i] = ratio;
            p.tr_acc[trace_base + i] = acc ? 1 : 0;
          }
          if (acc) {
            const int nxt = cur ^ 1;
            if (t > i && t < NP) sm.d[spin * NP + t] = rank1<EXACT>(sm.d[spin * NP + t], ec[spin * NP + t], cc[spin * NP + t]);
            double* en = sm.e + nxt * 2 * NP + spin * NP;
            double* cn = sm.c + nxt * 2 * NP + spin * NP;
            const double gs = spin ? gd : gu;
            if (!PHYS) {
              // parity: gamma_up = exp(-arg)-1, gamma_dn = exp(+arg)-1 (lqmc.py:320-323)
              const double gamma = spin ? fu : fd;
              if (ty == tyi) {
                double cv[TR];
#pragma unroll
                for (int b = 0; b < TR; ++b) cv[b] = __dmul_rn(-gamma, g[li][b]);
                if (tx == tyi) cv[li] = __dadd_rn(cv[li], gamma);
#pragma unroll
                for (int b = 0; b < TR; b += 2)
                  *reinterpret_cast<double2*>(cn + 2 * tx + 16 * (b >> 1)) = make_double2(cv[b], cv[b + 1]);
              }
              if (tx == tyi) {
                const double ci = __dadd_rn(__dmul_rn(-gamma, gs), gamma);
                const double den = __dadd_rn(1.0, ci);
                double ev[TR];
                if (EXACT) {
#pragma unroll
                  for (int a = 0; a < TR; ++a) ev[a] = __ddiv_rn(g[a][li], den);
                } else {
                  const double r = __drcp_rn(den);
#pragma unroll
                  for (int a = 0; a < TR; ++a) ev[a] = g[a][li] * r;
                }
#pragma unroll
                for (int a = 0; a < TR; a += 2)
                  *reinterpret_cast<double2*>(en + 2 * ty + 16 * (a >> 1)) = make_double2(ev[a], ev[a + 1]);
              }
            } else {
              // physics: G <- G - (e_i - G[:,i]) (Delta/R) G[i,:],  Delta = exp(2 sigma lamb h) - 1
              const double delta = spin ? fd : fu;
              const double rr = spin ? dd : du;
              if (ty == tyi) {
#pragma unroll
                for (int b = 0; b < TR; b += 2)
                  *reinterpret_cast<double2*>(cn + 2 * tx + 16 * (b >> 1)) = make_double2(g[li][b], g[li][b + 1]);
              }
              if (tx == tyi) {
                const double fac = delta / rr;
                double ev[TR];
#pragma unroll
                for (int a = 0; a < TR; ++a) ev[a] = -g[a][li] * fac;
                if (ty == tyi) ev[li] = (1.0 - g[li][li]) * fac;
#pragma unroll
                for (int a = 0; a < TR; a += 2)
                  *reinterpret_cast<double2*>(en + 2 * ty + 16 * (a >> 1)) = make_double2(ev[a], ev[a + 1]);
              }
            }
            __syncthreads();
            cur = nxt;
            double ev[TR], cv[TR];
#pragma unroll
            for (int a = 0; a < TR; a += 2) {
              const double2 x = *reinterpret_cast<const double2*>(en + 2 * ty + 16 * (a >> 1));
              const double2 y = *reinterpret_cast<const double2*>(cn + 2 * tx + 16 * (a >> 1));
              ev[a] = x.x; ev[a + 1] = x.y; cv[a] = y.x; cv[a + 1] = y.y;
            }
#pragma unroll
            for (int a = 0; a < TR; ++a)
#pragma unroll
              for (int b = 0; b < TR; ++b) g[a][b] = rank1<EXACT>(g[a][b], ev[a], cv[b]);
            if (threadIdx.x == 0) sm.h[i] = -h;
            ++n_accepted;
          }
        }
      }
    }
  }
}

template <int NP, bool EXACT, bool PHYS>
__global__ void __launch_bounds__(128, (NP == 64) ? 3 : 4) sweep_reg_kernel(const SweepParams p) {
  constexpr int TR = RegCfg<NP>::TR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RegSmem<NP> sm(smem_raw);
  const int chain = blockIdx.x;
  const int tid = threadIdx.x;
  const int spin = tid >> 6, t = tid & 63, ty = t >> 3, tx = t & 7;
  const int N = p.n_sites, L = p.n_slices;
  int8_t* field = p.field + (size_t)chain * L * NP;
  double* Gc = p.G + ((size_t)chain * 2 + spin) * NP * NP;
  const int n_steps = p.step_hi - p.step_lo;
  double g[TR][TR];
  int n_accepted = 0;

  if (!p.do_recompute) load_tile<NP>(Gc, NP, g, ty, tx);

  for (int sweep = 0; sweep < p.n_sweeps; ++sweep) {
    if (p.do_recompute) recompute_g<NP>(g, sm, p, field, p.recompute_l0, spin, t, ty, tx);
    for (int step = p.step_lo; step < p.step_hi; ++step) {
      const int l = L - 1 - step;
      const long long base = (((long long)chain * p.n_sweeps + sweep) * n_steps + (step - p.step_lo)) * N;
      if (p.do_propose) {
        __syncthreads();             // previous users of sm.u / sm.h are done
        if (tid < NP) {
          sm.h[tid] = field[l * NP + tid];
          double u = 2.0;
          if (tid < N) {
            u = (p.uniforms != nullptr)
                    ? p.uniforms[base + tid]
                    : lqmc_philox_uniform(p.seed, (uint64_t)(p.chain0 + chain), (uint64_t)(p.sweep0 + sweep), (uint32_t)(step * N + tid));
          }
          sm.u[tid] = u;
        }
        // propose_slice starts with its own barrier after the diagonal / buffer setup
        propose_slice<NP, EXACT, PHYS>(g, sm, p, base, spin, t, ty, tx, n_accepted);
        __syncthreads();
        if (tid < NP) field[l * NP + tid] = sm.h[tid];
      }
      if (p.do_wrap && l > 0) {
        __syncthreads();
        if (tid < NP) sm.hn[tid] = field[(l - 1) * NP + tid];
        // wrap_g's first barrier (after store_tile) also publishes sm.hn
        wrap_g<NP, PHYS>(g, sm, p, spin, ty, tx);
      }
    }
    if (p.measure) {
      double* gs = p.g_sum + ((size_t)chain * 2 + spin) * N * N;
#pragma unroll
      for (int a = 0; a < TR; ++a) {
        const int row = 2 * ty + 16 * (a >> 1) + (a & 1);
#pragma unroll
        for (int b = 0; b < TR; ++b) {
          const int col = 2 * tx + 16 * (b >> 1) + (b & 1);
          if (row < N && col < N) gs[row * N + col] += g[a][b];
        }
      }
      __syncthreads();
      if (ty == tx) {
#pragma unroll
        for (int a = 0; a < TR; ++a) sm.d[spin * NP + 2 * ty + 16 * (a >> 1) + (a & 1)] = g[a][a];
      }
      __syncthreads();
      if (tid < N) {
        const double nu = 1.0 - sm.d[tid], nd = 1.0 - sm.d[NP + tid];
        double* ob = p.obs_sum + (size_t)chain * 3 * N;
        ob[tid] += nu;
        ob[N + tid] += nd;
        ob[2 * N + tid] += nu * nd;
      }
      if (tid == 0) p.n_meas[chain] += 1;
    }
  }
  // hand the Green's functions back (lqmc.py:347) and account the accepted flips
#pragma unroll
  for (int a = 0; a < TR; ++a) {
    const int row = 2 * ty + 16 * (a >> 1) + (a & 1);
#pragma unroll
    for (int q = 0; q < TR / 2; ++q)
      *reinterpret_cast<double2*>(Gc + row * NP + 2 * tx + 16 * q) = make_double2(g[a][2 * q], g[a][2 * q + 1]);
  }
  if (tid == 0 && n_accepted) p.n_acc[chain] += n_accepted;
}

}  // namespace lqmc
