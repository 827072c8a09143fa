// Register-resident sweep kernel: one CTA per Markov chain, N <= 64 sites.
//
// Replaces, for a whole batch of chains at once, LatticeQMC._update_step of the reference
// (/root/reference/lqmc/lqmc.py:301-347) and the helpers it calls (get_m :156-185, get_exp_v :132-154,
// np.linalg.inv :306-307, Configuration.update configuration.py:126-136, measure_loop :356-375).
//
// Layout.  A CTA = 2 spins x (GY x GX) threads.  Each thread owns a TR x TC register tile of its
// spin's Green's function (TR = NP/GY rows, TC = NP/GX columns; NP = N padded to 16/32/64): rows
// {2*ty + 2*GY*q + s}, columns {2*tx + 2*GX*q + s}.  The pairwise interleave makes every shared-memory
// fragment access of the GEMMs, and both the plain and the transposed tile store, a conflict-free (or
// minimum-wavefront) 128-bit access with the row stride S = NP+2, and lets the site loop be unrolled
// so that "the row / column of site i" is a compile-time register index.  G never leaves the register
// file during the N proposals of a slice: a flip publishes row i and column i (2*NP doubles per spin)
// through shared memory, one bar.sync, then every thread applies the rank-1 update to its own tile.
// Per flip that is ~1 KB of shared traffic instead of the 16*N^2 B a shared-memory-resident G would
// move, so the phase is bound by the FP64 pipe (2*N^2 DFMA per flip), not by shared-memory bandwidth.
//
// The wrap G <- B G B^-1 and the sweep-start product are register-tiled DFMA GEMMs with k-major
// operands: one staged in shared memory (the chain's own G / running product, stored transposed when
// it is the left operand), the other (exp(-dtau K) or its inverse, shared by every chain on the SM)
// read through L1 with ld.global.nc.  exp(V_l) is diagonal and is folded into the epilogue as row /
// column scales (lqmc.py:339-345 builds it dense).  FP64 has no tcgen05 kind and DMMA issues to the
// same pipe as DFMA on sm_100a (profiles/fp64_peaks_r01.json: 36.8 vs 37.2 TFLOP/s alone, 30.9 mixed),
// so the roofline is the FP64 pipe either way.
//
// Arithmetic modes (template flags):
//   EXACT  - ratio, rank-1 vectors and update use the reference's roundings: separate multiply and
//            subtract, correctly rounded division (lqmc.py:314-331).  Given the same G, field and
//            uniforms the decisions and the updated G of a slice are bit-identical to NumPy's.
//   !EXACT - update contracted to one FMA, e = column * (1/denominator).
//   PHYS   - textbook DQMC instead of the reference recurrence.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "philox.h"

// 1: propose_slice_delayed (rank-1 updates of a slice delayed, lane-parallel accept scan).  Bit-identical to the immediate path
// and measured SLOWER (cfg2 3.37 vs 3.18 ms per sweep, cfg3 23.7 vs 22.2): both are bound by the per-flip chain of dependent
// instructions at ~5 clocks each with 4 warps per scheduler, not by the tile update (profiles/r02j_*).
#ifndef LQMC_REG_DELAYED
#define LQMC_REG_DELAYED 0
#endif
#ifndef LQMC_REG_PRE
#define LQMC_REG_PRE 0      // next-site ratio evaluated inside the tile-update block: bit-identical, cfg2 3.32 vs 3.15 ms, cfg3 23.3 vs 21.9: off
#endif
#ifndef LQMC_REG_KD
#define LQMC_REG_KD 16
#endif

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
  int wrap_first;       // G arrives one slice above step_lo's (stabilised G(l+1)): wrap it down before the first proposals
  int skip_last_wrap;   // no wrap after step_hi-1 (the caller recomputes G for the next slice: stabilised segments)
  // indexing of the uniforms / trace buffers [chain][buf_sweeps][buf_steps][N]: this launch's sweep s and step t sit at
  // sweep index buf_sweep0 + s and step index t - buf_step0 (a segmented sweep is several launches over one buffer)
  int buf_sweeps, buf_steps, buf_sweep0, buf_step0;
  double exp_pl, exp_ml, f_p2, f_m2;  // exp(+lamb), exp(-lamb), exp(+2 lamb)-1, exp(-2 lamb)-1
};

template <int NP_, int GY_, int GX_>
struct RegCfg {
  static constexpr int NP = NP_, GY = GY_, GX = GX_;
  static constexpr int TR = NP / GY;      // rows per thread
  static constexpr int TC = NP / GX;      // columns per thread
  static constexpr int TPS = GY * GX;     // threads per spin
  static constexpr int THREADS = 2 * TPS;
  static constexpr int WPS = TPS / 32;    // warps per spin
  static constexpr int GMIN = GY < GX ? GY : GX;
  static constexpr int S = NP + 2;        // shared row stride in doubles: even (16-B vectors), odd/2 (banks)
  static constexpr int HS = TPS / NP;     // threads per matrix row in the inverse
  static constexpr int KD = LQMC_REG_KD;  // delay depth of the rank-1 updates inside a slice (propose_slice_delayed)
  static constexpr size_t smem_bytes = (size_t(2) * NP * S + 2 * NP + 8 * NP + NP + 2 * WPS) * sizeof(double)
                                       + (2 * WPS + 2 * NP) * sizeof(int) + 2 * NP
                                       + (LQMC_REG_DELAYED ? (size_t(4) * NP + size_t(4) * (KD + 2) * NP) * sizeof(double) : 0);
  static_assert(TR % 2 == 0 && TC % 2 == 0, "tiles are made of element pairs");
  static_assert(TPS % 32 == 0 && TPS >= NP && TPS % NP == 0, "thread grid");
  static_assert(TPS >= 2 * NP, "a spin group rebuilds row and column of the flipped site with one thread per element");
};

template <class C>
struct RegSmem {
  double* stage;   // [2][NP][S]
  double* d;       // [2][NP]      diagonal as of the previous accepted flip
  double* e;       // [2 buf][2 spin][NP]
  double* c;       // [2 buf][2 spin][NP]
  double* u;       // [NP]
  double* red_v;   // [2*WPS] pivot search partials (per warp)
  int* red_i;      // [2*WPS]
  int* piv;        // [2][NP]
  int8_t* h;       // [NP] field column of the slice being updated
  int8_t* hn;      // [NP] field column the wrap scales with
  double* dh;      // [2 buf][2 spin][NP]     delayed path: diagonal as of the flip before the last one
  double* eh;      // [2 spin][NP][KD+2]      delayed path: e vectors of the pending flips, site-major
  double* ch;      // [2 spin][NP][KD+2]      c vectors
  __device__ explicit RegSmem(unsigned char* base) {
    constexpr int NP = C::NP, S = C::S, WPS = C::WPS;
    stage = reinterpret_cast<double*>(base);
    d = stage + 2 * NP * S;
    e = d + 2 * NP;
    c = e + 4 * NP;
    u = c + 4 * NP;
    red_v = u + NP;
    red_i = reinterpret_cast<int*>(red_v + 2 * WPS);
    piv = red_i + 2 * WPS;
    h = reinterpret_cast<int8_t*>(piv + 2 * NP);
    hn = h + NP;
    dh = reinterpret_cast<double*>(hn + NP);
    eh = dh + 4 * NP;
    ch = eh + 2 * (C::KD + 2) * NP;
  }
};

template <class C> __device__ __forceinline__ int row_of(int ty, int a) { return 2 * ty + 2 * C::GY * (a >> 1) + (a & 1); }
template <class C> __device__ __forceinline__ int col_of(int tx, int b) { return 2 * tx + 2 * C::GX * (b >> 1) + (b & 1); }

template <bool EXACT>
__device__ __forceinline__ double rank1(double g, double e, double c) {
  if (EXACT) return __dsub_rn(g, __dmul_rn(e, c));
  return fma(-e, c, g);
}

// ---- correctly rounded x / d from a shared reciprocal --------------------------------------------------
// IEEE division costs ~50 FP64-pipe issue slots on sm_100a (tools/fp64_peak.cu) and a flip needs N of them
// per spin with one common denominator.  With r = RN(1/d), two residual corrections
//     q <- q + fma(-q, d, x) * r
// give the correctly rounded quotient (Markstein's theorem: r correctly rounded, q faithful after the first
// correction, exact residual from the FMA).  Operands outside a wide safe exponent window take the
// IEEE path.  lqmc_selftest_division() checks bit-equality with __ddiv_rn on the device.
__device__ __forceinline__ bool div_safe(double v) {
  const unsigned ex = ((unsigned)__double2hiint(v) >> 20) & 0x7ffu;
  return ex > 0x3ffu - 400u && ex < 0x3ffu + 400u;
}
__device__ __forceinline__ double div_shared_rcp(double x, double d, double r, bool d_safe) {
  if (!(d_safe && (x == 0.0 || div_safe(x)))) return __ddiv_rn(x, d);
  double q = __dmul_rn(x, r);
  q = fma(fma(-q, d, x), r, q);
  q = fma(fma(-q, d, x), r, q);
  return q;
}

// ---- register-tiled DFMA GEMM, k-major operands ----------------------------------------------------------
// acc[a][b] += sum_k A[k*lda + row_a] * B[k*ldb + col_b];  GA / GB: operand is global (read-only path)
template <class C, bool GA, bool GB>
__device__ __forceinline__ void gemm_kmajor(double (&acc)[C::TR][C::TC], const double* __restrict__ A, int lda,
                                            const double* __restrict__ B, int ldb, int ty, int tx) {
  constexpr int TR = C::TR, TC = C::TC, NP = C::NP;
  const double* ap = A + 2 * ty;
  const double* bp = B + 2 * tx;
#pragma unroll 4
  for (int k = 0; k < NP; ++k) {
    double a[TR], b[TC];
#pragma unroll
    for (int q = 0; q < TR / 2; ++q) {
      const double2* src = reinterpret_cast<const double2*>(ap + k * lda + 2 * C::GY * q);
      const double2 v = GA ? __ldg(src) : *src;
      a[2 * q] = v.x; a[2 * q + 1] = v.y;
    }
#pragma unroll
    for (int q = 0; q < TC / 2; ++q) {
      const double2* src = reinterpret_cast<const double2*>(bp + k * ldb + 2 * C::GX * q);
      const double2 w = GB ? __ldg(src) : *src;
      b[2 * q] = w.x; b[2 * q + 1] = w.y;
    }
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
      for (int j = 0; j < TC; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
  }
}

// FP64 tensor-core MMA, D(8x8) += A(8x4) * B(4x8).  Fragment layout (PTX ISA, mma.m8n8k4 .f64):
//   a : A[row = lane/4][k = lane%4]      b : B[k = lane%4][col = lane/4]
//   c0, c1 : C[row = lane/4][col = 2*(lane%4) + {0,1}]
// On sm_100a DMMA issues to the same FP64 pipe as DFMA (same peak); what it buys is operand traffic: a lane
// loads 1 double per 8 multiply-adds instead of 1 per <= 2.7 for a 4x8 register-tiled DFMA loop, which is the
// difference between a shared-memory-bound and an FP64-pipe-bound GEMM (profiles/r01_cfg4_summary.md).
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}


// ---- NP = 64: the 64 x 64 x 64 GEMMs on DMMA fragments --------------------------------------------------------------
// One spin group = 128 threads = 4 warps in a 2 x 2 grid of 32 x 32 warp tiles; acc[m][n][s] is element
// (32 wm + 8 m + lane/4, 32 wn + 8 n + 2 (lane%4) + s).  A is k-major (A[k*lda + row]), B is row-major by k.
// The register-tiled DFMA loop above needs one shared-memory double per 2.7 FMAs and is LSU-bound (52 % wavefronts at
// 40 % FP64 pipe, profiles/r01_cfg2_summary.md); a DMMA fragment needs one per 8.
template <bool GA, bool GB>
__device__ __forceinline__ void gemm64_dmma(double (&acc)[4][4][2], const double* __restrict__ A, int lda,
                                            const double* __restrict__ B, int ldb, int wm, int wn, int lr, int lk) {
  const double* ap = A + lk * lda + 32 * wm + lr;
  const double* bp = B + lk * ldb + 32 * wn + lr;
#pragma unroll 4
  for (int k4 = 0; k4 < 16; ++k4) {
    double a[4], b[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) a[m] = GA ? __ldg(ap + 4 * k4 * lda + 8 * m) : ap[4 * k4 * lda + 8 * m];
#pragma unroll
    for (int n = 0; n < 4; ++n) b[n] = GB ? __ldg(bp + 4 * k4 * ldb + 8 * n) : bp[4 * k4 * ldb + 8 * n];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int n = 0; n < 4; ++n) dmma884(acc[m][n], a[m], b[n]);
  }
}
__device__ __forceinline__ void zero_acc(double (&acc)[4][4][2]) {
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
}
// fragment tile -> shared, row-major M[row][col] / transposed M[col][row], row stride S; element scaled by rs[m] * cs[n][s]
template <int S, bool TRANSPOSED, bool ADD_ID>
__device__ __forceinline__ void store_acc(double* M, const double (&acc)[4][4][2], const double (&rs)[4], const double (&cs)[4][2],
                                          int wm, int wn, int lr, int lk) {
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int row = 32 * wm + 8 * m + lr;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int col0 = 32 * wn + 8 * n + 2 * lk;
      double v0 = acc[m][n][0] * rs[m] * cs[n][0], v1 = acc[m][n][1] * rs[m] * cs[n][1];
      if (ADD_ID) { if (row == col0) v0 += 1.0; if (row == col0 + 1) v1 += 1.0; }
      if (!TRANSPOSED) {
        *reinterpret_cast<double2*>(M + row * S + col0) = make_double2(v0, v1);
      } else {
        M[col0 * S + row] = v0;
        M[(col0 + 1) * S + row] = v1;
      }
    }
  }
}

// tile -> shared, row-major M[row][col]
template <class C>
__device__ __forceinline__ void store_tile(double* M, const double (&g)[C::TR][C::TC], int ty, int tx) {
#pragma unroll
  for (int a = 0; a < C::TR; ++a) {
    const int row = row_of<C>(ty, a);
#pragma unroll
    for (int q = 0; q < C::TC / 2; ++q)
      *reinterpret_cast<double2*>(M + row * C::S + 2 * tx + 2 * C::GX * q) = make_double2(g[a][2 * q], g[a][2 * q + 1]);
  }
}
// tile -> shared, transposed M[col][row]  (k-major left operand of the next GEMM)
template <class C>
__device__ __forceinline__ void store_tile_t(double* M, const double (&g)[C::TR][C::TC], int ty, int tx) {
#pragma unroll
  for (int b = 0; b < C::TC; ++b) {
    const int col = col_of<C>(tx, b);
#pragma unroll
    for (int q = 0; q < C::TR / 2; ++q)
      *reinterpret_cast<double2*>(M + col * C::S + 2 * ty + 2 * C::GY * q) = make_double2(g[2 * q][b], g[2 * q + 1][b]);
  }
}
template <class C>
__device__ __forceinline__ void load_tile(const double* M, int stride, double (&g)[C::TR][C::TC], int ty, int tx) {
#pragma unroll
  for (int a = 0; a < C::TR; ++a) {
    const int row = row_of<C>(ty, a);
#pragma unroll
    for (int q = 0; q < C::TC / 2; ++q) {
      const double2 v = *reinterpret_cast<const double2*>(M + row * stride + 2 * tx + 2 * C::GX * q);
      g[a][2 * q] = v.x; g[a][2 * q + 1] = v.y;
    }
  }
}
template <class C>
__device__ __forceinline__ void zero_tile(double (&acc)[C::TR][C::TC]) {
#pragma unroll
  for (int a = 0; a < C::TR; ++a)
#pragma unroll
    for (int b = 0; b < C::TC; ++b) acc[a][b] = 0.0;
}
// publish the tile's diagonal entries: dst[row] = g[row][row]
template <class C>
__device__ __forceinline__ void store_diag(double* dst, const double (&g)[C::TR][C::TC], int ty, int tx) {
#pragma unroll
  for (int a = 0; a < C::TR; ++a)
#pragma unroll
    for (int b = 0; b < C::TC; ++b)
      if (row_of<C>(ty, a) == col_of<C>(tx, b)) dst[row_of<C>(ty, a)] = g[a][b];
}

// exp(-sigma*lamb*h) for spin index `spin` (0: sigma=+1, 1: sigma=-1)   [get_exp_v, lqmc.py:149-154]
__device__ __forceinline__ double hs_v(int8_t h, int spin, const SweepParams& p) {
  return ((h > 0) != (spin != 0)) ? p.exp_ml : p.exp_pl;
}
__device__ __forceinline__ double hs_vinv(int8_t h, int spin, const SweepParams& p) {
  return ((h > 0) != (spin != 0)) ? p.exp_pl : p.exp_ml;
}

// ---- in-place Gauss-Jordan inverse with partial (row) pivoting ------------------------------------------
// Stands in for np.linalg.inv (LAPACK getrf/getri, lqmc.py:306-307): same pivot rule (first entry of
// largest magnitude in the column), different elimination order.  HS threads share a matrix row.
template <class C>
__device__ void gj_inverse(double* M, RegSmem<C>& sm, int spin, int t) {
  constexpr int NP = C::NP, S = C::S, WPS = C::WPS, HS = C::HS, SEG = NP / HS;
  const int lane = threadIdx.x & 31, warp_in_spin = (t >> 5);
  const int row = t % NP, seg = t / NP;
  int* piv = sm.piv + spin * NP;
  for (int k = 0; k < NP; ++k) {
    const double akk = M[k * S + k];
    const bool cand = (seg == 0 && row >= k);
    double pv = cand ? M[row * S + k] : 0.0;
    double av = cand ? fabs(pv) : -1.0;
    int idx = cand ? row : NP;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const double oa = __shfl_down_sync(0xffffffffu, av, off);
      const double op = __shfl_down_sync(0xffffffffu, pv, off);
      const int oi = __shfl_down_sync(0xffffffffu, idx, off);
      if (oa > av || (oa == av && oi < idx)) { av = oa; pv = op; idx = oi; }
    }
    if (lane == 0) { sm.red_v[spin * WPS + warp_in_spin] = pv; sm.red_i[spin * WPS + warp_in_spin] = idx; }
    __syncthreads();
    pv = sm.red_v[spin * WPS]; idx = sm.red_i[spin * WPS];
#pragma unroll
    for (int w = 1; w < WPS; ++w) {
      const double ov = sm.red_v[spin * WPS + w];
      const int oi = sm.red_i[spin * WPS + w];
      if (oi < NP && (idx >= NP || fabs(ov) > fabs(pv) || (fabs(ov) == fabs(pv) && oi < idx))) { pv = ov; idx = oi; }
    }
    if (idx >= NP) idx = k;
    // this row's multiplier, taken before anybody rewrites column k: after the swap row `idx` holds the old
    // row k, every other row keeps its own entry
    const double f = (row == idx && idx != k) ? akk : M[row * S + k];
    if (t == 0) piv[k] = idx;
    if (t < NP) {
      const double ak = M[k * S + t];
      const double ap = M[idx * S + t];
      const double x = (t == k) ? 1.0 : ap;
      M[k * S + t] = x / pv;
      if (idx != k) M[idx * S + t] = ak;
    }
    __syncthreads();
    if (row != k) {
      double* rp = M + row * S;
      const double* rk = M + k * S;
#pragma unroll 4
      for (int j = seg * SEG; j < (seg + 1) * SEG; j += 2) {
        double2 r = *reinterpret_cast<double2*>(rp + j);
        const double2 q = *reinterpret_cast<const double2*>(rk + j);
        if (j == (k & ~1)) { if (k & 1) r.y = 0.0; else r.x = 0.0; }
        r.x = fma(-f, q.x, r.x);
        r.y = fma(-f, q.y, r.y);
        *reinterpret_cast<double2*>(rp + j) = r;
      }
    }
    __syncthreads();
  }
  if (seg == 0) {
    double* rp = M + row * S;
    for (int k = NP - 1; k >= 0; --k) {
      const int p = piv[k];
      if (p != k) { const double tmp = rp[k]; rp[k] = rp[p]; rp[p] = tmp; }
    }
  }
  __syncthreads();
}

// ---- sweep-start Green's function: G = inv(I + B_{s0} B_{s1} ... ), slice order (l0-1-m) mod L ---------
// get_m + np.linalg.inv (lqmc.py:156-185,303-307).  The first factor is taken as is (the reference starts
// its left-to-right product from the scalar 1); each later one costs one GEMM with the column scale
// exp(V_l) in the epilogue.
template <class C>
__device__ void recompute_g(double (&g)[C::TR][C::TC], RegSmem<C>& sm, const SweepParams& p, const int8_t* field, int l0,
                            int spin, int t, int ty, int tx) {
  constexpr int TR = C::TR, TC = C::TC, NP = C::NP, S = C::S;
  const int L = p.n_slices;
  double* stage = sm.stage + spin * NP * S;
  int l = (l0 - 1 + L) % L;
  if constexpr (NP == 64) {
    // DMMA chain: the running product stays in shared memory (k-major), the fragment accumulators never become a tile
    const int lane = threadIdx.x & 31, w = t >> 5, wm = w >> 1, wn = w & 1, lr = lane >> 2, lk = lane & 3;
    const double one4[4] = {1.0, 1.0, 1.0, 1.0};
    {
      const int8_t* hl = field + l * NP;
      for (int q = t; q < NP * NP; q += C::TPS) {               // stage[col][row] = E[row][col] * v[col]
        const int col = q >> 6, row = q & 63;
        stage[col * S + row] = p.Et[col * NP + row] * hs_v(hl[col], spin, p);
      }
    }
    double acc[4][4][2];
    for (int m = 1; m < L; ++m) {
      l = (l0 - 1 - m + 2 * L) % L;
      __syncthreads();
      zero_acc(acc);
      gemm64_dmma<false, true>(acc, stage, S, p.E, NP, wm, wn, lr, lk);
      const int8_t* hl = field + l * NP;
      double cs[4][2];
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) cs[n][s2] = hs_v(hl[32 * wn + 8 * n + 2 * lk + s2], spin, p);
      __syncthreads();                                          // every warp is done reading the old product
      if (m < L - 1) store_acc<S, true, false>(stage, acc, one4, cs, wm, wn, lr, lk);
      else store_acc<S, false, true>(stage, acc, one4, cs, wm, wn, lr, lk);
    }
    if (L == 1) {
      // single factor: I + E diag(v), row-major
      __syncthreads();
      const int8_t* hl = field + l * NP;
      for (int q = t; q < NP * NP; q += C::TPS) {
        const int row = q >> 6, col = q & 63;
        stage[row * S + col] = p.E[row * NP + col] * hs_v(hl[col], spin, p) + (row == col ? 1.0 : 0.0);
      }
    }
    __syncthreads();
    gj_inverse<C>(stage, sm, spin, t);
    load_tile<C>(stage, S, g, ty, tx);
    return;
  }
  {
    const int8_t* hl = field + l * NP;
#pragma unroll
    for (int b = 0; b < TC; ++b) {
      const int col = col_of<C>(tx, b);
      const double v = hs_v(hl[col], spin, p);
#pragma unroll
      for (int a = 0; a < TR; ++a) g[a][b] = p.E[row_of<C>(ty, a) * NP + col] * v;
    }
  }
  for (int m = 1; m < L; ++m) {
    l = (l0 - 1 - m + 2 * L) % L;
    __syncthreads();                 // everyone is done reading the previous stage contents
    store_tile_t<C>(stage, g, ty, tx);
    __syncthreads();
    double acc[TR][TC];
    zero_tile<C>(acc);
    gemm_kmajor<C, false, true>(acc, stage, S, p.E, NP, ty, tx);
    const int8_t* hl = field + l * NP;
#pragma unroll
    for (int b = 0; b < TC; ++b) {
      const double v = hs_v(hl[col_of<C>(tx, b)], spin, p);
#pragma unroll
      for (int a = 0; a < TR; ++a) g[a][b] = acc[a][b] * v;
    }
  }
#pragma unroll
  for (int a = 0; a < TR; ++a)
#pragma unroll
    for (int b = 0; b < TC; ++b)
      if (row_of<C>(ty, a) == col_of<C>(tx, b)) g[a][b] += 1.0;
  __syncthreads();
  store_tile<C>(stage, g, ty, tx);
  __syncthreads();
  gj_inverse<C>(stage, sm, spin, t);
  load_tile<C>(stage, S, g, ty, tx);
}

// ---- wrap from slice l to l-1 ------------------------------------------------------------------------
// parity : G <- diag(v) E G E^-1 diag(1/v),   v = exp(-sigma lamb h[:, l-1])   (lqmc.py:338-345)
// physics: G <- diag(1/v) E^-1 G E diag(v)                                      (Appendix C)
template <class C, bool PHYS>
__device__ void wrap_g(double (&g)[C::TR][C::TC], RegSmem<C>& sm, const SweepParams& p, int spin, int ty, int tx) {
  constexpr int TR = C::TR, TC = C::TC, NP = C::NP, S = C::S;
  double* stage = sm.stage + spin * NP * S;
  store_tile<C>(stage, g, ty, tx);
  __syncthreads();
  if constexpr (NP == 64) {
    const int t = threadIdx.x % C::TPS, lane = threadIdx.x & 31, w = t >> 5, wm = w >> 1, wn = w & 1, lr = lane >> 2, lk = lane & 3;
    const double one4[4] = {1.0, 1.0, 1.0, 1.0};
    const double one42[4][2] = {{1.0, 1.0}, {1.0, 1.0}, {1.0, 1.0}, {1.0, 1.0}};
    double acc[4][4][2];
    zero_acc(acc);
    gemm64_dmma<true, false>(acc, PHYS ? p.Eit : p.Et, NP, stage, S, wm, wn, lr, lk);      // T = E G  (or E^-1 G)
    __syncthreads();
    store_acc<S, true, false>(stage, acc, one4, one42, wm, wn, lr, lk);                     // k-major for the second product
    __syncthreads();
    zero_acc(acc);
    gemm64_dmma<false, true>(acc, stage, S, PHYS ? p.E : p.Ei, NP, wm, wn, lr, lk);
    double rs[4], cs[4][2];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int8_t hr = sm.hn[32 * wm + 8 * m + lr];
      rs[m] = PHYS ? hs_vinv(hr, spin, p) : hs_v(hr, spin, p);
    }
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        const int8_t hc = sm.hn[32 * wn + 8 * n + 2 * lk + s2];
        cs[n][s2] = PHYS ? hs_v(hc, spin, p) : hs_vinv(hc, spin, p);
      }
    __syncthreads();
    store_acc<S, false, false>(stage, acc, rs, cs, wm, wn, lr, lk);
    __syncthreads();
    load_tile<C>(stage, S, g, ty, tx);
    return;
  }
  double acc[TR][TC];
  zero_tile<C>(acc);
  gemm_kmajor<C, true, false>(acc, PHYS ? p.Eit : p.Et, NP, stage, S, ty, tx);
  __syncthreads();
  store_tile_t<C>(stage, acc, ty, tx);
  __syncthreads();
  zero_tile<C>(acc);
  gemm_kmajor<C, false, true>(acc, stage, S, PHYS ? p.E : p.Ei, NP, ty, tx);
  double rs[TR], cs[TC];
#pragma unroll
  for (int a = 0; a < TR; ++a) {
    const int8_t hr = sm.hn[row_of<C>(ty, a)];
    rs[a] = PHYS ? hs_vinv(hr, spin, p) : hs_v(hr, spin, p);
  }
#pragma unroll
  for (int b = 0; b < TC; ++b) {
    const int8_t hc = sm.hn[col_of<C>(tx, b)];
    cs[b] = PHYS ? hs_v(hc, spin, p) : hs_vinv(hc, spin, p);
  }
#pragma unroll
  for (int a = 0; a < TR; ++a)
#pragma unroll
    for (int b = 0; b < TC; ++b) g[a][b] = acc[a][b] * rs[a] * cs[b];
}

// ---- the N proposals of one time slice ---------------------------------------------------------------
// lqmc.py:311-335.  Every thread evaluates the same ratio from the same shared-memory numbers, so the
// accept decision is CTA-uniform without communication.  The diagonal is kept lazily: d[] holds G_jj as of
// the previous accepted flip and the current value is d[j] - e[j]*c[j] with that flip's vectors, which is
// exactly the arithmetic the tile owner applies.  One bar.sync per accepted flip, none per rejected one.
template <class C, bool EXACT, bool PHYS>
__device__ void propose_slice(double (&g)[C::TR][C::TC], RegSmem<C>& sm, const SweepParams& p, long long trace_base, int spin,
                              int t, int ty, int tx, int& n_accepted) {
  constexpr int TR = C::TR, TC = C::TC, NP = C::NP, GY = C::GY, GX = C::GX, GMIN = C::GMIN;
  constexpr int RY = GY / GMIN, RX = GX / GMIN;
  const int N = p.n_sites;
  store_diag<C>(sm.d + spin * NP, g, ty, tx);
  if (t < NP) {
    sm.e[spin * NP + t] = 0.0; sm.e[2 * NP + spin * NP + t] = 0.0;
    sm.c[spin * NP + t] = 0.0; sm.c[2 * NP + spin * NP + t] = 0.0;
  }
  __syncthreads();
  int cur = 0;
#if LQMC_REG_PRE
  // ratio of the site after the last accepted flip, evaluated INSIDE that flip's tile-update block (see below)
  int pre_site = -1;
  double pre_gu = 0.0, pre_gd = 0.0, pre_ratio = 0.0;
#endif
#pragma unroll
  for (int j = 0; j < NP / (2 * GMIN); ++j) {
#pragma unroll 1
    for (int m = 0; m < GMIN; ++m) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int i = 2 * GMIN * j + 2 * m + s;
        const int ty_i = GMIN * (j % RY) + m, tx_i = GMIN * (j % RX) + m;   // owners of row i / column i
        const int ai = 2 * (j / RY) + s, bi = 2 * (j / RX) + s;            // their local row / column index
        if (i < N) {
          const int8_t h = sm.h[i];
          const double* ec = sm.e + cur * 2 * NP;
          const double* cc = sm.c + cur * 2 * NP;
          // exp(+arg)-1 for spin up and exp(-arg)-1 for spin down, arg = 2*lamb*h   (lqmc.py:313-315)
          const double fu = (h > 0) ? p.f_p2 : p.f_m2;
          const double fd = (h > 0) ? p.f_m2 : p.f_p2;
#if LQMC_REG_PRE
          double gu, gd, ratio, du = 0.0, dd = 0.0;
          if (!PHYS && pre_site == i) {
            gu = pre_gu; gd = pre_gd; ratio = pre_ratio;
          } else {
            gu = rank1<EXACT>(sm.d[i], ec[i], cc[i]);
            gd = rank1<EXACT>(sm.d[NP + i], ec[NP + i], cc[NP + i]);
            du = __dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, gu), fu));
            dd = __dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, gd), fd));
            ratio = __dmul_rn(du, dd);
          }
#else
          const double gu = rank1<EXACT>(sm.d[i], ec[i], cc[i]);
          const double gd = rank1<EXACT>(sm.d[NP + i], ec[NP + i], cc[NP + i]);
          const double du = __dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, gu), fu));
          const double dd = __dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, gd), fd));
          const double ratio = __dmul_rn(du, dd);
#endif
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
              if (ty == ty_i) {
                double cv[TC];
#pragma unroll
                for (int b = 0; b < TC; ++b) cv[b] = __dmul_rn(-gamma, g[ai][b]);
                if (tx == tx_i) cv[bi] = __dadd_rn(cv[bi], gamma);
#pragma unroll
                for (int b = 0; b < TC; b += 2)
                  *reinterpret_cast<double2*>(cn + 2 * tx + 2 * GX * (b >> 1)) = make_double2(cv[b], cv[b + 1]);
              }
              if (tx == tx_i) {
                const double ci = __dadd_rn(__dmul_rn(-gamma, gs), gamma);
                const double den = __dadd_rn(1.0, ci);
                const double r = __drcp_rn(den);
                double ev[TR];
                if (EXACT) {
                  const bool ok = div_safe(den);
#pragma unroll
                  for (int a = 0; a < TR; ++a) ev[a] = div_shared_rcp(g[a][bi], den, r, ok);
                } else {
#pragma unroll
                  for (int a = 0; a < TR; ++a) ev[a] = g[a][bi] * r;
                }
#pragma unroll
                for (int a = 0; a < TR; a += 2)
                  *reinterpret_cast<double2*>(en + 2 * ty + 2 * GY * (a >> 1)) = make_double2(ev[a], ev[a + 1]);
              }
            } else {
              // physics: G <- G - (e_i - G[:,i]) (Delta/R) G[i,:],  Delta = exp(2 sigma lamb h) - 1
              const double delta = spin ? fd : fu;
              const double rr = spin ? dd : du;
              if (ty == ty_i) {
#pragma unroll
                for (int b = 0; b < TC; b += 2)
                  *reinterpret_cast<double2*>(cn + 2 * tx + 2 * GX * (b >> 1)) = make_double2(g[ai][b], g[ai][b + 1]);
              }
              if (tx == tx_i) {
                const double fac = delta / rr;
                double ev[TR];
#pragma unroll
                for (int a = 0; a < TR; ++a) ev[a] = -g[a][bi] * fac;
                if (ty == ty_i) ev[ai] = (1.0 - g[ai][bi]) * fac;
#pragma unroll
                for (int a = 0; a < TR; a += 2)
                  *reinterpret_cast<double2*>(en + 2 * ty + 2 * GY * (a >> 1)) = make_double2(ev[a], ev[a + 1]);
              }
            }
            __syncthreads();
            cur = nxt;
#if LQMC_REG_PRE
            // A rejected proposal changes nothing, so the ratio of the NEXT site is already determined here.  Evaluated in the same
            // basic block as the tile update, its chain of dependent loads and FP64 operations is scheduled between the tile's
            // independent multiply-subtracts instead of stalling the warp at the top of the next iteration (index clamped: no branch).
            if (!PHYS) {
              const int ip = (i + 1 < NP) ? i + 1 : NP - 1;
              const int8_t hp = sm.h[ip];
              pre_gu = rank1<EXACT>(sm.d[ip], en[ip - spin * NP], cn[ip - spin * NP]);
              pre_gd = rank1<EXACT>(sm.d[NP + ip], en[NP + ip - spin * NP], cn[NP + ip - spin * NP]);
              const double fup = (hp > 0) ? p.f_p2 : p.f_m2;
              const double fdp = (hp > 0) ? p.f_m2 : p.f_p2;
              pre_ratio = __dmul_rn(__dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, pre_gu), fup)), __dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, pre_gd), fdp)));
              pre_site = i + 1;
            }
#endif
            double ev[TR], cv[TC];
#pragma unroll
            for (int a = 0; a < TR; a += 2) {
              const double2 x = *reinterpret_cast<const double2*>(en + 2 * ty + 2 * GY * (a >> 1));
              ev[a] = x.x; ev[a + 1] = x.y;
            }
#pragma unroll
            for (int b = 0; b < TC; b += 2) {
              const double2 y = *reinterpret_cast<const double2*>(cn + 2 * tx + 2 * GX * (b >> 1));
              cv[b] = y.x; cv[b + 1] = y.y;
            }
#pragma unroll
            for (int a = 0; a < TR; ++a)
#pragma unroll
              for (int b = 0; b < TC; ++b) g[a][b] = rank1<EXACT>(g[a][b], ev[a], cv[b]);
            if (threadIdx.x == 0) sm.h[i] = -h;
            ++n_accepted;
          }
        }
      }
    }
  }
}

// ---- the N proposals of one time slice, rank-1 updates delayed ------------------------------------------
// Same arithmetic as propose_slice, different schedule.  Applying every accepted flip to the register tiles right
// away puts a publish -> bar.sync -> 2*TR*TC FP64 instructions chain between two flips (2.3 K clocks per flip with
// two CTAs per SM).  Here the (e, c) vectors of up to KD flips stay in shared memory, the register tiles keep the
// Green's function G0 of the last flush (also mirrored row-major in `stage`), and a flip only needs row i and
// column i of the CURRENT G:
//     G[i][j] = (((G0[i][j] - e_0[i] c_0[j]) - e_1[i] c_1[j]) - ...)        one thread per element, 2*NP per spin
// which is exactly the sequence of roundings the element would have seen in the tile (separate multiply and
// subtract in exact arithmetic), so decisions, vectors and the flushed G are bit-identical to the undelayed path.
// The flush applies the pending updates to the tiles back to back (no barrier in between: FP64-pipe-bound).
//   * History layout: site-major, eh[site][m] / ch[site][m] with row stride HSTR = KD + 2 doubles: a thread's chain
//     operands are contiguous (one 128-bit load per two updates, conflict-free at that stride), unused slots of a
//     group of 8 are +0 (e = c = +0 changes no bit), so the chain is straight-line code of 8 or 16 updates that the
//     compiler interleaves with the rest of the flip's arithmetic.
//   * The accept scan is lane-parallel: a warp evaluates the ratios of 32 sites at once from the lazily updated
//     diagonal (d as of the flip before the last one, folded with the last flip's vectors, both kept compact in
//     sm.e / sm.c), the ballot finds the first accepted site, everything before it is rejected for free; every warp
//     does this redundantly, so the decision is CTA-uniform without communication.  Each lane also forms the flip's
//     denominator and its reciprocal for ITS site (SIMD: no more instructions than for one site), which takes the
//     reciprocal off the post-accept critical path.
template <class C>
struct RegHist {
  static constexpr int KD = C::KD, HSTR = KD + 2;
  static_assert(KD == 8 || KD == 16, "chains of 8 or 16 updates");
};

template <int STEPS, bool EXACT>
__device__ __forceinline__ double pending_chain(double x, const double* __restrict__ ea, const double* __restrict__ ca) {
#pragma unroll
  for (int m = 0; m < STEPS; m += 2) {
    const double2 e2 = *reinterpret_cast<const double2*>(ea + m);
    const double2 c2 = *reinterpret_cast<const double2*>(ca + m);
    x = rank1<EXACT>(x, e2.x, c2.x);
    x = rank1<EXACT>(x, e2.y, c2.y);
  }
  return x;
}

template <class C, bool EXACT>
__device__ __forceinline__ void apply_pending(double (&g)[C::TR][C::TC], const double* __restrict__ eh, const double* __restrict__ ch,
                                              int k, int ty, int tx) {
  constexpr int TR = C::TR, TC = C::TC, HSTR = RegHist<C>::HSTR;
  const double* er[TR];
  const double* cr[TC];
#pragma unroll
  for (int a = 0; a < TR; ++a) er[a] = eh + row_of<C>(ty, a) * HSTR;
#pragma unroll
  for (int b = 0; b < TC; ++b) cr[b] = ch + col_of<C>(tx, b) * HSTR;
  for (int m = 0; m < k; m += 2) {          // k odd: slot k holds +0
    double2 e2[TR], c2[TC];
#pragma unroll
    for (int a = 0; a < TR; ++a) e2[a] = *reinterpret_cast<const double2*>(er[a] + m);
#pragma unroll
    for (int b = 0; b < TC; ++b) c2[b] = *reinterpret_cast<const double2*>(cr[b] + m);
#pragma unroll
    for (int a = 0; a < TR; ++a)
#pragma unroll
      for (int b = 0; b < TC; ++b) g[a][b] = rank1<EXACT>(g[a][b], e2[a].x, c2[b].x);
#pragma unroll
    for (int a = 0; a < TR; ++a)
#pragma unroll
      for (int b = 0; b < TC; ++b) g[a][b] = rank1<EXACT>(g[a][b], e2[a].y, c2[b].y);
  }
}

#ifdef LQMC_PHASE_CLOCKS
struct RegClocks { long long rec, slice, wrap, scan, build, flush, flips, scans, b_load, b_chain, b_vec, b_bar; };
#define RCLK(...) __VA_ARGS__
#else
#define RCLK(...)
#endif
template <class C, bool EXACT, bool PHYS>
__device__ void propose_slice_delayed(double (&g)[C::TR][C::TC], RegSmem<C>& sm, const SweepParams& p, long long trace_base,
                                      int spin, int t, int ty, int tx, int& n_accepted
#ifdef LQMC_PHASE_CLOCKS
                                      , RegClocks& ck
#endif
                                      ) {
  constexpr int NP = C::NP, S = C::S, KD = C::KD, HSTR = RegHist<C>::HSTR;
  const int N = p.n_sites;
  const int lane = threadIdx.x & 31;
  double* stage = sm.stage + spin * NP * S;
  double* eh = sm.eh + spin * NP * HSTR;
  double* ch = sm.ch + spin * NP * HSTR;
  store_tile<C>(stage, g, ty, tx);
  store_diag<C>(sm.dh + spin * NP, g, ty, tx);
  if (t < NP) { sm.e[spin * NP + t] = 0.0; sm.c[spin * NP + t] = 0.0; }
  __syncthreads();
  // cur: buffer of the diagonal (as of the flip before the last one) and of the last flip's compact (e, c)
  int cur = 0, k = 0, i0 = 0;
  while (i0 < N) {
    // ---- scan: ratios of sites i0 .. i0+31 (lqmc.py:313-317), and this spin's denominator / reciprocal ----
    RCLK(const long long c0 = clock64(); ck.scans += 1;)
    const int i = i0 + lane;
    bool acc = false;
    double ratio = 0.0, den = 1.0, rcp = 1.0;       // den / rcp: parity 1 + c_i and its reciprocal; physics: Delta / R in rcp
    int8_t h = 1;
    if (i < N) {
      const double* dc = sm.dh + cur * 2 * NP;
      const double* el = sm.e + cur * 2 * NP;
      const double* cl = sm.c + cur * 2 * NP;
      h = sm.h[i];
      const double gu = rank1<EXACT>(dc[i], el[i], cl[i]);
      const double gd = rank1<EXACT>(dc[NP + i], el[NP + i], cl[NP + i]);
      const double fu = (h > 0) ? p.f_p2 : p.f_m2;
      const double fd = (h > 0) ? p.f_m2 : p.f_p2;
      const double du = __dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, gu), fu));
      const double dd = __dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, gd), fd));
      ratio = __dmul_rn(du, dd);
      acc = sm.u[i] <= ratio;
      if (!PHYS) {
        // parity: gamma_up = exp(-arg)-1, gamma_dn = exp(+arg)-1 (lqmc.py:320-323); denominator 1 + c_i
        const double gamma = spin ? fu : fd;
        den = __dadd_rn(1.0, __dadd_rn(__dmul_rn(-gamma, spin ? gd : gu), gamma));
        rcp = __drcp_rn(den);
      } else {
        rcp = (spin ? fd : fu) / (spin ? dd : du);
      }
    }
    const unsigned ball = __ballot_sync(0xffffffffu, acc);
    const int nf = ball ? (__ffs(ball) - 1) : 31;
    if (p.tr_ratio != nullptr && threadIdx.x < 32 && i < N && lane <= nf) {
      p.tr_ratio[trace_base + i] = ratio;
      p.tr_acc[trace_base + i] = acc ? 1 : 0;
    }
    if (!ball) { i0 += 32; RCLK(ck.scan += clock64() - c0;) continue; }
    RCLK(const long long c1 = clock64(); ck.scan += c1 - c0;)
    const int is = i0 + nf;
    const int8_t hs = (int8_t)__shfl_sync(0xffffffffu, (int)h, nf);
    den = __shfl_sync(0xffffffffu, den, nf);
    rcp = __shfl_sync(0xffffffffu, rcp, nf);
    // ---- build: row / column `is` of the current G, the flip's vectors into history slot k ----
    const int nxt = cur ^ 1;
    if (t < 2 * NP) {
      const bool isrow = t < NP;
      const int j = isrow ? t : t - NP;
      double x = isrow ? stage[is * S + j] : stage[j * S + is];
      const double* ea = eh + (isrow ? is : j) * HSTR;
      const double* ca = ch + (isrow ? j : is) * HSTR;
      RCLK(long long cb0 = clock64() + (__double_as_longlong(x) & 0); ck.b_load += cb0 - c1;)
      if (k > 0) {
        if (KD == 8 || k <= 8) x = pending_chain<8, EXACT>(x, ea, ca);
        else x = pending_chain<KD, EXACT>(x, ea, ca);
      }
      RCLK(long long cb1 = clock64() + (__double_as_longlong(x) & 0); ck.b_chain += cb1 - cb0;)
      double v;
      if (!PHYS) {
        if (isrow) {
          const double gamma = ((hs > 0) != (spin != 0)) ? p.f_m2 : p.f_p2;     // spin ? fu : fd
          v = __dmul_rn(-gamma, x);
          if (j == is) v = __dadd_rn(v, gamma);
        } else {
          v = EXACT ? div_shared_rcp(x, den, rcp, div_safe(den)) : x * rcp;
        }
      } else {
        // physics: G <- G - (e_i - G[:,i]) (Delta/R) G[i,:],  Delta = exp(2 sigma lamb h) - 1
        v = isrow ? x : ((j == is) ? (1.0 - x) * rcp : -x * rcp);
      }
      double* hv = (isrow ? ch : eh) + j * HSTR + k;
      *hv = v;
      (isrow ? sm.c : sm.e)[(nxt * 2 + spin) * NP + j] = v;
      if ((k & 7) == 0) {    // a new group of 8 slots: the unused ones read as +0
        hv[1] = 0.0;
#pragma unroll
        for (int q = 2; q < 8; q += 2) *reinterpret_cast<double2*>(hv + q) = make_double2(0.0, 0.0);
      }
      if (isrow)   // the diagonal as of the previous flip (the new flip's vectors become the lazy part)
        sm.dh[(nxt * 2 + spin) * NP + j] = rank1<EXACT>(sm.dh[(cur * 2 + spin) * NP + j], sm.e[(cur * 2 + spin) * NP + j],
                                                        sm.c[(cur * 2 + spin) * NP + j]);
    }
    RCLK(const long long cb2 = clock64(); ck.b_vec += cb2 - c1;)
    __syncthreads();
    RCLK(ck.b_bar += clock64() - cb2;)
    if (threadIdx.x == 0) sm.h[is] = -hs;   // after the barrier: slower warps were still scanning h[is]
    ++n_accepted;
    ++k; cur = nxt; i0 = is + 1;
    RCLK(const long long c2 = clock64(); ck.build += c2 - c1; ck.flips += 1;)
    if (k == KD && i0 < N) {
      // ---- flush: pending updates into the tiles, G0 mirror and diagonal brought up to date ----
      apply_pending<C, EXACT>(g, eh, ch, k, ty, tx);
      if (t < NP) {
        const int o = (cur * 2 + spin) * NP + t, on = ((cur ^ 1) * 2 + spin) * NP + t;
        sm.dh[on] = rank1<EXACT>(sm.dh[o], sm.e[o], sm.c[o]);
        sm.e[on] = 0.0; sm.c[on] = 0.0;
      }
      store_tile<C>(stage, g, ty, tx);
      __syncthreads();
      cur ^= 1; k = 0;
      RCLK(ck.flush += clock64() - c2;)
    }
  }
  RCLK(const long long c3 = clock64();)
  apply_pending<C, EXACT>(g, eh, ch, k, ty, tx);
  RCLK(ck.flush += clock64() - c3;)
}


template <class C, bool EXACT, bool PHYS>
__global__ void __launch_bounds__(C::THREADS, (C::THREADS == 256) ? 2 : 4) sweep_reg_kernel(const SweepParams p) {
  constexpr int TR = C::TR, TC = C::TC, NP = C::NP, GY = C::GY, TPS = C::TPS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RegSmem<C> sm(smem_raw);
  const int chain = blockIdx.x;
  const int tid = threadIdx.x;
  const int spin = tid / TPS, t = tid % TPS, ty = t % GY, tx = t / GY;
  const int N = p.n_sites, L = p.n_slices;
  int8_t* field = p.field + (size_t)chain * L * NP;
  double* Gc = p.G + ((size_t)chain * 2 + spin) * NP * NP;
  double g[TR][TC];
  int n_accepted = 0;
  RCLK(RegClocks ck = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long cka;)

  if (!p.do_recompute) load_tile<C>(Gc, NP, g, ty, tx);

  for (int sweep = 0; sweep < p.n_sweeps; ++sweep) {
    RCLK(cka = clock64();)
    if (p.do_recompute) recompute_g<C>(g, sm, p, field, p.recompute_l0, spin, t, ty, tx);
    RCLK(ck.rec += clock64() - cka;)
    for (int step = p.step_lo; step < p.step_hi; ++step) {
      const int l = L - 1 - step;
      const long long base = (((long long)chain * p.buf_sweeps + p.buf_sweep0 + sweep) * p.buf_steps + (step - p.buf_step0)) * N;
      if (p.wrap_first && step == p.step_lo) {
        __syncthreads();
        if (tid < NP) sm.hn[tid] = field[l * NP + tid];
        wrap_g<C, PHYS>(g, sm, p, spin, ty, tx);
      }
      if (p.do_propose) {
        RCLK(cka = clock64();)
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
#if LQMC_REG_DELAYED
        propose_slice_delayed<C, EXACT, PHYS>(g, sm, p, base, spin, t, ty, tx, n_accepted
#ifdef LQMC_PHASE_CLOCKS
                                              , ck
#endif
                                              );
#else
        propose_slice<C, EXACT, PHYS>(g, sm, p, base, spin, t, ty, tx, n_accepted);
#endif
        __syncthreads();
        if (tid < NP) field[l * NP + tid] = sm.h[tid];
        RCLK(ck.slice += clock64() - cka;)
      }
      if (p.do_wrap && l > 0 && !(p.skip_last_wrap && step == p.step_hi - 1)) {
        RCLK(cka = clock64();)
        __syncthreads();
        if (tid < NP) sm.hn[tid] = field[(l - 1) * NP + tid];
        // wrap_g's first barrier (after store_tile) also publishes sm.hn
        wrap_g<C, PHYS>(g, sm, p, spin, ty, tx);
        RCLK(ck.wrap += clock64() - cka;)
      }
    }
    if (p.measure) {
      double* gs = p.g_sum + ((size_t)chain * 2 + spin) * N * N;
#pragma unroll
      for (int a = 0; a < TR; ++a) {
        const int row = row_of<C>(ty, a);
#pragma unroll
        for (int b = 0; b < TC; ++b) {
          const int col = col_of<C>(tx, b);
          if (row < N && col < N) gs[row * N + col] += g[a][b];
        }
      }
      __syncthreads();
      store_diag<C>(sm.d + spin * NP, g, ty, tx);
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
    const int row = row_of<C>(ty, a);
#pragma unroll
    for (int q = 0; q < TC / 2; ++q)
      *reinterpret_cast<double2*>(Gc + row * NP + 2 * tx + 2 * C::GX * q) = make_double2(g[a][2 * q], g[a][2 * q + 1]);
  }
  if (tid == 0 && n_accepted) p.n_acc[chain] += n_accepted;
#ifdef LQMC_PHASE_CLOCKS
  if (tid == 0 && N >= 8) {     // tools/reg_split.py reads these back through lqmc_get_measurements
    double* ob = p.obs_sum + (size_t)chain * 3 * N;
    ob[0] = (double)ck.rec; ob[1] = (double)ck.slice; ob[2] = (double)ck.wrap; ob[3] = (double)ck.scan;
    ob[4] = (double)ck.build; ob[5] = (double)ck.flush; ob[6] = (double)ck.flips; ob[7] = (double)ck.scans;
    ob[8] = (double)ck.b_load; ob[9] = (double)ck.b_chain; ob[10] = (double)ck.b_vec; ob[11] = (double)ck.b_bar;
  }
  if (tid == NP && N >= 16) {   // a column-role thread
    double* ob = p.obs_sum + (size_t)chain * 3 * N;
    ob[12] = (double)ck.b_load; ob[13] = (double)ck.b_chain; ob[14] = (double)ck.b_vec; ob[15] = (double)ck.b_bar;
  }
#endif
}

// ---- device self-test of the shared-reciprocal division ---------------------------------------------------
__global__ void division_selftest_kernel(unsigned long long n_per_thread, uint64_t seed, unsigned long long* mismatches) {
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  unsigned long long bad = 0;
  for (unsigned long long it = 0; it < n_per_thread; ++it) {
    uint32_t c[4] = {(uint32_t)it, (uint32_t)(it >> 32), (uint32_t)tid, (uint32_t)(tid >> 32)};
    lqmc_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    // mantissas fully random; exponents within +-300 of 1; every 8th denominator gets an all-ones or
    // one-hot mantissa tail (the hard cases for reciprocal-based division)
    uint64_t mx = ((uint64_t)c[0] << 20) ^ c[1];
    uint64_t md = ((uint64_t)c[2] << 20) ^ c[3];
    mx &= 0x000fffffffffffffull; md &= 0x000fffffffffffffull;
    if ((it & 7) == 1) md |= 0x000ffffffffff000ull;
    if ((it & 7) == 2) md &= 0x000ff00000000001ull;
    if ((it & 7) == 3) mx |= 0x000fffffffffff00ull;
    const int ex = (int)(c[0] % 601u) - 300, ed = (int)(c[3] % 601u) - 300;
    const uint64_t sx = (uint64_t)(c[1] & 1u) << 63, sd = (uint64_t)(c[2] & 1u) << 63;
    const double x = __longlong_as_double((long long)(sx | ((uint64_t)(1023 + ex) << 52) | mx));
    const double d = __longlong_as_double((long long)(sd | ((uint64_t)(1023 + ed) << 52) | md));
    const double ref = __ddiv_rn(x, d);
    const double got = div_shared_rcp(x, d, __drcp_rn(d), div_safe(d));
    if (__double_as_longlong(ref) != __double_as_longlong(got)) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}

}  // namespace lqmc
