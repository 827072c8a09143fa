// Numerically stabilised Green's function for the HS-field sweep: batched Householder-QR / UDV kernels.
//
// The reference recomputes G = inv(I + B_{L-1} ... B_0) once per sweep as a plain matrix product followed by
// np.linalg.inv (/root/reference/lqmc/lqmc.py:156-185,303-307).  At beta >= 8 that product has a condition number
// of 1e21..1e24 and the result is roundoff.  Physics mode replaces it by the standard DQMC
// stabilisation: the product is accumulated `chunk` factors at a time as
//        A = U D V,      U orthogonal, D positive diagonal (graded), V well conditioned
// with a column-norm pre-pivoted Householder QR after every chunk, and
//        G = (D_b^-1 U^T + D_s V)^-1 D_b^-1 U^T,       D = D_b D_s split at 1.
//
// One CTA (256 threads) per (chain, spin) matrix; everything lives in HBM / L2 except the QR panel.
//   st_chain_kernel   M = B_s ... B_s' (U D): `count` GEMMs (DMMA m8n8k4, operand ring as in sweep_l2.cuh) with the
//                     diagonal exp(V_l) as a row scale and D as a column scale in the epilogues
//   st_qr_kernel      column norms -> permutation -> blocked Householder QR (panel of NB columns factored in shared
//                     memory, compact-WY trailing update streamed over the rest) -> D = |diag R| -> explicit Q
//   st_v_kernel       V <- (D^-1 R) P^T V, one GEMM whose right operand rows are gathered through the permutation
//   st_forms / st_inverse / st_final     the G formula above
// The two-sided variant (st_combine_*) joins a stack of left products with the running right product so that a
// stabilisation point costs O(1) QR factorizations instead of O(L / chunk).
#pragma once
#include "sweep_l2.cuh"

namespace lqmc {

constexpr int ST_THREADS = 256;
constexpr int ST_BM = 64, ST_BK = 16, ST_STAGES = 3;
constexpr int ST_LDA = ST_BM + 4;
constexpr int ST_NB = 16;                 // QR panel width
constexpr int ST_CB = 256;                // columns per trailing-update block (2 per thread, thread pairs split the reflectors)
constexpr int ST_LDS = 36;                // DMMA reflector application: row stride of a warp's private 16 x 32 W strip (72 words = 8 mod 32)
constexpr int ST_WT_LD = (ST_THREADS / 32) * ST_LDS > ST_CB ? (ST_THREADS / 32) * ST_LDS : ST_CB;
#ifndef LQMC_ST_APPLY_DMMA
#define LQMC_ST_APPLY_DMMA 1              // 0: the DFMA register-tiled block-reflector application (kept as the timed comparison)
#endif

struct HsConsts { double exp_pl, exp_ml; };

// exp(-sigma lamb h) (inv = false) or its inverse, sigma = +1 for spin index 0   [get_exp_v, lqmc.py:149-154]
__device__ __forceinline__ double st_hs(int8_t h, int spin, bool inv, const HsConsts& c) {
  const bool minus = ((h > 0) != (spin != 0));
  return (minus != inv) ? c.exp_ml : c.exp_pl;
}

inline int st_padded_size(int n_sites) { return n_sites <= 64 ? 64 : l2_padded_size(n_sites); }   // the sweep kernels' layout above 64

// ---- GEMM  C = A * B,  A given k-major, B row-major (rows optionally gathered through bperm) ----------------------
// how a scale vector enters: the value itself, 1 / max(v, 1)  (D_b^-1)  or  min(v, 1)  (D_s)
enum { ST_VEC_PLAIN = 0, ST_VEC_INV_BIG = 1, ST_VEC_SMALL = 2 };
__device__ __forceinline__ double st_vscale(double v, int mode) {
  if (mode == ST_VEC_INV_BIG) return 1.0 / fmax(v, 1.0);
  if (mode == ST_VEC_SMALL) return fmin(v, 1.0);
  return v;
}

struct StEpilogue {
  const int8_t* hrow = nullptr;   // row scale exp(-+sigma lamb h[row]) for row < nvalid
  bool row_inv = false;
  const double* rvec = nullptr;   // row scale by a vector (of the product A*B, whatever the store layout)
  const double* cvec = nullptr;   // column scale by a vector
  int rmode = ST_VEC_PLAIN, cmode = ST_VEC_PLAIN;
  const double* addend = nullptr; // added element-wise, indexed like the output (may be the output itself)
  bool transposed_out = false;
  int nvalid = 0;
};

template <int NFRAG>
__device__ void st_gemm(const double* __restrict__ At, const double* __restrict__ B, const int* __restrict__ bperm,
                        double* __restrict__ Cout, int NP, int spin, const StEpilogue& ep, const HsConsts& hc, double* pa,
                        double* pb) {
  constexpr int BN = 32 * NFRAG, LDB = BN + 4;
  constexpr int BSH = (NFRAG == 4) ? 6 : 5, BQ = (NFRAG == 4) ? 4 : 2, BSTEP = (NFRAG == 4) ? 4 : 8;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const int lr = lane >> 2, lk = lane & 3;
  const int nk = NP / ST_BK;
  const int brow = tid >> BSH, bchunk = tid & ((1 << BSH) - 1);
  __syncthreads();
  for (int i0 = 0; i0 < NP; i0 += ST_BM) {
    for (int j0 = 0; j0 < NP; j0 += BN) {
      double acc[4][NFRAG][2];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < NFRAG; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
      const double* srcA = At + (size_t)(tid >> 5) * NP + i0 + 2 * (tid & 31);
      double* dstA = pa + (tid >> 5) * ST_LDA + 2 * (tid & 31);
      double* dstB = pb + brow * LDB + 2 * bchunk;
      const bool b_ok = j0 + 2 * bchunk < NP;               // half tile at the right edge (NP = 576): no copy, no compute, no store
      const bool w_ok = j0 + 8 * NFRAG * wn < NP;           // warp-uniform
      auto issue = [&](int kpanel, int stage) {
        const int k0 = kpanel * ST_BK;
#pragma unroll
        for (int q = 0; q < 2; ++q)
          __pipeline_memcpy_async(dstA + stage * ST_BK * ST_LDA + q * 8 * ST_LDA, srcA + (size_t)(k0 + q * 8) * NP, 16);
        if (b_ok) {
#pragma unroll
          for (int q = 0; q < BQ; ++q) {
            const int krow = k0 + brow + q * BSTEP;
            const int srow = bperm ? bperm[krow] : krow;
            __pipeline_memcpy_async(dstB + stage * ST_BK * LDB + q * BSTEP * LDB, B + (size_t)srow * NP + j0 + 2 * bchunk, 16);
          }
        }
      };
#pragma unroll
      for (int s0 = 0; s0 < ST_STAGES - 1; ++s0) {
        if (s0 < nk) issue(s0, s0);
        __pipeline_commit();
      }
      for (int kp = 0; kp < nk; ++kp) {
        const int st = kp % ST_STAGES;
        __pipeline_wait_prior(ST_STAGES - 2);
        __syncthreads();
        if (kp + ST_STAGES - 1 < nk) issue(kp + ST_STAGES - 1, (kp + ST_STAGES - 1) % ST_STAGES);
        __pipeline_commit();
        const double* ap = pa + st * ST_BK * ST_LDA + lk * ST_LDA + 32 * wm + lr;
        const double* bp = pb + st * ST_BK * LDB + lk * LDB + 8 * NFRAG * wn + lr;
        if (w_ok) {
#pragma unroll
          for (int k4 = 0; k4 < ST_BK / 4; ++k4) {
            double a[4], b[NFRAG];
#pragma unroll
            for (int m = 0; m < 4; ++m) a[m] = ap[4 * k4 * ST_LDA + 8 * m];
#pragma unroll
            for (int n = 0; n < NFRAG; ++n) b[n] = bp[4 * k4 * LDB + 8 * n];
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
              for (int n = 0; n < NFRAG; ++n) dmma884(acc[m][n], a[m], b[n]);
          }
        }
      }
      __pipeline_wait_prior(0);
      __syncthreads();
      if (!w_ok) continue;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int row = i0 + 32 * wm + 8 * m + lr;
        double rs = 1.0;
        if (ep.hrow && row < ep.nvalid) rs = st_hs(ep.hrow[row], spin, ep.row_inv, hc);
        if (ep.rvec) rs *= st_vscale(ep.rvec[row], ep.rmode);
#pragma unroll
        for (int n = 0; n < NFRAG; ++n) {
          const int col0 = j0 + 8 * NFRAG * wn + 8 * n + 2 * lk;
          double v0 = acc[m][n][0] * rs, v1 = acc[m][n][1] * rs;
          if (ep.cvec) { v0 *= st_vscale(ep.cvec[col0], ep.cmode); v1 *= st_vscale(ep.cvec[col0 + 1], ep.cmode); }
          if (!ep.transposed_out) {
            if (ep.addend) { v0 += ep.addend[(size_t)row * NP + col0]; v1 += ep.addend[(size_t)row * NP + col0 + 1]; }
            *reinterpret_cast<double2*>(Cout + (size_t)row * NP + col0) = make_double2(v0, v1);
          } else {
            if (ep.addend) { v0 += ep.addend[(size_t)col0 * NP + row]; v1 += ep.addend[(size_t)(col0 + 1) * NP + row]; }
            Cout[(size_t)col0 * NP + row] = v0;
            Cout[(size_t)(col0 + 1) * NP + row] = v1;
          }
        }
      }
    }
  }
  __syncthreads();
}

// out[j][i] = f(in[i][j], i, j) over the full padded matrix, through 32 x 33 shared tiles (both sides coalesced)
template <class F>
__device__ void st_transpose_map(const double* __restrict__ in, double* __restrict__ out, int NP, double* tile, F f) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i0 = 0; i0 < NP; i0 += 32)
    for (int j0 = 0; j0 < NP; j0 += 32) {
#pragma unroll
      for (int q = 0; q < 4; ++q) tile[(warp + 8 * q) * 33 + lane] = in[(size_t)(i0 + warp + 8 * q) * NP + j0 + lane];
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = j0 + warp + 8 * q, i = i0 + lane;
        out[(size_t)j * NP + i] = f(tile[lane * 33 + warp + 8 * q], i, j);
      }
      __syncthreads();
    }
}

// reciprocal to ~1 ulp: hardware seed + one Newton step (no denormal / overflow fix-ups: operands here are O(norms))
__device__ __forceinline__ double st_rcp(double d) {
  double r = __drcp_rn(d);
  return r;
}

__device__ __forceinline__ double st_warp_sum(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// ---- st_chain_kernel ---------------------------------------------------------------------------------------------
// normal     : X <- B_{s_last} ... B_{s_first} * src * diag(dcol),   B_s = E diag(v_s)      (left products)
// transposed : X <- B_{s_last}^T ... B_{s_first}^T * src * diag(dcol),  B_s^T = diag(v_s) E^T  (right products, transposed)
// slices s_i = (s_start + i s_step) mod L.  The result is in buf[count & 1].
struct StChainArgs {
  const double* src; size_t src_stride;      // nullptr: identity
  const double* dcol; size_t dcol_stride;    // nullptr: no column scale
  double* buf0; double* buf1;                // [2C][NPs^2]
  const int8_t* field;                       // [chain][L][NPf]
  const double* Eop;                         // k-major left operand: E^T stored row-major (normal) / E row-major (transposed)
  int N, NPs, NPf, L;
  int s_start, s_step, count, transposed;
  HsConsts hc;
};

template <int NFRAG>
__global__ void __launch_bounds__(ST_THREADS) st_chain_kernel(const StChainArgs a) {
  extern __shared__ __align__(16) unsigned char st_smem[];
  double* pa = reinterpret_cast<double*>(st_smem);
  double* pb = pa + ST_STAGES * ST_BK * ST_LDA;
  const int m = blockIdx.x, chain = m >> 1, spin = m & 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NP = a.NPs;
  const int8_t* field = a.field + (size_t)chain * a.L * a.NPf;
  double* buf[2] = {a.buf0 + (size_t)m * NP * NP, a.buf1 + (size_t)m * NP * NP};
  const double* src = a.src ? a.src + (size_t)m * a.src_stride : nullptr;
  const double* dcol = a.dcol ? a.dcol + (size_t)m * a.dcol_stride : nullptr;
  auto slice_of = [&](int i) { int s = (a.s_start + i * a.s_step) % a.L; return s < 0 ? s + a.L : s; };
  {
    const int8_t* h0 = field + (size_t)slice_of(0) * a.NPf;
    for (int r = warp; r < NP; r += ST_THREADS / 32) {
      const double rs = (!a.transposed && r < a.N) ? st_hs(h0[r], spin, false, a.hc) : 1.0;
      for (int c = lane; c < NP; c += 32) {
        const double x = src ? src[(size_t)r * NP + c] : (r == c ? 1.0 : 0.0);
        buf[0][(size_t)r * NP + c] = x * rs;
      }
    }
  }
  for (int i = 0; i < a.count; ++i) {
    StEpilogue ep;
    ep.nvalid = a.N;
    const bool last = (i == a.count - 1);
    if (!a.transposed) {
      if (!last) ep.hrow = field + (size_t)slice_of(i + 1) * a.NPf;
    } else {
      ep.hrow = field + (size_t)slice_of(i) * a.NPf;
    }
    if (last) ep.cvec = dcol;
    st_gemm<NFRAG>(a.Eop, buf[i & 1], nullptr, buf[(i + 1) & 1], NP, spin, ep, a.hc, pa, pb);
  }
}

// ---- blocked Householder QR ----------------------------------------------------------------------------------------
// Compact-WY block reflector applied from the left to A[row0 : row0+m, c0 : c1):
//     A <- (I - V Top V^T) A,     Top = T^T (TRANS_T, the factorization's Q^T A) or T (forming Q)
// V (m x NB, unit lower trapezoid, explicit) and T (NB x NB upper triangular) are in shared memory.  Thread pairs
// (tid, tid + 128) share two columns {c, c + 128} of the block and split the reflector index range.
template <int NB, bool TRANS_T>
__device__ void st_apply_reflector_dfma(double* __restrict__ Amat, int ld, int row0, int m, int c0, int c1,
                                        const double* __restrict__ Vs, const double* __restrict__ Ts, double* __restrict__ Wt) {
  constexpr int LDV = NB + 4, NA = NB / 2;
  const int tid = threadIdx.x, cl = tid & 127, half = tid >> 7;
  const int a0 = half * NA;
  double* const abase = Amat + (size_t)row0 * ld;
  for (int cb = c0; cb < c1; cb += ST_CB) {
    const int ca = cb + cl, cb2 = ca + 128;
    const bool va = ca < c1, vb = cb2 < c1;
    // W = V^T A_blk
    double acc0[NA], acc1[NA];
#pragma unroll
    for (int q = 0; q < NA; ++q) acc0[q] = acc1[q] = 0.0;
#pragma unroll 4
    for (int r = 0; r < m; ++r) {
      const double x0 = va ? abase[(size_t)r * ld + ca] : 0.0;
      const double x1 = vb ? abase[(size_t)r * ld + cb2] : 0.0;
      const double* vr = Vs + r * LDV + a0;
#pragma unroll
      for (int q = 0; q < NA; q += 2) {
        const double2 v = *reinterpret_cast<const double2*>(vr + q);
        acc0[q] = fma(v.x, x0, acc0[q]); acc0[q + 1] = fma(v.y, x0, acc0[q + 1]);
        acc1[q] = fma(v.x, x1, acc1[q]); acc1[q + 1] = fma(v.y, x1, acc1[q + 1]);
      }
    }
#pragma unroll
    for (int q = 0; q < NA; ++q) { Wt[(a0 + q) * ST_CB + cl] = acc0[q]; Wt[(a0 + q) * ST_CB + 128 + cl] = acc1[q]; }
    __syncthreads();
    // W2 = Top W  (this thread's reflector range, its two columns)
#pragma unroll
    for (int q = 0; q < NA; ++q) acc0[q] = acc1[q] = 0.0;
    for (int b = 0; b < NB; ++b) {
      const double y0 = Wt[b * ST_CB + cl], y1 = Wt[b * ST_CB + 128 + cl];
#pragma unroll
      for (int q = 0; q < NA; ++q) {
        const double t = TRANS_T ? Ts[b * NB + a0 + q] : Ts[(a0 + q) * NB + b];
        acc0[q] = fma(t, y0, acc0[q]);
        acc1[q] = fma(t, y1, acc1[q]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NA; ++q) { Wt[(a0 + q) * ST_CB + cl] = acc0[q]; Wt[(a0 + q) * ST_CB + 128 + cl] = acc1[q]; }
    __syncthreads();
    // A_blk <- A_blk - V W2 : the pair splits the rows
    double f0[NB], f1[NB];
#pragma unroll
    for (int q = 0; q < NB; ++q) { f0[q] = Wt[q * ST_CB + cl]; f1[q] = Wt[q * ST_CB + 128 + cl]; }
#pragma unroll 2
    for (int r = half; r < m; r += 2) {
      double x0 = va ? abase[(size_t)r * ld + ca] : 0.0;
      double x1 = vb ? abase[(size_t)r * ld + cb2] : 0.0;
      const double* vr = Vs + r * LDV;
#pragma unroll
      for (int q = 0; q < NB; q += 2) {
        const double2 v = *reinterpret_cast<const double2*>(vr + q);
        x0 = fma(-v.x, f0[q], x0); x0 = fma(-v.y, f0[q + 1], x0);
        x1 = fma(-v.x, f1[q], x1); x1 = fma(-v.y, f1[q + 1], x1);
      }
      if (va) abase[(size_t)r * ld + ca] = x0;
      if (vb) abase[(size_t)r * ld + cb2] = x1;
    }
    __syncthreads();
  }
}

// The same block reflector on DMMA fragments (mma.sync.m8n8k4.f64).  The DFMA version above is LSU-bound: every V value it
// loads (a broadcast shared-memory read) feeds two FMAs (ncu, profiles/r01d_stab_summary.md: 43 % LSU wavefronts, FP64 pipe
// 22 %); a DMMA fragment is one double per 8 FMAs.  Work is split by COLUMN STRIPS, one warp per strip of 8 NT columns, and all
// three steps of a strip are warp-local - no block barrier inside:
//   W  = V^T A_strip     2 x NT accumulator tiles, k = rows: A fragments straight from global memory (a B fragment is 4 rows x
//                        8 consecutive doubles: whole 32-byte sectors), V^T fragments from shared memory;
//   W2 = -Top W          through the warp's private [16][ST_LDS] shared strip (accumulator layout -> B-fragment layout);
//   A_strip += V W2      W2 fragments stay in registers for the whole strip, the accumulator tiles ARE the 16-row slabs of A
//                        (128-bit loads / stores), the next slab's loads in flight behind the DMMAs of the current one.
// Rows are processed in slabs of 16 up to mp = roundup(m, 16): the caller keeps V rows >= m zero, and the matrix padding
// (rows / columns >= N) holds zeros (A) or the identity (Q), so no element masks are needed - only warp-uniform tile guards.
template <int NT, bool TRANS_T>
__device__ __forceinline__ void st_apply_strip(double* __restrict__ abase, int ld, int mp, int cw, int c1, const double* __restrict__ Vs,
                                               const double* __restrict__ Ts, double* __restrict__ Ws, int lr, int lk) {
  constexpr int NB = ST_NB, LDV = NB + 4;
  bool nv[NT];
#pragma unroll
  for (int n = 0; n < NT; ++n) nv[n] = cw + 8 * n < c1;
  // ---- W = V^T A_strip
  double w[2][NT][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int n = 0; n < NT; ++n) w[mt][n][0] = w[mt][n][1] = 0.0;
  for (int r = 0; r < mp; r += 16) {
    double b[4][NT];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int n = 0; n < NT; ++n) b[ks][n] = nv[n] ? abase[(size_t)(r + 4 * ks + lk) * ld + cw + 8 * n + lr] : 0.0;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const double* vp = Vs + (r + 4 * ks + lk) * LDV + lr;
      const double a0 = vp[0], a1 = vp[8];
#pragma unroll
      for (int n = 0; n < NT; ++n) { dmma884(w[0][n], a0, b[ks][n]); dmma884(w[1][n], a1, b[ks][n]); }
    }
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int n = 0; n < NT; ++n)
      *reinterpret_cast<double2*>(Ws + (8 * mt + lr) * ST_LDS + 8 * n + 2 * lk) = make_double2(w[mt][n][0], w[mt][n][1]);
  __syncwarp();
  // ---- W2 = -Top W   (Top = T^T for the factorization's Q^T A, T when forming Q)
  double bw[4][NT];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int n = 0; n < NT; ++n) bw[ks][n] = Ws[(4 * ks + lk) * ST_LDS + 8 * n + lr];
  __syncwarp();
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    double acc[NT][2];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const double t = TRANS_T ? Ts[(4 * ks + lk) * NB + 8 * mt + lr] : Ts[(8 * mt + lr) * NB + 4 * ks + lk];
#pragma unroll
      for (int n = 0; n < NT; ++n) dmma884(acc[n], -t, bw[ks][n]);
    }
#pragma unroll
    for (int n = 0; n < NT; ++n)
      *reinterpret_cast<double2*>(Ws + (8 * mt + lr) * ST_LDS + 8 * n + 2 * lk) = make_double2(acc[n][0], acc[n][1]);
  }
  __syncwarp();
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int n = 0; n < NT; ++n) bw[ks][n] = Ws[(4 * ks + lk) * ST_LDS + 8 * n + lr];
  __syncwarp();
  // ---- A_strip += V W2, 16-row slabs
  double2 nxt[2][NT];
  auto load_slab = [&](int r, double2 (&dst)[2][NT]) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int n = 0; n < NT; ++n)
        dst[mt][n] = nv[n] ? *reinterpret_cast<const double2*>(abase + (size_t)(r + 8 * mt + lr) * ld + cw + 8 * n + 2 * lk) : make_double2(0.0, 0.0);
  };
  if (mp > 0) load_slab(0, nxt);
  for (int r = 0; r < mp; r += 16) {
    double c[2][NT][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int n = 0; n < NT; ++n) { c[mt][n][0] = nxt[mt][n].x; c[mt][n][1] = nxt[mt][n].y; }
    if (r + 16 < mp) load_slab(r + 16, nxt);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const double* vp = Vs + (r + lr) * LDV + 4 * ks + lk;
      const double a0 = vp[0], a1 = vp[8 * LDV];
#pragma unroll
      for (int n = 0; n < NT; ++n) { dmma884(c[0][n], a0, bw[ks][n]); dmma884(c[1][n], a1, bw[ks][n]); }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int n = 0; n < NT; ++n)
        if (nv[n]) *reinterpret_cast<double2*>(abase + (size_t)(r + 8 * mt + lr) * ld + cw + 8 * n + 2 * lk) = make_double2(c[mt][n][0], c[mt][n][1]);
  }
}

template <int NB, bool TRANS_T>
__device__ void st_apply_reflector(double* __restrict__ Amat, int ld, int row0, int m, int c0, int c1,
                                   const double* __restrict__ Vs, const double* __restrict__ Ts, double* __restrict__ Wt) {
#if LQMC_ST_APPLY_DMMA
  static_assert(NB == 16, "two 8-row accumulator tiles per reflector block");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lr = lane >> 2, lk = lane & 3;
  const int mp = (m + 15) & ~15;
  double* const abase = Amat + (size_t)row0 * ld;
  double* const Ws = Wt + warp * NB * ST_LDS;
  constexpr int NW = ST_THREADS / 32;
  if (c1 - c0 > 16 * NW) {
    for (int cw = c0 + 32 * warp; cw < c1; cw += 32 * NW) st_apply_strip<4, TRANS_T>(abase, ld, mp, cw, c1, Vs, Ts, Ws, lr, lk);
  } else {
    for (int cw = c0 + 16 * warp; cw < c1; cw += 16 * NW) st_apply_strip<2, TRANS_T>(abase, ld, mp, cw, c1, Vs, Ts, Ws, lr, lk);
  }
  __syncthreads();
#else
  st_apply_reflector_dfma<NB, TRANS_T>(Amat, ld, row0, m, c0, c1, Vs, Ts, Wt);
#endif
}

template <int NB>
struct StQrSmem {
  static constexpr int LDV = NB + 4;     // 2 LDV = 8 (mod 32): the DMMA fragment loads of V are bank-conflict-free
  double* P;      // [rows][LDV] panel / reflectors
  double* Ts;     // [NB][NB]
  double* Gm;     // [NB][NB]
  double* taus;   // [NB]
  double* dots;   // [NB]
  double* Wt;     // [NB][ST_WT_LD]: DFMA variant [NB][ST_CB]; DMMA variant 8 warps x [NB][ST_LDS]
  __device__ StQrSmem(unsigned char* base, int rows) {
    P = reinterpret_cast<double*>(base);
    Ts = P + (size_t)rows * LDV;
    Gm = Ts + NB * NB;
    taus = Gm + NB * NB;
    dots = taus + NB;
    Wt = dots + NB;
  }
  static size_t bytes(int rows) { return ((size_t)rows * (NB + 4) + 2 * NB * NB + 2 * NB + (size_t)NB * ST_WT_LD) * sizeof(double); }
};

// T factor of the panel's block reflector from the explicit V in sm.P (m rows, w valid columns) and sm.taus
template <int NB>
__device__ void st_form_t(StQrSmem<NB>& sm, int m, int w) {
  constexpr int LDV = NB + 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int q = tid; q < NB * NB; q += ST_THREADS) { sm.Ts[q] = 0.0; sm.Gm[q] = 0.0; }
  __syncthreads();
  for (int q = warp; q < NB * NB; q += ST_THREADS / 32) {
    const int a = q / NB, b = q % NB;
    if (a < b && b < w) {
      double s = 0.0;
      for (int r = lane; r < m; r += 32) s = fma(sm.P[r * LDV + a], sm.P[r * LDV + b], s);
      s = st_warp_sum(s);
      if (lane == 0) sm.Gm[a * NB + b] = s;
    }
  }
  __syncthreads();
  if (warp == 0) {
    for (int j = 0; j < w; ++j) {
      if (lane < j) {
        double s = 0.0;
        for (int l = lane; l < j; ++l) s = fma(sm.Ts[lane * NB + l], sm.Gm[l * NB + j], s);
        sm.Ts[lane * NB + j] = -sm.taus[j] * s;
      }
      if (lane == j) sm.Ts[j * NB + j] = sm.taus[j];
      __syncwarp();
    }
  }
  __syncthreads();
}

struct StQrArgs {
  const double* M;     // [2C][NPs^2] matrix to factor
  double* A;           // [2C][NPs^2] work: column-permuted copy, then R (upper) + reflectors (lower)
  double* Q; size_t q_stride;
  double* dvec; size_t d_stride;
  double* tfac;        // [2C][NPs * NB]
  int* perm;           // [2C][NPs]
  int N, NPs;
};

template <int NB>
__global__ void __launch_bounds__(ST_THREADS) st_qr_kernel(const StQrArgs a) {
  constexpr int LDV = NB + 4;
  extern __shared__ __align__(16) unsigned char st_smem[];
  const int N = a.N, NP = a.NPs;
  StQrSmem<NB> sm(st_smem, NP);
  const int m_idx = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* M = a.M + (size_t)m_idx * NP * NP;
  double* A = a.A + (size_t)m_idx * NP * NP;
  double* Q = a.Q + (size_t)m_idx * a.q_stride;
  double* dvec = a.dvec + (size_t)m_idx * a.d_stride;
  double* tfac = a.tfac + (size_t)m_idx * NP * NB;
  int* perm = a.perm + (size_t)m_idx * NP;

  // 1. column norms and the descending-norm permutation (ties: lower index first)
  double* norms = sm.P;
  for (int c = tid; c < N; c += ST_THREADS) {
    double s = 0.0;
    for (int r = 0; r < N; ++r) { const double v = M[(size_t)r * NP + c]; s = fma(v, v, s); }
    norms[c] = s;
  }
  __syncthreads();
  for (int c = tid; c < NP; c += ST_THREADS) {
    if (c < N) {
      const double nc = norms[c];
      int rank = 0;
      for (int o = 0; o < N; ++o) { const double no = norms[o]; rank += (no > nc || (no == nc && o < c)) ? 1 : 0; }
      perm[rank] = c;
    } else {
      perm[c] = c;
    }
  }
  __syncthreads();
  // 2. A = M[:, perm]
  // (rows up to roundup(N, 16) and all NP columns: the DMMA reflector application runs over whole tiles of zero padding)
  const int n16 = (N + 15) & ~15;
  for (int r = warp; r < n16; r += ST_THREADS / 32)
    for (int j = lane; j < NP; j += 32) A[(size_t)r * NP + j] = (r < N && j < N) ? M[(size_t)r * NP + perm[j]] : 0.0;
  __syncthreads();

  // 3. panels
  for (int j0 = 0; j0 < N; j0 += NB) {
    const int w = min(NB, N - j0), m = N - j0;
    for (int r = warp; r < ((m + 15) & ~15); r += ST_THREADS / 32)
      if (lane < NB) sm.P[r * LDV + lane] = (lane < w && r < m) ? A[(size_t)(j0 + r) * NP + j0 + lane] : 0.0;
    if (tid < NB) sm.taus[tid] = 0.0;
    __syncthreads();
    double scale_prev = 0.0, beta_prev = 0.0, tau_prev = 0.0;
    for (int j = 0; j < w; ++j) {
      // finish column j-1 (nobody reads it any more): scale the reflector, store beta and tau
      if (j > 0) {
        for (int r = j + tid; r < m; r += ST_THREADS) sm.P[r * LDV + j - 1] *= scale_prev;
        if (tid == 0) { sm.P[(j - 1) * LDV + j - 1] = beta_prev; sm.taus[j - 1] = tau_prev; }
      }
      // dots[c] = sum_{r > j} P[r][j] P[r][c]  for c = j .. w-1
      for (int c = j + warp; c < w; c += ST_THREADS / 32) {
        double s = 0.0;
        for (int r = j + 1 + lane; r < m; r += 32) s = fma(sm.P[r * LDV + j], sm.P[r * LDV + c], s);
        s = st_warp_sum(s);
        if (lane == 0) sm.dots[c] = s;
      }
      __syncthreads();
      const double alpha = sm.P[j * LDV + j], sigma = sm.dots[j];
      double tau = 0.0, beta = alpha, scale = 0.0;
      if (sigma != 0.0) {
        const double nrm = sqrt(fma(alpha, alpha, sigma));
        beta = (alpha >= 0.0) ? -nrm : nrm;
        tau = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
        for (int c = j + 1 + warp; c < w; c += ST_THREADS / 32) {
          const double wv = fma(scale, sm.dots[c], sm.P[j * LDV + c]);
          const double tw = tau * wv, tws = tw * scale;
          for (int r = j + 1 + lane; r < m; r += 32) sm.P[r * LDV + c] = fma(-tws, sm.P[r * LDV + j], sm.P[r * LDV + c]);
          __syncwarp();
          if (lane == 0) sm.P[j * LDV + c] -= tw;
        }
      }
      scale_prev = scale; beta_prev = beta; tau_prev = tau;
      __syncthreads();
    }
    {
      const int j = w;
      for (int r = j + tid; r < m; r += ST_THREADS) sm.P[r * LDV + j - 1] *= scale_prev;
      if (tid == 0) { sm.P[(j - 1) * LDV + j - 1] = beta_prev; sm.taus[j - 1] = tau_prev; }
    }
    __syncthreads();
    // write the factored panel back (R on and above the diagonal, reflectors below) and make V explicit
    for (int r = warp; r < m; r += ST_THREADS / 32)
      if (lane < w) A[(size_t)(j0 + r) * NP + j0 + lane] = sm.P[r * LDV + lane];
    __syncthreads();
    for (int q = tid; q < NB * NB; q += ST_THREADS) {
      const int r = q / NB, c = q % NB;
      if (r < m && c >= r) sm.P[r * LDV + c] = (c == r && c < w) ? 1.0 : 0.0;
    }
    __syncthreads();
    st_form_t<NB>(sm, m, w);
    for (int q = tid; q < NB * NB; q += ST_THREADS) tfac[(size_t)j0 * NB + q] = sm.Ts[q];
    if (j0 + w < N) st_apply_reflector<NB, true>(A, NP, j0, m, j0 + w, N, sm.P, sm.Ts, sm.Wt);
    __syncthreads();
  }

  // 4. D = |diag R|
  for (int j = tid; j < NP; j += ST_THREADS) {
    double d = 1.0;
    if (j < N) { d = fabs(A[(size_t)j * NP + j]); if (!(d > 0.0)) d = 1e-300; }
    dvec[j] = d;
  }
  // 5. explicit Q = H_1 ... H_N
  for (int r = warp; r < NP; r += ST_THREADS / 32)
    for (int c = lane; c < NP; c += 32) Q[(size_t)r * NP + c] = (r == c) ? 1.0 : 0.0;
  __syncthreads();
  const int n_panels = (N + NB - 1) / NB;
  for (int pb = n_panels - 1; pb >= 0; --pb) {
    const int j0 = pb * NB, w = min(NB, N - j0), m = N - j0;
    for (int r = warp; r < ((m + 15) & ~15); r += ST_THREADS / 32)
      if (lane < NB) {
        double v = 0.0;
        if (lane < w && r < m) v = (r > lane) ? A[(size_t)(j0 + r) * NP + j0 + lane] : (r == lane ? 1.0 : 0.0);
        sm.P[r * LDV + lane] = v;
      }
    for (int q = tid; q < NB * NB; q += ST_THREADS) sm.Ts[q] = tfac[(size_t)j0 * NB + q];
    __syncthreads();
    st_apply_reflector<NB, false>(Q, NP, j0, m, j0, N, sm.P, sm.Ts, sm.Wt);
  }
}

// ---- NPs = 64: the whole factorization in shared memory ------------------------------------------------------------
// For N <= 64 the blocked kernel above is pure latency (900 K clocks per 64 x 64 matrix: global round trips per row of
// the reflector application).  Here the matrix (33 KB) and Q live in shared memory; unblocked Householder, 256 threads =
// 64 columns x 4 row quarters, two barriers per column; same outputs (A = R + reflectors is not needed afterwards, only
// R, D, Q and the permutation).
__global__ void __launch_bounds__(ST_THREADS) st_qr_small_kernel(const StQrArgs a) {
  constexpr int NP = 64, LD = 65;
  extern __shared__ __align__(16) unsigned char st_smem[];
  double* S0 = reinterpret_cast<double*>(st_smem);      // M, later Q
  double* S1 = S0 + NP * LD;                            // A = M[:, perm], factored in place
  double* red = S1 + NP * LD;                           // [4][64] partial dot products
  double* norms = red + 4 * NP;                         // [64]
  double* taus = norms + NP;                            // [64]
  int* perm_s = reinterpret_cast<int*>(taus + NP);      // [64]
  const int N = a.N;
  const int m_idx = blockIdx.x, tid = threadIdx.x;
  const int c = tid & 63, q = tid >> 6, r_lo = 16 * q, r_hi = 16 * q + 16;
  const double* M = a.M + (size_t)m_idx * NP * NP;
  double* A = a.A + (size_t)m_idx * NP * NP;
  double* Q = a.Q + (size_t)m_idx * a.q_stride;
  double* dvec = a.dvec + (size_t)m_idx * a.d_stride;
  int* perm = a.perm + (size_t)m_idx * NP;
  for (int idx = tid; idx < NP * NP; idx += ST_THREADS) S0[(idx >> 6) * LD + (idx & 63)] = M[idx];
  __syncthreads();
  // column norms (4 partial sums per column) -> descending permutation
  {
    double s = 0.0;
    if (c < N)
      for (int r = r_lo; r < r_hi; ++r) if (r < N) { const double v = S0[r * LD + c]; s = fma(v, v, s); }
    red[q * NP + c] = s;
  }
  __syncthreads();
  if (tid < NP) norms[tid] = red[tid] + red[NP + tid] + red[2 * NP + tid] + red[3 * NP + tid];
  __syncthreads();
  if (tid < NP) {
    if (tid < N) {
      const double nc = norms[tid];
      int rank = 0;
      for (int o = 0; o < N; ++o) { const double no = norms[o]; rank += (no > nc || (no == nc && o < tid)) ? 1 : 0; }
      perm_s[rank] = tid;
    } else {
      perm_s[tid] = tid;
    }
  }
  __syncthreads();
  if (tid < NP) perm[tid] = perm_s[tid];
  for (int idx = tid; idx < NP * NP; idx += ST_THREADS) {
    const int r = idx >> 6, j = idx & 63;
    S1[r * LD + j] = (r < N && j < N) ? S0[r * LD + perm_s[j]] : (r == j ? 1.0 : 0.0);
  }
  __syncthreads();
  // Householder QR of the leading N x N block of S1.  The 16-row loops are fully unrolled with predicates so that the
  // shared loads of a phase are issued together (a rolled loop is one load-to-use latency per row: 600 K clocks per matrix).
  const double* col_c = S1 + r_lo * LD + c;
  double scale_prev = 0.0, beta_prev = 0.0, tau_prev = 0.0;
  for (int j = 0; j < N; ++j) {
    if (j > 0 && c == j - 1) {                              // finish column j-1: nobody reads it any more
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) if (r_lo + rr >= j && r_lo + rr < N) S1[(r_lo + rr) * LD + j - 1] *= scale_prev;
      if (q == 0) { S1[(j - 1) * LD + j - 1] = beta_prev; taus[j - 1] = tau_prev; }
    }
    const double ajc = S1[j * LD + c];                       // row j is read here, written (by its owner) after the barrier
    const bool mine = (c >= j && c < N);
    // rows of this quarter below the diagonal and inside the matrix: local indices [lo, hi)
    const int lo = max(0, j + 1 - r_lo), hi = min(16, N - r_lo);
    double vj[16], ac[16];
    if (lo < hi && mine) {                                    // warp-uniform up to the `mine` edge; skipped quarters cost nothing
      const double* col_j = S1 + r_lo * LD + j;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) {
        const bool on = (rr >= lo) && (rr < hi);
        vj[rr] = on ? col_j[rr * LD] : 0.0;
        ac[rr] = on ? col_c[rr * LD] : 0.0;
      }
#pragma unroll
      for (int rr = 0; rr < 16; rr += 2) { s0 = fma(vj[rr], ac[rr], s0); s1 = fma(vj[rr + 1], ac[rr + 1], s1); }
      red[q * NP + c] = s0 + s1;
    } else {
      red[q * NP + c] = 0.0;
    }
    __syncthreads();
    const double sigma = red[j] + red[NP + j] + red[2 * NP + j] + red[3 * NP + j];
    const double alpha = S1[j * LD + j];
    double tau = 0.0, beta = alpha, scale = 0.0;
    if (sigma != 0.0) {
      // every thread evaluates these scalars: keep them cheap (IEEE sqrt and divide are ~40-instruction software
      // sequences on the FP64 pipe and were 70 % of this kernel's FP64 issue slots); 1-ulp results are ample here
      const double x2 = fma(alpha, alpha, sigma);
      const double rs = rsqrt(x2);
      double nrm = x2 * rs;
      nrm = fma(fma(-nrm, nrm, x2), 0.5 * rs, nrm);
      beta = (alpha >= 0.0) ? -nrm : nrm;
      tau = (beta - alpha) * st_rcp(beta);
      scale = st_rcp(alpha - beta);
      if (c > j && c < N) {
        const double dot = red[c] + red[NP + c] + red[2 * NP + c] + red[3 * NP + c];
        const double wv = fma(scale, dot, ajc);
        const double tw = tau * wv, tws = tw * scale;
        if (lo < hi) {
#pragma unroll
          for (int rr = 0; rr < 16; ++rr)
            if (rr >= lo && rr < hi) S1[(r_lo + rr) * LD + c] = fma(-tws, vj[rr], ac[rr]);
        }
        if (j >= r_lo && j < r_hi) S1[j * LD + c] = ajc - tw;
      }
    }
    scale_prev = scale; beta_prev = beta; tau_prev = tau;
    __syncthreads();
  }
  if (c == N - 1 && q == 0) { S1[(N - 1) * LD + N - 1] = beta_prev; taus[N - 1] = tau_prev; }
  __syncthreads();
  // D, R (upper triangle) back to global; Q = H_0 ... H_{N-1} in S0
  if (tid < NP) {
    double d = 1.0;
    if (tid < N) { d = fabs(S1[tid * LD + tid]); if (!(d > 0.0)) d = 1e-300; }
    dvec[tid] = d;
  }
  for (int idx = tid; idx < NP * NP; idx += ST_THREADS) {
    const int r = idx >> 6, jj = idx & 63;
    A[idx] = S1[r * LD + jj];
    S0[r * LD + jj] = (r == jj) ? 1.0 : 0.0;
  }
  __syncthreads();
  double* qcol_c = S0 + r_lo * LD + c;
  for (int j = N - 1; j >= 0; --j) {
    const double tau = taus[j];
    const bool mine = (c >= j && c < N);
    const int lo = max(0, j - r_lo), hi = min(16, N - r_lo);      // rows j .. N-1 of this quarter
    const bool work = (lo < hi) && mine && (tau != 0.0);
    double vj[16], qc[16];
    if (work) {
      const double* col_j = S1 + r_lo * LD + j;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) {
        const bool on = (rr >= lo) && (rr < hi);
        vj[rr] = on ? ((r_lo + rr == j) ? 1.0 : col_j[rr * LD]) : 0.0;
        qc[rr] = on ? qcol_c[rr * LD] : 0.0;
      }
#pragma unroll
      for (int rr = 0; rr < 16; rr += 2) { s0 = fma(vj[rr], qc[rr], s0); s1 = fma(vj[rr + 1], qc[rr + 1], s1); }
      red[q * NP + c] = s0 + s1;
    } else {
      red[q * NP + c] = 0.0;
    }
    __syncthreads();
    if (work) {
      const double tw = tau * (red[c] + red[NP + c] + red[2 * NP + c] + red[3 * NP + c]);
#pragma unroll
      for (int rr = 0; rr < 16; ++rr)
        if (rr >= lo && rr < hi) qcol_c[rr * LD] = fma(-tw, vj[rr], qc[rr]);
    }
    __syncthreads();
  }
  for (int idx = tid; idx < NP * NP; idx += ST_THREADS) Q[idx] = S0[(idx >> 6) * LD + (idx & 63)];
}

// ---- V <- (D^-1 R) P^T V ---------------------------------------------------------------------------------------------
struct StVArgs {
  const double* R;       // [2C][NPs^2] factored matrix (upper triangle = R)
  const double* dvec; size_t d_stride;
  const int* perm;       // [2C][NPs]
  const double* Vold; size_t vold_stride;     // nullptr: identity
  double* At;            // [2C][NPs^2] scratch for the k-major left operand
  double* Vnew; size_t vnew_stride;
  int N, NPs;
  HsConsts hc;
};

template <int NFRAG>
__global__ void __launch_bounds__(ST_THREADS) st_v_kernel(const StVArgs a) {
  extern __shared__ __align__(16) unsigned char st_smem[];
  double* pa = reinterpret_cast<double*>(st_smem);
  double* pb = pa + ST_STAGES * ST_BK * ST_LDA;
  const int m = blockIdx.x, NP = a.NPs, N = a.N;
  const double* R = a.R + (size_t)m * NP * NP;
  const double* dvec = a.dvec + (size_t)m * a.d_stride;
  const int* perm = a.perm + (size_t)m * NP;
  double* At = a.At + (size_t)m * NP * NP;
  double* Vnew = a.Vnew + (size_t)m * a.vnew_stride;
  // At[k][i] = R[i][k] / D[i]  (i <= k < N), identity elsewhere
  st_transpose_map(R, At, NP, pa, [&](double x, int i, int k) {
    if (k < N && i <= k) return x / dvec[i];
    return (i == k) ? 1.0 : 0.0;
  });
  if (a.Vold) {
    StEpilogue ep;
    st_gemm<NFRAG>(At, a.Vold + (size_t)m * a.vold_stride, perm, Vnew, NP, m & 1, ep, a.hc, pa, pb);
  } else {
    // V_old = I: V_new[i][perm[k]] = At[k][i]
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = warp; k < NP; k += ST_THREADS / 32) {
      const int pk = perm[k];
      for (int i = lane; i < NP; i += 32) Vnew[(size_t)i * NP + pk] = At[(size_t)k * NP + i];
    }
  }
}

// ---- one-sided G = (D_b^-1 U^T + D_s V)^-1 D_b^-1 U^T ---------------------------------------------------------------
// st_forms:   lhsT[i][j] = U[i][j] / db[j] + V[j][i] ds[j]   (the transposed left-hand side, row-major)
//             rhs[k][i]  = U[i][k] / db[k]                   (= D_b^-1 U^T row-major = k-major left operand of the last GEMM)
// then  W = inv(lhsT)  and  G^T = (U D_b^-1) W.
struct StFormsArgs {
  const double* U; size_t u_stride;
  const double* V; size_t v_stride;
  const double* dvec; size_t d_stride;
  double* lhsT; double* rhs;      // [2C][NPs^2]
  int NPs;
};

__global__ void __launch_bounds__(ST_THREADS) st_forms_kernel(const StFormsArgs a) {
  __shared__ double tile[32 * 33];
  const int m = blockIdx.x, NP = a.NPs;
  const double* U = a.U + (size_t)m * a.u_stride;
  const double* V = a.V + (size_t)m * a.v_stride;
  const double* d = a.dvec + (size_t)m * a.d_stride;
  double* lhsT = a.lhsT + (size_t)m * NP * NP;
  double* rhs = a.rhs + (size_t)m * NP * NP;
  // rhs[k][i] = U[i][k] / max(d[k], 1)
  st_transpose_map(U, rhs, NP, tile, [&](double x, int i, int k) { return x / fmax(d[k], 1.0); });
  // lhsT[i][j] = V[j][i] * min(d[j], 1)   (+ U[i][j] / max(d[j], 1) below)
  st_transpose_map(V, lhsT, NP, tile, [&](double x, int j, int i) { return x * fmin(d[j], 1.0); });
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = warp; i < NP; i += ST_THREADS / 32)
    for (int j = lane; j < NP; j += 32) lhsT[(size_t)i * NP + j] += U[(size_t)i * NP + j] / fmax(d[j], 1.0);
}

// in-place inverse of one matrix per CTA, large sizes: delayed Gauss-Jordan of sweep_l2.cuh on a single matrix
struct StInvArgs { double* M; int* piv; int NPs, KD; };

__global__ void __launch_bounds__(L2_THREADS, 2) st_inverse_l2_kernel(const StInvArgs a) {
  extern __shared__ __align__(16) unsigned char st_smem[];
  L2Smem sm(st_smem, a.NPs, a.KD, 1, false);
  l2_gj_inverse<1>(a.M + (size_t)blockIdx.x * a.NPs * a.NPs, a.NPs, a.KD, sm, a.piv + (size_t)blockIdx.x * a.NPs);
}

// small sizes (NPs = 64): shared-memory Gauss-Jordan of sweep_reg.cuh, 128 threads
__global__ void __launch_bounds__(128) st_inverse_small_kernel(const StInvArgs a) {
  using C = RegCfg<64, 16, 8>;
  extern __shared__ __align__(16) unsigned char st_smem[];
  RegSmem<C> sm(st_smem);
  double* M = a.M + (size_t)blockIdx.x * 64 * 64;
  const int t = threadIdx.x;
  for (int q = t; q < 64 * 64; q += 128) sm.stage[(q >> 6) * C::S + (q & 63)] = M[q];
  __syncthreads();
  gj_inverse<C>(sm.stage, sm, 0, t);
  for (int q = t; q < 64 * 64; q += 128) M[q] = sm.stage[(q >> 6) * C::S + (q & 63)];
}

// ---- generic batched GEMM and transpose kernels for the combination formulas -----------------------------------------
struct StGemmArgs {
  const double* At; size_t at_stride;       // k-major left operand
  const double* B; size_t b_stride;
  double* out; size_t out_stride;
  const double* rvec; size_t rvec_stride; int rmode;
  const double* cvec; size_t cvec_stride; int cmode;
  const double* addend; size_t addend_stride;
  double* G;             // optional: copy the leading N x N block of the result into [2C][NPg][NPg]
  int N, NPs, NPg, transposed_out;
  HsConsts hc;
};

template <int NFRAG>
__global__ void __launch_bounds__(ST_THREADS) st_gemm_kernel(const StGemmArgs a) {
  extern __shared__ __align__(16) unsigned char st_smem[];
  double* pa = reinterpret_cast<double*>(st_smem);
  double* pb = pa + ST_STAGES * ST_BK * ST_LDA;
  const int m = blockIdx.x, NP = a.NPs;
  double* out = a.out + (size_t)m * a.out_stride;
  StEpilogue ep;
  ep.transposed_out = a.transposed_out != 0;
  if (a.rvec) { ep.rvec = a.rvec + (size_t)m * a.rvec_stride; ep.rmode = a.rmode; }
  if (a.cvec) { ep.cvec = a.cvec + (size_t)m * a.cvec_stride; ep.cmode = a.cmode; }
  if (a.addend) ep.addend = a.addend + (size_t)m * a.addend_stride;
  st_gemm<NFRAG>(a.At + (size_t)m * a.at_stride, a.B + (size_t)m * a.b_stride, nullptr, out, NP, m & 1, ep, a.hc, pa, pb);
  if (a.G) {
    double* G = a.G + (size_t)m * a.NPg * a.NPg;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < a.N; r += ST_THREADS / 32)
      for (int c = lane; c < a.N; c += 32) G[(size_t)r * a.NPg + c] = out[(size_t)r * NP + c];
  }
}

// out[k][i] = in[i][k] * scale(vec[k])   (vec optional)
struct StTransposeArgs {
  const double* in; size_t in_stride;
  double* out; size_t out_stride;
  const double* vec; size_t vec_stride; int mode;
  int NPs;
};

__global__ void __launch_bounds__(ST_THREADS) st_transpose_kernel(const StTransposeArgs a) {
  __shared__ double tile[32 * 33];
  const int m = blockIdx.x;
  const double* vec = a.vec ? a.vec + (size_t)m * a.vec_stride : nullptr;
  const int mode = a.mode;
  st_transpose_map(a.in + (size_t)m * a.in_stride, a.out + (size_t)m * a.out_stride, a.NPs, tile,
                   [&](double x, int i, int k) { return vec ? x * st_vscale(vec[k], mode) : x; });
}

// identity matrices / unit vectors for the empty right product
struct StIdentityArgs { double* M0; double* M1; double* d; int NPs; };
__global__ void __launch_bounds__(ST_THREADS) st_identity_kernel(const StIdentityArgs a) {
  const int m = blockIdx.x, NP = a.NPs;
  double* M0 = a.M0 + (size_t)m * NP * NP;
  double* M1 = a.M1 ? a.M1 + (size_t)m * NP * NP : nullptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < NP; r += ST_THREADS / 32)
    for (int c = lane; c < NP; c += 32) {
      const double v = (r == c) ? 1.0 : 0.0;
      M0[(size_t)r * NP + c] = v;
      if (M1) M1[(size_t)r * NP + c] = v;
    }
  if (a.d) for (int j = threadIdx.x; j < NP; j += ST_THREADS) a.d[(size_t)m * NP + j] = 1.0;
}

}  // namespace lqmc
