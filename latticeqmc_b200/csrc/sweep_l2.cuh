// Large-lattice sweep kernel (64 < N <= 1024): one CTA per Markov chain, Green's functions in HBM / L2.
//
// Same reference path as sweep_reg.cuh (LatticeQMC._update_step, /root/reference/lqmc/lqmc.py:301-347), for
// lattices whose two N x N FP64 Green's functions (1 MiB at 16x16, 5 MiB at 24x24) no longer fit one SM's
// registers.  Three things change:
//
//  * Delayed rank-k updates.  An accepted flip does not touch G in memory.  Its Sherman-Morrison vectors
//    (e, c) are appended to shared-memory buffers U, W; the row and column of the *current* G that the next
//    flip needs are rebuilt on the fly as  G0[i,:] - sum_m U_m[i] W_m[:]  (and likewise the column), and the
//    diagonal is carried along in shared memory.  After KD flips (or at the end of the slice) the block
//    G0 <- G0 - U W^T is applied in one pass.  The undelayed update moves 32 N^2 bytes per accepted flip
//    (2 MiB at N = 256); delaying divides that by KD.  In EXACT mode every element sees the same sequence of
//    multiply-then-subtract roundings as the reference's one-flip-at-a-time loop (lqmc.py:328-331), so the
//    delayed path is bit-identical to the undelayed one, not merely close.
//  * The accept/reject scan is lane-parallel: a warp evaluates the ratios of the next 32 sites at once from
//    the shared diagonal; the first accepted site is found with a ballot, everything before it is rejected
//    for free (no flip happened in between, so those ratios were final).
//  * Wrap and sweep-start product are tiled GEMMs on FP64 tensor-core fragments (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4): 64 x 128
//    block tile, 8 warps x (32 x 32) warp tiles, k-major operand panels staged global -> shared through a 3-stage cp.async ring
//    (a cp.async.bulk + mbarrier variant is kept behind LQMC_L2_STAGING_TMA); the left operand is always kept k-major
//    (transposed) in memory so both panels are contiguous row copies.  exp(V_l) is applied as a row / column scale in the
//    epilogue.  Matrices are padded to a multiple of 64 (at least 128); the half tile at the right edge is skipped.
//  * For NP <= 256 the c history of the delayed updates lives in tensor memory (tcgen05.ld / st as a per-thread scratchpad),
//    which doubles the delay depth; 384 < NP <= 768 uses the same idea with two or three columns per thread (one CTA per SM).
//
// FP64 has no tcgen05 kind, and DMMA issues to the same pipe as DFMA on sm_100a (profiles/fp64_peaks_r01.json): the roofline of
// every phase here is the FP64 pipe peak; tensor-core fragments buy operand traffic (one double per 8 FMAs), not flops.
#pragma once
#include <type_traits>
#include <cuda.h>                 // CUtensorMap (types and enums only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <cuda_pipeline_primitives.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sweep_reg.cuh"

namespace lqmc {


constexpr int L2_THREADS = 256;
constexpr int L2_BM = 64, L2_BN = 128, L2_BK = 16;     // block tile and k-panel depth
constexpr int L2_GY = 16, L2_GX = 16;                  // thread grid inside a block tile: 4 x 8 per thread
#ifndef LQMC_L2_STAGING_TMA
#define LQMC_L2_STAGING_TMA 1                          // 1: cp.async.bulk.tensor (tensor maps, 128-byte swizzle) + full / empty mbarrier ring; 0: LDGSTS (cp.async) ring
#endif
#ifndef LQMC_L2_EARLY_RCP
#define LQMC_L2_EARLY_RCP 1
#endif
#ifndef LQMC_L2_MERGE_RC
#define LQMC_L2_MERGE_RC 1
#endif
#ifndef LQMC_L2_ZPAD
#define LQMC_L2_ZPAD 1
#endif
#ifndef LQMC_FLUSH2_EPF
#define LQMC_FLUSH2_EPF 0
#endif
#ifndef LQMC_L2_ROTATE_PRODUCER
#define LQMC_L2_ROTATE_PRODUCER 1
#endif
#ifndef LQMC_L2_COL_OUTER
#define LQMC_L2_COL_OUTER false     // measured neutral to slightly negative (cfg4 271.5 vs 270.9 ms, cfg5 2076 vs 2064 ms): the operand re-reads hit L2 either way
#endif
#ifndef LQMC_FLUSH2_L2PF
#define LQMC_FLUSH2_L2PF 0
#endif
#ifndef LQMC_FLUSHX_L2PF
#define LQMC_FLUSHX_L2PF 1      // chunks ahead of the register prefetch that are pulled into L2 (0: off)
#endif
#ifndef LQMC_FLUSHX_EPF
#define LQMC_FLUSHX_EPF 1
#endif
#ifndef LQMC_L2_WIDE_TILE
#define LQMC_L2_WIDE_TILE 1
#endif
constexpr int L2_STAGES = 3;                           // operand-panel ring depth of the LDGSTS variant
constexpr int L2_TMA_STAGES_MAX = 8;                   // mbarrier slots per direction
constexpr int L2_TMA_STAGES = 4;                       // ... of the TMA variant: 4 x (8 KB + 16 KB) dense, swizzled panels
constexpr int L2_TMA_SUB = 16 * 16 * 8;                // one TMA box: 16 k-rows x 16 doubles (128-byte rows, the swizzle span) = 2 KB
constexpr size_t L2_GEMM_DOUBLES_LDGSTS = (size_t)3 * 16 * (64 + 128 + 8);
constexpr size_t L2_GEMM_DOUBLES_TMA = (size_t)4 * 16 * (64 + 128);
constexpr size_t L2_GEMM_DOUBLES = L2_GEMM_DOUBLES_TMA > L2_GEMM_DOUBLES_LDGSTS ? L2_GEMM_DOUBLES_TMA : L2_GEMM_DOUBLES_LDGSTS;   // operand-stage region (aliases U / W)
constexpr int L2_MAXQ = 4;                             // indices per thread in the vector phases (NP <= 1024)
#ifndef LQMC_L2_KDT
#define LQMC_L2_KDT 24
#endif
constexpr int L2_KDT = LQMC_L2_KDT;                    // delay depth of the tensor-memory slice path (NP <= 256)
constexpr int L2_TMEM_COLS = 256;
constexpr int L2_RING = 8;                             // slots of the "history of the next candidate sites" ring (site & 7)
constexpr int L2_PUB = 4;                              // sites after a flip whose c history is published with the flip itself

// Padded matrix size: a multiple of 64.  The GEMM block tile is 64 x 128 and the flush tiles are 128 columns wide, but both
// (here and in stab.cuh) handle a half tile at the right edge, and padding is expensive: BASELINE configs[4] is N = 576, which a
// multiple of 128 would pad to 640 - (640/576)^3 = 1.37 x the GEMM work (ncu on the cfg5 wrap at NP = 640: DMMA pipe 80 % busy
// at 21.5 algorithmic TFLOP/s).  The TMA staging variant copies whole 128-column rows and keeps the multiple of 128.
inline int l2_padded_size(int n_sites) {
  const int np = (n_sites + 63) / 64 * 64 < 128 ? 128 : (n_sites + 63) / 64 * 64;
  return np <= 1024 ? np : -1;
}

// Tensor maps of the GEMM operands (kernel parameter space, __grid_constant__): every buffer is one 2-D f64 tensor
// [rows = matrices x NP][cols = NP], box 16 x 16, 128-byte swizzle.  G / T hold all chains' matrices stacked as rows.
struct L2TmaMaps {
  alignas(64) CUtensorMap G, T, E, Et, Ei, Eit;
  const double* pG; const double* pT; const double* pE; const double* pEt; const double* pEi; const double* pEit;
  long long n_elems;      // doubles in the G (and T) buffer
  int valid;
};

struct L2Workspace {
  double* T = nullptr;      // [chain][2][NP][NP] second matrix buffer (two-GEMM wrap, running product)
  int kd = 0;               // delay depth the shared-memory budget allows
  int tmem_mode = 0;        // tensor-memory slice path that fits the shared-memory region: 0 none, 1 NP <= 256, 2 384 < NP <= 512, 3 512 < NP <= 768
  int gemm_stages = 4;      // operand-ring depth of the 64 x 192 GEMMs (tmem_mode 2 / 3 kernels)
  size_t smem = 0;          // dynamic shared memory per CTA
  L2TmaMaps maps;           // tensor maps of the GEMM operands (built at the first launch)
  int last_cluster = 1;     // CTAs per chain of the most recent launch
};

// ---- one chain on a thread-block cluster (strong scaling: fewer chains than SMs) --------------------------------------------
// With fewer chains than SMs one CTA per chain leaves most of the GPU idle (37 chains: a quarter of a B200).  The sweep kernel can
// instead run one chain on a cluster of CS = 2 / 4 / 8 CTAs (one per SM, co-scheduled by the hardware):
//   * the two phases that carry the work - GEMM tiles (wrap, sweep-start product) and the rows of the delayed-update flush - are
//     SPLIT over the CTAs of the cluster;
//   * the latency-bound build of a flip (acceptance scan, row / column rebuild, Sherman-Morrison vectors) is REPLICATED: every CTA
//     reads the same G0 rows / columns, performs the same arithmetic in the same order and therefore takes the same decisions
//     and holds the same (e, c) histories - which it needs anyway for its share of the flush.  No vector is exchanged; the only
//     communication is the hardware cluster barrier (barrier.cluster arrive.release / wait.acquire) that orders the global-memory
//     writes of one phase before the reads of the next;
//   * the serial sweep-start inverse runs on rank 0 alone (6 % of the recompute).
// Every element of G still sees the reference's operations in the reference's order: bit-identical to the one-CTA kernel.
// A pointer that reaches a __noinline__ subroutine inside a by-value argument is a plain generic value to the compiler: every
// access through it becomes a generic LD.E / ST.E.  Round-tripping it through its real state space tells the address-space inference
// which one it is (LDS / STS, LDG / STG again).
template <class T> __device__ __forceinline__ T* as_shared(T* p) {
  return reinterpret_cast<T*>(__cvta_shared_to_generic(__cvta_generic_to_shared(const_cast<void*>(static_cast<const void*>(p)))));
}
template <class T> __device__ __forceinline__ T* as_global(T* p) {
  return reinterpret_cast<T*>(__cvta_global_to_generic(__cvta_generic_to_global(const_cast<void*>(static_cast<const void*>(p)))));
}
#ifndef LQMC_CLUSTER_LDCG
#define LQMC_CLUSTER_LDCG 1
#endif
template <bool CL>
__device__ __forceinline__ double l2_ld_peer(const double* p) {
  if (CL && LQMC_CLUSTER_LDCG) return __ldcg(p);        // a cluster peer may have written the line: read it from L2
  return *p;
}
__device__ __forceinline__ void l2_cluster_sync(int cs) {
#ifdef LQMC_NO_CLUSTER
  __syncthreads(); return;
#endif
  if (cs > 1) {
    __threadfence();                                          // this CTA's global writes (flush, epilogue) before the release
    asm volatile("fence.proxy.async;" ::: "memory");          // ... also for the TMA (async-proxy) reads of the peers
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    __threadfence();
  } else {
    __syncthreads();
  }
}

struct L2Smem {
  const L2TmaMaps* maps = nullptr;     // set by the sweep kernel when the launch carries tensor maps
  int cs = 1, crank = 0;               // cluster size and this CTA's rank in it (1 / 0: one CTA per chain)
  // vector phase
  double* U;      // [2 spin][KD][NP]   e vectors of the delayed flips
  double* W;      // [2 spin][KD][NP]   c vectors
  double* d;      // [2 buf][2 spin][NP] current diagonal, double-buffered across flips
  double* u;      // [NP]               uniforms of the slice
  double* red_v;  // [16]
  double* hist;   // [64]  (TMEM path: history of the flipped site, 2 x L2_KDT; [63] doubles as the TMEM base-address slot)
  double* ring;   // [L2_RING][2][L2_KDT]  (TMEM path, NP <= 256: c histories of the next candidate sites, slot = site & 7)
  int* red_i;     // [16]
  int8_t* h;      // [NP]
  int8_t* hn;     // [NP]
  // GEMM phase (aliases U / W)
  double* pa;     // [L2_STAGES][BK][BM + 4]
  double* pb;     // [L2_STAGES][BK][BN + 4]
  uint64_t* full; // [L2_STAGES] mbarriers: "panel landed"
  uint32_t pipe_iter = 0;   // panels consumed so far by this CTA (uniform across threads): stage and phase parity
  // ns = matrices the CTA works on at once (2: both spins of a chain; 1: the single-matrix inverse of stab.cuh)
  // with_gemm: reserve the operand-stage region of the GEMMs (the single-matrix inverse of stab.cuh has no GEMM and must not pay for it)
  __device__ L2Smem(unsigned char* base, int NP, int KD, int ns = 2, bool with_gemm = true) {
    U = reinterpret_cast<double*>(base);
    W = U + (size_t)ns * KD * NP;
    pa = U;
    pb = pa + L2_STAGES * L2_BK * (L2_BM + 4);
    double* tail = W + (size_t)ns * KD * NP;
    const size_t gemm_end = L2_GEMM_DOUBLES;
    if (with_gemm && (size_t)2 * ns * KD * NP < gemm_end) tail = U + gemm_end;
    d = tail;
    u = d + 4 * NP;
    red_v = u + NP;
    hist = red_v + 16;
    ring = hist + 64;
    full = reinterpret_cast<uint64_t*>(ring + L2_RING * 2 * L2_KDT);     // [0..3] "panel landed", [4..7] "panel consumed"
    red_i = reinterpret_cast<int*>(full + 16);    // full[0..7] "panel landed", full[8..15] "panel consumed"
    h = reinterpret_cast<int8_t*>(red_i + 16);
    hn = h + NP;
  }
};

inline size_t l2_smem_bytes(int NP, int KD, int ns = 2, bool with_gemm = true) {
  size_t vec = (size_t)2 * ns * KD * NP;
  const size_t gemm = L2_GEMM_DOUBLES;
  if (with_gemm && vec < gemm) vec = gemm;
  return (vec + 5 * (size_t)NP + 16 + 64 + L2_RING * 2 * L2_KDT + 16) * sizeof(double) + 16 * sizeof(int) + 2 * (size_t)NP + 16;
}

// ---- tiled GEMM:  C = A * B  with A given k-major (At[k*NP + i] = A[i][k]) and B row-major ------------------
// Epilogue: C[i][j] *= rs(i) * cs(j) (+ 1 on the diagonal if add_identity); stored row-major or transposed.
struct L2Epilogue {
  const int8_t* hrow = nullptr;   // field column for the row scale, nullptr = 1
  const int8_t* hcol = nullptr;   // field column for the column scale, nullptr = 1
  bool row_inv = false;           // row scale uses exp(+sigma lamb h) instead of exp(-sigma lamb h)
  bool col_inv = false;
  bool transposed_out = false;
  bool add_identity = false;
  // Tile order.  Every tile re-fetches its operand panels (nothing is kept between tiles), so an operand is read NP / BM (B) or
  // NP / BN (A) times per product.  The chains' own matrices do not fit L2 together (cfg4: 592 MB, cfg5: 785 MB against 126 MB);
  // the exp(-+dtau K) matrices are shared and do.  Row tiles outer (default) re-reads an A row tile back to back - right when A is
  // the chain's matrix; column tiles outer does the same for B (the first product of the wrap, B = G).
  bool col_outer = false;
};

// (dmma884, the FP64 tensor-core MMA wrapper, lives in sweep_reg.cuh)
constexpr int L2_LDA = L2_BM + 4;    // panel row strides = 4 mod 16 doubles: the 4 k-rows a fragment load touches
constexpr int L2_LDB = L2_BN + 4;    // fall into disjoint bank groups (conflict-free LDS.64)

// ---- TMA bulk-copy plumbing (cp.async.bulk + mbarrier complete_tx; SASS: UBLKCP / SYNCS) --------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// One operand panel = BK k-rows of the k-major left operand (BM doubles each) + BK rows of the right operand
// (BN doubles each).  Warp 0 issues it: lane r < BK copies A row r and B row r as two bulk transfers that complete
// on the stage's mbarrier.  All addresses are per-lane registers set up once per block tile.
struct L2PanelIssue {
  const double* srcA;   // At + lane*NP + i0     (+ k0*NP per panel)
  const double* srcB;   // B  + lane*NP + j0
  uint32_t dstA, dstB;  // shared addresses of row `lane` in stage 0
  uint32_t bar;         // shared address of full[0]
};
__device__ __forceinline__ void l2_issue_panel(const L2PanelIssue& pi, int NP, int k0, int stage, int lane) {
  const uint32_t bar = pi.bar + 8u * stage;
  if (lane == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(L2_BK * (L2_BM + L2_BN) * sizeof(double))) : "memory");
  __syncwarp();
  if (lane < L2_BK) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     pi.dstA + (uint32_t)(stage * L2_BK * L2_LDA * sizeof(double))),
                 "l"(pi.srcA + (size_t)k0 * NP), "r"((uint32_t)(L2_BM * sizeof(double))), "r"(bar)
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     pi.dstB + (uint32_t)(stage * L2_BK * L2_LDB * sizeof(double))),
                 "l"(pi.srcB + (size_t)k0 * NP), "r"((uint32_t)(L2_BN * sizeof(double))), "r"(bar)
                 : "memory");
  }
}

// Compiled as a subroutine of its own (__noinline__, every argument by value): the register allocation of the GEMM main
// loop must not depend on what else lives in the sweep kernel - measured: the same source ran 10 % slower inlined next to
// the tensor-memory slice path.
struct L2GemmCtx { double* pa; double* pb; uint64_t* full; uint32_t* pipe_iter; double exp_pl, exp_ml; };
__device__ __forceinline__ double hs_v2(int8_t h, int spin, bool inv, double exp_pl, double exp_ml) {
  return (((h > 0) != (spin != 0)) != inv) ? exp_ml : exp_pl;
}
__device__ __noinline__ void l2_gemm_sub(const double* __restrict__ At, const double* __restrict__ B, double* __restrict__ Cout, int NP,
                                         int spin, const L2Epilogue ep, const L2GemmCtx sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;            // 2 x 4 warps, 32 x 32 warp tiles
  const int lr = lane >> 2, lk = lane & 3;
  const int nk = NP / L2_BK;
  const double* pa = sm.pa;
  const double* pb = sm.pb;
  const uint32_t pa_u32 = smem_u32(sm.pa), pb_u32 = smem_u32(sm.pb), bar_u32 = smem_u32(sm.full);
  // operands may have been written by this CTA's ordinary stores (previous epilogue, flush) and the panel
  // region by ordinary shared stores (U / W): order them before the async-proxy copies
  fence_proxy_async();
  __syncthreads();
  for (int i0 = 0; i0 < NP; i0 += L2_BM) {
    for (int j0 = 0; j0 < NP; j0 += L2_BN) {
      double acc[4][4][2];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
      // field bytes of this thread's rows / column pairs for the epilogue scales: in flight during the main loop
      int hr[4] = {1, 1, 1, 1};
      unsigned hc[4] = {0x0101u, 0x0101u, 0x0101u, 0x0101u};
      if (ep.hrow) {
#pragma unroll
        for (int m = 0; m < 4; ++m) hr[m] = ep.hrow[i0 + 32 * wm + 8 * m + lr];
      }
      if (ep.hcol && j0 + 32 * wn < NP) {
#pragma unroll
        for (int n = 0; n < 4; ++n) hc[n] = *reinterpret_cast<const unsigned short*>(ep.hcol + j0 + 32 * wn + 8 * n + 2 * lk);
      }
#if 0   // (round 1's row-wise cp.async.bulk staging; superseded by l2_gemm_tma_sub below)
      const uint32_t it0 = *sm.pipe_iter;
      L2PanelIssue pi;
      pi.srcA = At + (size_t)lane * NP + i0;
      pi.srcB = B + (size_t)lane * NP + j0;
      pi.dstA = pa_u32 + (uint32_t)(lane * L2_LDA * sizeof(double));
      pi.dstB = pb_u32 + (uint32_t)(lane * L2_LDB * sizeof(double));
      pi.bar = bar_u32;
      if (warp == 0) {
#pragma unroll
        for (int s0 = 0; s0 < L2_STAGES; ++s0)
          if (s0 < nk) l2_issue_panel(pi, NP, s0 * L2_BK, (it0 + s0) % L2_STAGES, lane);
      }
      for (int kp = 0; kp < nk; ++kp) {
        const uint32_t it = it0 + kp;
        const int st = it % L2_STAGES;
        mbar_wait_u32(bar_u32 + 8u * st, (it / L2_STAGES) & 1u);
        const double* ap = pa + st * L2_BK * L2_LDA + lk * L2_LDA + 32 * wm + lr;
        const double* bp = pb + st * L2_BK * L2_LDB + lk * L2_LDB + 32 * wn + lr;
#pragma unroll
        for (int k4 = 0; k4 < L2_BK / 4; ++k4) {
          double a[4], b[4];
#pragma unroll
          for (int m = 0; m < 4; ++m) a[m] = ap[4 * k4 * L2_LDA + 8 * m];
#pragma unroll
          for (int n = 0; n < 4; ++n) b[n] = bp[4 * k4 * L2_LDB + 8 * n];
#pragma unroll
          for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int n = 0; n < 4; ++n) dmma884(acc[m][n], a[m], b[n]);
        }
        __syncthreads();                     // every warp is done with stage st: refill it
        if (warp == 0 && kp + L2_STAGES < nk) l2_issue_panel(pi, NP, (kp + L2_STAGES) * L2_BK, st, lane);
      }
      __syncthreads();
      if (tid == 0) *sm.pipe_iter = it0 + nk;
      __syncthreads();
#else
      // LDGSTS ring: every thread copies 2 (A) + 4 (B) 16-byte chunks per panel; addresses set up once per tile
      const double* srcA = At + (size_t)(tid >> 5) * NP + i0 + 2 * (tid & 31);          // rows tid/32 (+8), chunk tid%32
      const double* srcB = B + (size_t)(tid >> 6) * NP + j0 + 2 * (tid & 63);           // rows tid/64 (+4 q), chunk tid%64
      const bool b_ok = j0 + 2 * (tid & 63) < NP;          // half tile at the right edge (NP = 64 mod 128): no copy, no compute
      const bool w_ok = j0 + 32 * wn < NP;                 // warp-uniform
      double* dstA = sm.pa + (tid >> 5) * L2_LDA + 2 * (tid & 31);
      double* dstB = sm.pb + (tid >> 6) * L2_LDB + 2 * (tid & 63);
      auto issue = [&](int kpanel, int stage) {
        const size_t koff = (size_t)kpanel * L2_BK * NP;
#pragma unroll
        for (int q = 0; q < 2; ++q)
          __pipeline_memcpy_async(dstA + stage * L2_BK * L2_LDA + q * 8 * L2_LDA, srcA + koff + (size_t)q * 8 * NP, 16);
        if (b_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            __pipeline_memcpy_async(dstB + stage * L2_BK * L2_LDB + q * 4 * L2_LDB, srcB + koff + (size_t)q * 4 * NP, 16);
        }
      };
#pragma unroll
      for (int s0 = 0; s0 < L2_STAGES - 1; ++s0) {
        if (s0 < nk) issue(s0, s0);
        __pipeline_commit();
      }
      for (int kp = 0; kp < nk; ++kp) {
        const int st = kp % L2_STAGES;
        __pipeline_wait_prior(L2_STAGES - 2);        // this thread's copies of panel kp have landed
        __syncthreads();                             // everyone's have; and every warp is done with panel kp-1
        if (kp + L2_STAGES - 1 < nk) issue(kp + L2_STAGES - 1, (kp + L2_STAGES - 1) % L2_STAGES);
        __pipeline_commit();
        const double* ap = pa + st * L2_BK * L2_LDA + lk * L2_LDA + 32 * wm + lr;
        const double* bp = pb + st * L2_BK * L2_LDB + lk * L2_LDB + 32 * wn + lr;
        if (w_ok) {
#pragma unroll
          for (int k4 = 0; k4 < L2_BK / 4; ++k4) {
            double a[4], b[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) a[m] = ap[4 * k4 * L2_LDA + 8 * m];
#pragma unroll
            for (int n = 0; n < 4; ++n) b[n] = bp[4 * k4 * L2_LDB + 8 * n];
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
              for (int n = 0; n < 4; ++n) dmma884(acc[m][n], a[m], b[n]);
          }
        }
      }
      __pipeline_wait_prior(0);
      __syncthreads();                               // panels free for the next tile
#endif
      // epilogue: element (row, col) = acc[m][n][s]; the field bytes were fetched before the main loop
      if (j0 + 32 * wn >= NP) continue;        // half tile at the right edge: this warp owns no columns (warp-uniform)
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int row = i0 + 32 * wm + 8 * m + lr;
        const double rs = hs_v2((int8_t)hr[m], spin, ep.row_inv, sm.exp_pl, sm.exp_ml);
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const int col0 = j0 + 32 * wn + 8 * n + 2 * lk;
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            double v = acc[m][n][s2];
            if (ep.hrow) v *= rs;
            if (ep.hcol) v *= hs_v2((int8_t)(s2 ? (hc[n] >> 8) : (hc[n] & 0xff)), spin, ep.col_inv, sm.exp_pl, sm.exp_ml);
            if (ep.add_identity && row == col0 + s2) v += 1.0;
            acc[m][n][s2] = v;
          }
          if (!ep.transposed_out) {
            *reinterpret_cast<double2*>(Cout + (size_t)row * NP + col0) = make_double2(acc[m][n][0], acc[m][n][1]);
          } else {
            Cout[(size_t)col0 * NP + row] = acc[m][n][0];
            Cout[(size_t)(col0 + 1) * NP + row] = acc[m][n][1];
          }
        }
      }
    }
  }
  __syncthreads();
}

// ---- the same GEMM with TMA-staged operands (north_star item 2) ---------------------------------------------------------------
// Operand panels arrive by cp.async.bulk.tensor (SASS UTMALDG): one elected thread issues 2 KB boxes of 16 k-rows x 16 doubles
// through a 2-D tensor map of the whole buffer (all chains' matrices stacked as rows) with the 128-byte swizzle, four boxes for the
// 64-wide left panel and eight for the 128-wide right panel; they complete on the stage's "full" mbarrier (expect_tx = 24 KB).
// Nobody but that thread spends issue slots or LSU wavefronts on staging, and the panels are dense (no padding columns).
//
// Bank conflicts: a DMMA fragment load reads, per half warp, 4 k-rows x 4 consecutive doubles.  In a dense 128-byte-row box all
// k-rows start in bank 0; the hardware swizzle XORs the 16-byte chunk index with (row & 7), which separates rows r and r' only if
// (r ^ r') touches bit 1 or 2 of the chunk index - so a lane's k index within a k4 step is mapped to rows {0,2,4,6} (+1 for odd
// steps, +8 for the second half of the panel) instead of {0,1,2,3}: the contraction index may be visited in any order as long as
// both operands use the same one.  Every fragment load is then conflict-free (checked: l1tex__data_bank_conflicts in
// profiles/r02g_*).
//
// Pipeline: L2_TMA_STAGES stages, "full" (TMA -> consumers, tx-count) and "empty" (consumers -> producer, one arrive per warp)
// mbarriers, no __syncthreads in the main loop.  The panel sequence runs across the block tiles of the GEMM: while the warps are
// in a tile's epilogue the first panels of the next tile are already in flight.  Out-of-range boxes of the half tile at the right
// edge (NP = 64 mod 128) are zero-filled by the TMA unit.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
struct L2TmaOperand { const void* map; int row0; };        // tensor map of the buffer and the matrix's first row in it
struct L2GemmTmaCtx { uint32_t stage0; uint32_t bars; uint32_t* pipe_iter; const unsigned char* stage_ptr; double exp_pl, exp_ml; int cs, crank; };
// CL: the launch runs one chain per thread-block cluster.  A template flag, not a run-time test: with the cluster code compiled into
// the one-CTA kernel its slice phase ran 20 % slower (1.77 -> 2.12 ms; cause not isolated - not the barrier instructions, not the
// L2-only loads, not the run-time chunk ranges), so the one-CTA instantiation contains none of it.
// NT: 8-column fragments per warp in N.  4: block tile 64 x 128 (two CTAs per SM, 128 registers).  6: block tile 64 x 192, warp tile
// 32 x 48 - the one-CTA-per-SM sizes (NP > 384, 255 registers): 10 fragment loads per 24 DMMAs instead of 8 per 16, and 576 = 3 x 192
// has no half tile at the right edge (with 128-column tiles 9 of its 45 tiles were half empty but took a full tile's time).
template <bool CL, int NT>
__device__ __noinline__ void l2_gemm_tma_sub(const L2TmaOperand opA, const L2TmaOperand opB, double* __restrict__ Cout, int NP, int spin,
                                             const L2Epilogue ep, const L2GemmTmaCtx sm) {
  const int ccs = CL ? sm.cs : 1, ccrank = CL ? sm.crank : 0;
  // ring depth: 4 stages of 24 KB with the 64 x 128 tile (two CTAs per SM); the one-CTA-per-SM kernels take as many 32 KB stages as
  // the U / W region holds (up to L2_TMA_STAGES_MAX; the host writes the count next to the panel counter)
  const int S = (NT == 4) ? L2_TMA_STAGES : (int)sm.pipe_iter[1];
  constexpr int BN = 32 * NT;                                                      // 4 warps in N x NT fragments of 8 columns
  static_assert(NT == 4 || NT == 6, "warp tile 32 x 32 or 32 x 48");
  constexpr uint32_t STAGE_BYTES = (L2_BM + BN) * L2_BK * sizeof(double);          // 24 KB / 32 KB
  constexpr uint32_t A_BYTES = L2_BM * L2_BK * sizeof(double);                     // 8 KB: boxes 0..3 of a stage, then 8 B boxes
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;            // 2 x 4 warps, 32 x 32 warp tiles
  const int lr = lane >> 2, lk = lane & 3;
  const int nk = NP / L2_BK;
  const int tiles_n = (NP + BN - 1) / BN, tiles_m = NP / L2_BM, n_tiles_all = tiles_m * tiles_n;
  // block tiles crank, crank + cs, ... of the GEMM are this CTA's (cs = 1: all of them)
  const int n_tiles = (n_tiles_all - ccrank + ccs - 1) / ccs;
  const int n_panels = n_tiles * nk;
  const uint32_t full0 = sm.bars, empty0 = sm.bars + 8u * L2_TMA_STAGES_MAX;
  const unsigned char* const stage_ptr = sm.stage_ptr;
  __builtin_assume(__isShared(stage_ptr));
  // generic-proxy writes of this CTA (previous epilogue, flush, the U / W region the stages alias) before async-proxy traffic
  fence_proxy_async();
  __syncthreads();
  // the barriers are initialised once per kernel; panels are numbered across the GEMM calls of the launch (stage and phase parity)
  const uint32_t q0 = *sm.pipe_iter;
  // producer: one thread; panel q = (tile q / nk, k-panel q % nk) goes to stage (q0 + q) % S
  auto issue = [&](int q) {
    const int tl = q / nk, kp = q - tl * nk;
    const int t = ccrank + tl * ccs;
    const int ti = ep.col_outer ? t % tiles_m : t / tiles_n, tj = ep.col_outer ? t / tiles_m : t - ti * tiles_n;
    const int st = (q0 + q) % S;
    const uint32_t dst = sm.stage0 + (uint32_t)st * STAGE_BYTES, bar = full0 + 8u * st;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(STAGE_BYTES) : "memory");
#pragma unroll
    for (int b = 0; b < L2_BM / 16; ++b) tma_load_2d(dst + b * L2_TMA_SUB, opA.map, ti * L2_BM + 16 * b, opA.row0 + kp * L2_BK, bar);
#pragma unroll
    for (int b = 0; b < BN / 16; ++b)
      tma_load_2d(dst + A_BYTES + b * L2_TMA_SUB, opB.map, tj * BN + 16 * b, opB.row0 + kp * L2_BK, bar);
  };
  // the producer is ONE elected lane of warp 0 (elect.sync: the compiler then knows the TMA operands are uniform and emits the
  // UTMALDG straight from uniform registers instead of a per-lane loop)
  auto elected = [&]() -> bool {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
  };
  if (warp == 0) {
    if (elected())
      for (int q = 0; q < S - 1 && q < n_panels; ++q) issue(q);
    __syncwarp();
  }
  // per-lane fragment offsets inside a stage (bytes): box of the fragment's 16-column group, k-row 2 lk (+ step parity, + 8 for
  // the second half), 16-byte chunk XOR (row & 7), 8-byte half
  uint32_t offA[4], offB[NT];                      // [m] / [n] for step parity 0, k-half 0; the other three steps are derived
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int e = 8 * (m & 1) + lr;                // element inside the 16-wide box
    offA[m] = (uint32_t)((2 * wm + (m >> 1)) * L2_TMA_SUB + (2 * lk) * 128 + (((e >> 1) ^ (2 * lk)) << 4) + ((e & 1) << 3));
  }
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    const int e = 8 * (n & 1) + lr;
    offB[n] = A_BYTES + (uint32_t)(((NT / 2) * wn + (n >> 1)) * L2_TMA_SUB + (2 * lk) * 128 + (((e >> 1) ^ (2 * lk)) << 4) + ((e & 1) << 3));
  }
  int q = 0;
  for (int tl = 0; tl < n_tiles; ++tl) {
    const int t = ccrank + tl * ccs;
    const int ti = ep.col_outer ? t % tiles_m : t / tiles_n, tj = ep.col_outer ? t / tiles_m : t - ti * tiles_n;
    const int i0 = ti * L2_BM, j0 = tj * BN, jw = j0 + 8 * NT * wn;     // jw: first column of this warp
    const bool w_ok = jw < NP;                           // partial tile at the right edge: this warp owns no columns (warp-uniform)
    double acc[4][NT][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < NT; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    // field bytes of this thread's rows / column pairs for the epilogue scales: in flight during the main loop
    int hr[4] = {1, 1, 1, 1};
    unsigned hc[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) hc[n] = 0x0101u;
    if (ep.hrow) {
#pragma unroll
      for (int m = 0; m < 4; ++m) hr[m] = ep.hrow[i0 + 32 * wm + 8 * m + lr];
    }
    if (ep.hcol && w_ok) {
#pragma unroll
      for (int n = 0; n < NT; ++n)
        if (NT == 4 || jw + 8 * n < NP) hc[n] = *reinterpret_cast<const unsigned short*>(ep.hcol + jw + 8 * n + 2 * lk);
    }
    for (int kp = 0; kp < nk; ++kp, ++q) {
      const uint32_t qg = q0 + (uint32_t)q;
      const int st = qg % S;
      mbar_wait_u32(full0 + 8u * st, (qg / S) & 1u);
      // ordinary (non-volatile) loads: the compiler may run the next step's fragment loads under the current step's DMMAs, but not
      // above the barrier wait (a volatile asm with a memory clobber)
      const unsigned char* base = stage_ptr + (size_t)st * STAGE_BYTES;
      if (w_ok) {
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
          // step s4: k-rows 8 (s4 >> 1) + 2 lk + (s4 & 1); an odd row flips bit 0 of the chunk XOR, the second half adds 1 KB
          const uint32_t delta = (uint32_t)((s4 >> 1) * 1024 + (s4 & 1) * 128), flip = (uint32_t)((s4 & 1) << 4);
          double a[4], b[NT];
#pragma unroll
          for (int m = 0; m < 4; ++m) a[m] = *reinterpret_cast<const double*>(base + ((offA[m] + delta) ^ flip));
#pragma unroll
          for (int n = 0; n < NT; ++n) b[n] = *reinterpret_cast<const double*>(base + ((offB[n] + delta) ^ flip));
#pragma unroll
          for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int n = 0; n < NT; ++n) dmma884(acc[m][n], a[m], b[n]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_u32(empty0 + 8u * st);          // this warp is done with the stage
      // The producer refills the stage panel q-1 sat in.  It does so AFTER its own warp's work on panel q: by then the other
      // warps have long released panel q-1, so the wait does not hold warp 0 back (issued before the compute it made warp 0 the
      // last warp of every panel: 0.93 ms per wrap against 0.78 with the cp.async ring, one CTA per SM).
      // (r2) The producer role rotates over the warps: with warp 0 issuing every panel (an empty-barrier wait, 12 - 16 TMA instructions
      // and their coordinates: several hundred clocks of one warp's dependent instruction stream) warp 0's loop was the slowest
      // stage of the pipeline and the other seven warps waited for it on the "full" barrier - 28 % of the stall samples at N = 576,
      // whatever the ring depth and the number of chains.
      if (warp == (LQMC_L2_ROTATE_PRODUCER ? (q & 7) : 0) && q + S - 1 < n_panels) {
        if (elected()) {
          if (q >= 1) mbar_wait_u32(empty0 + 8u * ((qg - 1) % S), ((qg - 1) / S) & 1u);
          issue(q + S - 1);
        }
        __syncwarp();
      }
    }
    if (!w_ok) continue;
    // epilogue: element (row, col) = acc[m][n][s]
    double cs[NT][2];
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      cs[n][0] = ep.hcol ? hs_v2((int8_t)(hc[n] & 0xff), spin, ep.col_inv, sm.exp_pl, sm.exp_ml) : 1.0;
      cs[n][1] = ep.hcol ? hs_v2((int8_t)(hc[n] >> 8), spin, ep.col_inv, sm.exp_pl, sm.exp_ml) : 1.0;
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int row = i0 + 32 * wm + 8 * m + lr;
      const double rs = hs_v2((int8_t)hr[m], spin, ep.row_inv, sm.exp_pl, sm.exp_ml);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const int col0 = jw + 8 * n + 2 * lk;
        if (NT != 4 && jw + 8 * n >= NP) continue;       // fragment past the right edge (zero-filled operands): nothing to store
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          double v = acc[m][n][s2];
          if (ep.hrow) v *= rs;
          if (ep.hcol) v *= cs[n][s2];
          if (ep.add_identity && row == col0 + s2) v += 1.0;
          acc[m][n][s2] = v;
        }
        if (!ep.transposed_out) {
          *reinterpret_cast<double2*>(Cout + (size_t)row * NP + col0) = make_double2(acc[m][n][0], acc[m][n][1]);
        } else {
          Cout[(size_t)col0 * NP + row] = acc[m][n][0];
          Cout[(size_t)(col0 + 1) * NP + row] = acc[m][n][1];
        }
      }
    }
  }
  __syncthreads();          // all panels consumed, all results stored: the stages may be reused as U / W
  if (tid == 0) *sm.pipe_iter = q0 + (uint32_t)n_panels;          // ordered before the next call's read by its entry barrier
}

__device__ __forceinline__ L2TmaOperand l2_tma_operand(const L2TmaMaps& m, const double* ptr, int NP) {
  if (ptr == m.pE) return L2TmaOperand{&m.E, 0};
  if (ptr == m.pEt) return L2TmaOperand{&m.Et, 0};
  if (ptr == m.pEi) return L2TmaOperand{&m.Ei, 0};
  if (ptr == m.pEit) return L2TmaOperand{&m.Eit, 0};
  // a matrix of some chain inside the G or the T buffer: NP x NP blocks, so its first row is the element offset / NP
  if (ptr >= m.pG && ptr < m.pG + m.n_elems) return L2TmaOperand{&m.G, (int)((ptr - m.pG) / NP)};
  return L2TmaOperand{&m.T, (int)((ptr - m.pT) / NP)};
}

// One staging variant per binary: with both compiled in, every GEMM call site marshals two argument sets around two calls and the
// one-launch sweep ran 2 % slower (274 vs 269 ms) whichever path was taken at run time.
template <bool CL, int NT = 4>
__device__ __forceinline__ void l2_gemm(const double* __restrict__ At, const double* __restrict__ B, double* __restrict__ Cout, int NP, int spin,
                                        const L2Epilogue& ep, const SweepParams& p, L2Smem& sm) {
#if LQMC_L2_STAGING_TMA
  L2GemmTmaCtx c;
  c.stage0 = smem_u32(sm.U); c.bars = smem_u32(sm.full); c.pipe_iter = reinterpret_cast<uint32_t*>(sm.hist + 62);
  c.stage_ptr = reinterpret_cast<const unsigned char*>(sm.U);
  c.exp_pl = p.exp_pl; c.exp_ml = p.exp_ml; c.cs = sm.cs; c.crank = sm.crank;
  l2_gemm_tma_sub<CL, NT>(l2_tma_operand(*sm.maps, At, NP), l2_tma_operand(*sm.maps, B, NP), Cout, NP, spin, ep, c);
  if (CL && sm.cs > 1) l2_cluster_sync(sm.cs);    // every tile of the product is in memory before any CTA reads it as an operand
#else
  L2GemmCtx c;
  c.pa = sm.pa; c.pb = sm.pb; c.full = sm.full; c.pipe_iter = reinterpret_cast<uint32_t*>(sm.hist + 62);
  c.exp_pl = p.exp_pl; c.exp_ml = p.exp_ml;
  l2_gemm_sub(At, B, Cout, NP, spin, ep, c);
#endif
}

// ---- G0 <- G0 - sum_m U_m W_m^T for both spins (the delayed block update) -----------------------------------
// A streaming pass over G (read + write 16 N^2 bytes for both spins): HBM-bound once the chains' matrices exceed
// L2.  Block tile 32 rows x 128 columns: warp w owns rows 4w..4w+3, lane l the column pairs {2l, 2l+1} and
// {64+2l, 64+2l+1}, so every global access of a warp is one fully coalesced 512-byte row segment, the U
// fragment is a warp-wide broadcast and the W fragment a conflict-free 512-byte read.  The next tile's loads
// are kept in flight while the nd updates are applied to the current one.
template <bool EXACT, int NS = 2>
__device__ void l2_flush(double* __restrict__ Gc, int NP, int nd, L2Smem& sm, int KD) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_tiles = NS * (NP / 32) * ((NP + 127) / 128);     // NP = 64 mod 128: the last tile of a row band is half a tile
  double* const thread_base = Gc + (size_t)(4 * warp) * NP + 2 * lane;
  // tile walk (spin, i0, j0) kept as running offsets: no integer division in the loop (the XU pipe is 1/4 rate)
  struct Walk { int spin, i0, j0; };
  auto advance = [&](Walk& w) {
    w.j0 += 128;
    if (w.j0 >= NP) { w.j0 = 0; w.i0 += 32; if (w.i0 == NP) { w.i0 = 0; w.spin += 1; } }
  };
  auto tile_ptr_w = [&](const Walk& w) -> double* { return thread_base + ((size_t)w.spin * NP + w.i0) * NP + w.j0; };
  auto load_tile4 = [&](const double* base, double (&g)[4][4], bool full) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const double2 v = (q == 0 || full) ? *reinterpret_cast<const double2*>(base + (size_t)a * NP + 64 * q) : make_double2(0.0, 0.0);
        g[a][2 * q] = v.x; g[a][2 * q + 1] = v.y;
      }
  };
  Walk wc{0, 0, 0}, wn{0, 0, 0};
  double nxt[4][4];
  load_tile4(tile_ptr_w(wn), nxt, wn.j0 + 64 < NP);
  for (int t = 0; t < n_tiles; ++t) {
    double g[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) g[a][b] = nxt[a][b];
    double* base = tile_ptr_w(wc);
    const bool full = wc.j0 + 64 < NP;
    advance(wn);
    if (t + 1 < n_tiles) load_tile4(tile_ptr_w(wn), nxt, wn.j0 + 64 < NP);
    const double* U = sm.U + (size_t)wc.spin * KD * NP + wc.i0 + 4 * warp;
    const double* W = sm.W + (size_t)wc.spin * KD * NP + wc.j0 + 2 * lane;
    advance(wc);
    // fragments of update m+1 are fetched while update m is applied (shared-memory latency off the critical path)
    double e[4], c[4];
    auto load_frag = [&](int m, double (&ef)[4], double (&cf)[4]) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const double2 v = *reinterpret_cast<const double2*>(U + (size_t)m * NP + 2 * q);
        ef[2 * q] = v.x; ef[2 * q + 1] = v.y;
        const double2 w = *reinterpret_cast<const double2*>(W + (size_t)m * NP + 64 * q);
        cf[2 * q] = w.x; cf[2 * q + 1] = w.y;
      }
    };
    load_frag(0, e, c);
    for (int m = 0; m < nd; ++m) {
      double en[4], cn[4];
      load_frag((m + 1 < nd) ? m + 1 : m, en, cn);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) g[a][b] = rank1<EXACT>(g[a][b], e[a], c[b]);
#pragma unroll
      for (int a = 0; a < 4; ++a) { e[a] = en[a]; c[a] = cn[a]; }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int q = 0; q < 2; ++q)
        if (q == 0 || full) *reinterpret_cast<double2*>(base + (size_t)a * NP + 64 * q) = make_double2(g[a][2 * q], g[a][2 * q + 1]);
  }
  __syncthreads();
}

// ---- the N proposals of one slice with delayed updates ------------------------------------------------------
// QMAX = entries per thread in the vector phases: 1 for NP <= 256 (no dead registers for larger lattices), else L2_MAXQ
template <bool EXACT, bool PHYS, int QMAX>
__device__ void l2_propose_slice(double* __restrict__ Gc, int NP, int KD, L2Smem& sm, const SweepParams& p, long long trace_base,
                                 int& n_accepted) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int N = p.n_sites;
  // diagonal of both spins
  for (int spin = 0; spin < 2; ++spin)
    for (int j = tid; j < NP; j += L2_THREADS) sm.d[spin * NP + j] = Gc[(size_t)spin * NP * NP + (size_t)j * NP + j];
  __syncthreads();
  int nd = 0;
  int i0 = 0;
#ifdef LQMC_PHASE_CLOCKS
  long long tk_scan = 0, tk_build = 0, tk_flush = 0, tk0 = clock64();
  long long tk_b1 = 0, tk_b2 = 0, tk_b3 = 0;     // build split: global loads landed / pending updates applied / vectors written
#endif
  int cur = 0;                  // diagonal buffer the scan reads; a flip writes the other one (slower warps may still be scanning)
  while (i0 < N) {
    const double* dcur = sm.d + cur * 2 * NP;
    double* dnxt = sm.d + (cur ^ 1) * 2 * NP;
    // lane-parallel scan of sites i0 .. i0+31
    const int i = i0 + lane;
    bool acc = false;
    double gu = 0.0, gd = 0.0, ratio = 0.0;
    int8_t h = 1;
    if (i < N) {
      h = sm.h[i];
      gu = dcur[i];
      gd = dcur[NP + i];
      const double fu = (h > 0) ? p.f_p2 : p.f_m2;
      const double fd = (h > 0) ? p.f_m2 : p.f_p2;
      const double du = __dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, gu), fu));
      const double dd = __dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, gd), fd));
      ratio = __dmul_rn(du, dd);
      acc = sm.u[i] <= ratio;
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, acc);
    const int first = ballot ? (__ffs(ballot) - 1) : 32;
    if (p.tr_ratio != nullptr && tid < 32 && i < N && lane <= first) {
      p.tr_ratio[trace_base + i] = ratio;
      p.tr_acc[trace_base + i] = (lane == first) ? 1 : 0;
    }
    if (!ballot) { i0 += 32; continue; }
#ifdef LQMC_PHASE_CLOCKS
    { const long long tk1 = clock64(); tk_scan += tk1 - tk0; tk0 = tk1; }
#endif
    const int is = i0 + first;
    gu = __shfl_sync(0xffffffffu, gu, first);
    gd = __shfl_sync(0xffffffffu, gd, first);
    const int hs = __shfl_sync(0xffffffffu, (int)h, first);
    const double fu = (hs > 0) ? p.f_p2 : p.f_m2;
    const double fd = (hs > 0) ? p.f_m2 : p.f_p2;
    // rebuild row `is` and column `is` of the current G: thread handles entries j = tid + 256 q of both spins.
    // The (strided) global loads of both spins are issued together, ahead of any use.
    double row2[2][QMAX], col2[2][QMAX];
#pragma unroll
    for (int spin = 0; spin < 2; ++spin) {
      const double* G = Gc + (size_t)spin * NP * NP;
#pragma unroll
      for (int q = 0; q < QMAX; ++q) {
        const int j = tid + L2_THREADS * q;
        if (j < NP) { row2[spin][q] = G[(size_t)is * NP + j]; col2[spin][q] = G[(size_t)j * NP + is]; }
      }
    }
#ifdef LQMC_PHASE_CLOCKS
    { double chk = row2[0][0] + col2[0][0] + row2[1][0] + col2[1][0]; if (chk == 1.2345e300) tk_b1 -= 1;   // wait for the loads
      const long long tk1 = clock64(); tk_b1 += tk1 - tk0; }
    long long tkb = clock64();
#endif
#pragma unroll
    for (int spin = 0; spin < 2; ++spin) {
      const double* U = sm.U + (size_t)spin * KD * NP;
      const double* W = sm.W + (size_t)spin * KD * NP;
      const double gs = spin ? gd : gu;
      double (&row)[QMAX] = row2[spin];
      double (&col)[QMAX] = col2[spin];
      for (int m = 0; m < nd; ++m) {
        const double ui = U[(size_t)m * NP + is], wi = W[(size_t)m * NP + is];
#pragma unroll
        for (int q = 0; q < QMAX; ++q) {
          const int j = tid + L2_THREADS * q;
          if (j < NP) {
            row[q] = rank1<EXACT>(row[q], ui, W[(size_t)m * NP + j]);
            col[q] = rank1<EXACT>(col[q], U[(size_t)m * NP + j], wi);
          }
        }
      }
#ifdef LQMC_PHASE_CLOCKS
      { double chk = row[0] + col[0]; if (chk == 1.2345e300) tk_b2 -= 1; const long long tk1 = clock64(); tk_b2 += tk1 - tkb; tkb = tk1; }
#endif
      double* Un = sm.U + ((size_t)spin * KD + nd) * NP;
      double* Wn = sm.W + ((size_t)spin * KD + nd) * NP;
      if (!PHYS) {
        const double gamma = spin ? fu : fd;            // exp(-arg)-1 for up, exp(+arg)-1 for down (lqmc.py:320-323)
        const double ci = __dadd_rn(__dmul_rn(-gamma, gs), gamma);
        const double den = __dadd_rn(1.0, ci);
        const double r = __drcp_rn(den);
        const bool ok = div_safe(den);
#pragma unroll
        for (int q = 0; q < QMAX; ++q) {
          const int j = tid + L2_THREADS * q;
          if (j < NP) {
            double c = __dmul_rn(-gamma, row[q]);
            if (j == is) c = __dadd_rn(c, gamma);
            const double e = EXACT ? div_shared_rcp(col[q], den, r, ok) : col[q] * r;
            Un[j] = e; Wn[j] = c;
            dnxt[spin * NP + j] = rank1<EXACT>(dcur[spin * NP + j], e, c);
          }
        }
      } else {
        const double delta = spin ? fd : fu;
        const double rr = __dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, gs), delta));
        const double fac = delta / rr;
#pragma unroll
        for (int q = 0; q < QMAX; ++q) {
          const int j = tid + L2_THREADS * q;
          if (j < NP) {
            const double e = ((j == is) ? (1.0 - col[q]) : -col[q]) * fac;
            const double c = row[q];
            Un[j] = e; Wn[j] = c;
            dnxt[spin * NP + j] = rank1<EXACT>(dcur[spin * NP + j], e, c);
          }
        }
      }
    }
#ifdef LQMC_PHASE_CLOCKS
    { const long long tk1 = clock64(); tk_b3 += tk1 - tkb; }
#endif
    ++n_accepted;
    ++nd;
    cur ^= 1;
    __syncthreads();
    if (tid == 0) sm.h[is] = (int8_t)(-hs);     // after the barrier: no warp is still scanning site `is`
#ifdef LQMC_PHASE_CLOCKS
    { const long long tk1 = clock64(); tk_build += tk1 - tk0; tk0 = tk1; }
#endif
    if (nd == KD) { l2_flush<EXACT>(Gc, NP, nd, sm, KD); nd = 0; }
#ifdef LQMC_PHASE_CLOCKS
    { const long long tk1 = clock64(); tk_flush += tk1 - tk0; tk0 = tk1; }
#endif
    i0 = is + 1;
  }
  if (nd > 0) l2_flush<EXACT>(Gc, NP, nd, sm, KD);
#ifdef LQMC_PHASE_CLOCKS
  { const long long tk1 = clock64(); tk_flush += tk1 - tk0;
    if (tid == 0) { double* ob = p.obs_sum + (size_t)blockIdx.x * 3 * N; ob[0] = (double)tk_scan; ob[1] = (double)tk_build; ob[2] = (double)tk_flush; ob[3] = (double)n_accepted; ob[4] = (double)tk_b1; ob[5] = (double)tk_b2; ob[6] = (double)tk_b3; } }
#endif
}

// ---- NP <= 256: twice the delay depth, c vectors parked in tensor memory ---------------------------------------------------
// The slice phase streams G through HBM once per flush (32 N^2 bytes for both spins) and, with every chain's G in flight
// (296 MiB at 16x16), that traffic - not the FP64 pipe - bounds it: per-flush time is the same at delay depth 3, 6 and 12,
// in EXACT and FMA arithmetic alike (profiles/r01_cfg4_summary.md).  The only lever is the delay depth, and shared memory
// (4 KD NP doubles for U and W) caps it at 12 with two CTAs per SM.  Blackwell's tensor memory is idle in an FP64 code:
// 256 KB per SM, 128 lanes x 512 32-bit columns, reached with tcgen05.ld / tcgen05.st.  In the 32x32b shape thread t of
// warp w addresses lane 32 (w % 4) + t % 32, i.e. TMEM is a per-thread scratchpad.  So:
//   * U (the e vectors) stays in shared memory, [2 spin][L2_KDT][NP]  (98 KB at NP = 256 for L2_KDT = 24);
//   * thread j parks the c history of ITS column, c_m[j], in TMEM (2 L2_KDT doubles = 96 columns; warps w and w+4 share a
//     lane quarter and use disjoint column windows; 256 columns per CTA, two CTAs per SM);
//   * row rebuild   G0[i][j] - sum_m e_m[i] c_m[j] : e_m[i] broadcast from shared memory, c_m[j] own (TMEM);
//     column rebuild G0[j][i] - sum_m e_m[j] c_m[i] : e_m[j] from shared memory, c_m[i] published by the warp that owns site i;
//   * the flush walks G with thread <-> column (a warp covers 32 consecutive columns, 256-byte row segments), c_m[j] in
//     registers for one spin at a time, e_m of eight rows as broadcast 128-bit shared loads.
// Same roundings in the same order as the generic path and the reference's one-flip-at-a-time loop (lqmc.py:328-331).
static_assert(L2_KDT % 8 == 0 && 8 * L2_KDT <= L2_TMEM_COLS, "two windows of 4 KDT columns; history read in chunks of 8 doubles");

__device__ __forceinline__ void tmem_st_f64(uint32_t taddr, double v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"((uint32_t)__double2loint(v)),
               "r"((uint32_t)__double2hiint(v))
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_f64x8(uint32_t taddr, double (&v)[8]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}
// two 8-double loads in flight, one wait (both spins' history chunks of one thread)
__device__ __forceinline__ void tmem_ld_f64x8_pair(uint32_t taddr0, uint32_t taddr1, double (&v0)[8], double (&v1)[8]) {
  uint32_t r[16], q[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr0)
               : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]), "=r"(q[9]),
                 "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
               : "r"(taddr1)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v0[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
    v1[i] = __hiloint2double((int)q[2 * i + 1], (int)q[2 * i]);
  }
}
// one warp allocates L2_TMEM_COLS columns for the CTA; every thread gets the base address
__device__ __forceinline__ uint32_t tmem_alloc_cta(uint32_t* slot) {
  if ((threadIdx.x >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(L2_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return *slot;
}
__device__ __forceinline__ void tmem_free_cta(uint32_t base) {
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(L2_TMEM_COLS) : "memory");
}

template <bool EXACT>
__device__ void l2_flush_tmem(double* __restrict__ Gc, int NP, int nd, const double* __restrict__ U3, uint32_t tm_my) {
  const int tid = threadIdx.x;
  if (tid < NP) {                                  // warp-uniform: NP is a multiple of 64
    for (int spin = 0; spin < 2; ++spin) {
      double cj[L2_KDT];
#pragma unroll
      for (int m0 = 0; m0 < L2_KDT; m0 += 8) {
        double v[8];
        tmem_ld_f64x8(tm_my + 2 * (spin * L2_KDT + m0), v);
#pragma unroll
        for (int q = 0; q < 8; ++q) cj[m0 + q] = v[q];
      }
      double* const col = Gc + (size_t)spin * NP * NP + tid;
      const double* const Us = U3 + (size_t)spin * L2_KDT * NP;
      double nxt[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) nxt[r] = col[(size_t)r * NP];
      for (int r0 = 0; r0 < NP; r0 += 8) {
        double g[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) g[r] = nxt[r];
        if (r0 + 8 < NP) {
#pragma unroll
          for (int r = 0; r < 8; ++r) nxt[r] = col[(size_t)(r0 + 8 + r) * NP];
        }
#pragma unroll
        for (int m = 0; m < L2_KDT; ++m) {
          if (m < nd) {
            const double* ur = Us + (size_t)m * NP + r0;
#pragma unroll
            for (int r = 0; r < 8; r += 2) {
              const double2 e = *reinterpret_cast<const double2*>(ur + r);
              g[r] = rank1<EXACT>(g[r], e.x, cj[m]);
              g[r + 1] = rank1<EXACT>(g[r + 1], e.y, cj[m]);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) col[(size_t)(r0 + r) * NP] = g[r];
      }
    }
  }
  __syncthreads();
}

// NP == 256: two columns per thread.  Every update needs e_m[row] from shared memory, and a broadcast load delivers one
// double per two LSU wavefronts whatever its width.  Warps w and w+4 address the same TMEM lane quarter, so thread t can read
// the c history of column t % 128 (window 0) AND of column t % 128 + 128 (window 1): it updates both columns with every e
// value it loads - half the shared-memory traffic per update - while the two thread halves split the rows.
//
// The flush is the FP64-pipe-bound half of the slice phase (exact arithmetic: one DMUL and one DADD per element and update,
// 2 x 2 x 256^2 x 24 / 64 lanes = 98 K clocks per flush), so its inner loop is written to keep that pipe fed:
//   * the history is zero-padded to a multiple of 8 updates (e = c = +0 leaves every element unchanged bit for bit in both
//     arithmetics: g - RN(0 * 0) = g, fma(-0, 0, g) = g), so the update loop has no per-update branch and the compiler
//     schedules the shared-memory loads of e ahead of the multiply-subtract chains of the previous update (the r01 version
//     had one basic block per update: every update started with an exposed LDS latency and ended with an exposed
//     DMUL -> DADD latency; alone on an SM it ran at 0.60 of the pipe floor, two co-resident CTAs at 0.74);
//   * the c history is read from tensor memory four updates at a time, the next four in flight while the current four are
//     applied (tcgen05.ld without a memory clobber; the wait carries the destination registers as operands so that no use
//     is scheduled above it);
//   * no register double buffer for G: a warp's 8-row chunk carries 24 x 32 FP64 instructions = 1.5 K pipe clocks, more
//     than the latency of its next loads, and the other warps of the sub-partition fill the gap.
// Same operations per element in the same order as the reference loop (lqmc.py:328-331): still bit-identical.
struct TmemQuad { uint32_t r[16]; };      // updates m .. m+3 of the thread's two columns: r[0..7] window 0, r[8..15] window 1
__device__ __forceinline__ void tmem_ld_quad_issue(uint32_t taddr0, uint32_t taddr1, TmemQuad& q) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(q.r[0]), "=r"(q.r[1]), "=r"(q.r[2]), "=r"(q.r[3]), "=r"(q.r[4]), "=r"(q.r[5]), "=r"(q.r[6]), "=r"(q.r[7])
               : "r"(taddr0));
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(q.r[8]), "=r"(q.r[9]), "=r"(q.r[10]), "=r"(q.r[11]), "=r"(q.r[12]), "=r"(q.r[13]), "=r"(q.r[14]), "=r"(q.r[15])
               : "r"(taddr1));
}
__device__ __forceinline__ void tmem_ld_quad_wait(TmemQuad& q) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(q.r[0]), "+r"(q.r[1]), "+r"(q.r[2]), "+r"(q.r[3]), "+r"(q.r[4]), "+r"(q.r[5]), "+r"(q.r[6]), "+r"(q.r[7]),
                 "+r"(q.r[8]), "+r"(q.r[9]), "+r"(q.r[10]), "+r"(q.r[11]), "+r"(q.r[12]), "+r"(q.r[13]), "+r"(q.r[14]), "+r"(q.r[15]));
}
#ifndef LQMC_FMA_FLUSH_DMMA
#define LQMC_FMA_FLUSH_DMMA 1      // fma arithmetic: flush on DMMA fragments (0: the scalar walk with one FMA per element)
#endif
#ifndef LQMC_FLUSH_ROWS
#define LQMC_FLUSH_ROWS 4          // rows per chunk of the two-column flush; the next chunk's G is prefetched into registers
#endif
template <bool EXACT, int R>
__device__ __forceinline__ void l2_flush_apply4(double (&g)[R][2], const TmemQuad& c, const double* __restrict__ ur, int NP) {
#if LQMC_FLUSH2_EPF
  // all e values of the quad requested up front (volatile: pinned), then the arithmetic
  double2 eb[4][R / 2];
  {
    const uint32_t ua = smem_u32(ur);
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int r = 0; r < R / 2; ++r)
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(eb[q][r].x), "=d"(eb[q][r].y) : "r"(ua + (uint32_t)(q * NP + 2 * r) * 8u) : "memory");
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double c0 = __hiloint2double((int)c.r[2 * q + 1], (int)c.r[2 * q]);
    const double c1 = __hiloint2double((int)c.r[8 + 2 * q + 1], (int)c.r[8 + 2 * q]);
#pragma unroll
    for (int r = 0; r < R; r += 2) {
      const double2 e = eb[q][r >> 1];
      g[r][0] = rank1<EXACT>(g[r][0], e.x, c0);
      g[r][1] = rank1<EXACT>(g[r][1], e.x, c1);
      g[r + 1][0] = rank1<EXACT>(g[r + 1][0], e.y, c0);
      g[r + 1][1] = rank1<EXACT>(g[r + 1][1], e.y, c1);
    }
  }
  return;
#endif
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double c0 = __hiloint2double((int)c.r[2 * q + 1], (int)c.r[2 * q]);
    const double c1 = __hiloint2double((int)c.r[8 + 2 * q + 1], (int)c.r[8 + 2 * q]);
#pragma unroll
    for (int r = 0; r < R; r += 2) {
      const double2 e = *reinterpret_cast<const double2*>(ur + (size_t)q * NP + r);
      g[r][0] = rank1<EXACT>(g[r][0], e.x, c0);
      g[r][1] = rank1<EXACT>(g[r][1], e.x, c1);
      g[r + 1][0] = rank1<EXACT>(g[r + 1][0], e.y, c0);
      g[r + 1][1] = rank1<EXACT>(g[r + 1][1], e.y, c1);
    }
  }
}

// __noinline__: the build loop around the call keeps its own register allocation.
// Column window: a flip needs column `is` of G0, a 256-line strided gather (2 K clocks of memory-system back-pressure per flip
// with every SM doing it, profiles/r02_cfg4_summary.md).  The flush has every element in registers anyway, so the threads that
// own the L2_COLWIN columns after the flush point (the sites the next 24 flips will come from) also store their values
// transposed into the idle second matrix buffer, Tc[spin][c][r] = G0[r][c]: 32 contiguous bytes per thread and chunk, + 12 % flush
// traffic, and the builder's column becomes one coalesced 2 KB read.
constexpr int L2_COLWIN = 64;
// One CTA per chain: the round-2 flush with compile-time chunk arithmetic (kept verbatim next to the cluster-capable version below)
template <bool EXACT>
__device__ __noinline__ void l2_flush_tmem2_single(double* __restrict__ Gc, int nd, double* __restrict__ U3, uint32_t tm_base,
                                            double* __restrict__ Tc, int wlo) {
  constexpr int NP = 256, R = LQMC_FLUSH_ROWS;
  U3 = as_shared(U3); Gc = as_global(Gc); Tc = as_global(Tc);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int jj = tid & 127, rbase = (tid >> 7) * (NP / 2);
  const uint32_t tm_lane = tm_base + ((uint32_t)(32 * (warp & 3)) << 16);
  const int nd8 = (nd + 7) & ~7;
  if (nd8 != nd) {                                   // zero-pad: thread t owns entry t of every e vector and the c history of column t
    const uint32_t tm_my = tm_lane + (uint32_t)((warp >> 2) * 4 * L2_KDT);
    for (int m = nd; m < nd8; ++m)
#pragma unroll
      for (int spin = 0; spin < 2; ++spin) {
        U3[((size_t)spin * L2_KDT + m) * NP + tid] = 0.0;
        tmem_st_f64(tm_my + 2 * (spin * L2_KDT + m), 0.0);
      }
    tmem_wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // the two spins are one walk of 2 x 128 / R chunks: the prefetch of the next chunk runs across the spin boundary
  constexpr int CH = NP / 2 / R;                     // chunks per spin
  auto chunk_ptr = [&](int ch) -> double* {
    const int spin = ch / CH, r0 = (ch % CH) * R;    // CH is a power of two
    return Gc + (size_t)spin * NP * NP + (size_t)(rbase + r0) * NP + jj;
  };
  double nxt[R][2];
  {
    const double* c0 = chunk_ptr(0);
#pragma unroll
    for (int r = 0; r < R; ++r) { nxt[r][0] = c0[(size_t)r * NP]; nxt[r][1] = c0[(size_t)r * NP + 128]; }
  }
  TmemQuad ca, cb;
  tmem_ld_quad_issue(tm_lane, tm_lane + 4 * L2_KDT, ca);
  for (int ch = 0; ch < 2 * CH; ++ch) {
    const int spin = ch / CH, r0 = (ch % CH) * R;
    double* const col = chunk_ptr(ch);
    const double* const Us = U3 + (size_t)spin * L2_KDT * NP + rbase + r0;
    const uint32_t tm0 = tm_lane + 2 * (spin * L2_KDT), tm1 = tm0 + 4 * L2_KDT;
    // history of the chunk after this one starts at update 0 of (possibly) the other spin
    const int spin_n = (ch + 1) / CH < 2 ? (ch + 1) / CH : 1;
    const uint32_t tn0 = tm_lane + 2 * (spin_n * L2_KDT), tn1 = tn0 + 4 * L2_KDT;
    double g[R][2];
#pragma unroll
    for (int r = 0; r < R; ++r) { g[r][0] = nxt[r][0]; g[r][1] = nxt[r][1]; }
    if (ch + 1 < 2 * CH) {
      const double* cn = chunk_ptr(ch + 1);
#pragma unroll
      for (int r = 0; r < R; ++r) { nxt[r][0] = cn[(size_t)r * NP]; nxt[r][1] = cn[(size_t)r * NP + 128]; }
    }
#if LQMC_FLUSH2_L2PF
    if (ch + 1 + LQMC_FLUSH2_L2PF < 2 * CH && (tid & 15) == 0) {     // one lane per 128-byte line
      const double* cp = chunk_ptr(ch + 1 + LQMC_FLUSH2_L2PF);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(cp + (size_t)r * NP));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(cp + (size_t)r * NP + 128));
      }
    }
#endif
    for (int m0 = 0; m0 < nd8; m0 += 8) {
      const bool last = m0 + 8 >= nd8;
      tmem_ld_quad_wait(ca);
      tmem_ld_quad_issue(tm0 + 2 * (m0 + 4), tm1 + 2 * (m0 + 4), cb);
      l2_flush_apply4<EXACT, R>(g, ca, Us + (size_t)m0 * NP, NP);
      tmem_ld_quad_wait(cb);
      tmem_ld_quad_issue(last ? tn0 : tm0 + 2 * (m0 + 8), last ? tn1 : tm1 + 2 * (m0 + 8), ca);
      l2_flush_apply4<EXACT, R>(g, cb, Us + (size_t)(m0 + 4) * NP, NP);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) { col[(size_t)r * NP] = g[r][0]; col[(size_t)r * NP + 128] = g[r][1]; }
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c = jj + 128 * cc;
      if ((unsigned)(c - wlo) < (unsigned)L2_COLWIN) {
        double* dst = Tc + (size_t)spin * NP * NP + (size_t)c * NP + rbase + r0;
#pragma unroll
        for (int r = 0; r < R; r += 2) *reinterpret_cast<double2*>(dst + r) = make_double2(g[r][cc], g[r + 1][cc]);
      }
    }
  }
  tmem_ld_quad_wait(ca);                               // drain the last (unused) history prefetch
  __syncthreads();
}

// SINGLE: one CTA per chain - the chunk range is a compile-time constant (the cluster variant's run-time range cost the
// one-CTA kernel 8 % of its slice phase when it was the only version)
template <bool EXACT, bool SINGLE>
__device__ __noinline__ void l2_flush_tmem2(double* __restrict__ Gc, int nd, double* __restrict__ U3, uint32_t tm_base,
                                            double* __restrict__ Tc, int wlo, int cs, int crank) {
  constexpr int NP = 256, R = LQMC_FLUSH_ROWS;
  U3 = as_shared(U3); Gc = as_global(Gc); Tc = as_global(Tc);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int jj = tid & 127, rbase = (tid >> 7) * (NP / 2);
  const uint32_t tm_lane = tm_base + ((uint32_t)(32 * (warp & 3)) << 16);
  const int nd8 = (nd + 7) & ~7;
  if (nd8 != nd) {                                   // zero-pad: thread t owns entry t of every e vector and the c history of column t
    const uint32_t tm_my = tm_lane + (uint32_t)((warp >> 2) * 4 * L2_KDT);
    for (int m = nd; m < nd8; ++m)
#pragma unroll
      for (int spin = 0; spin < 2; ++spin) {
        U3[((size_t)spin * L2_KDT + m) * NP + tid] = 0.0;
        tmem_st_f64(tm_my + 2 * (spin * L2_KDT + m), 0.0);
      }
    tmem_wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // the two spins are one walk of 2 x 128 / R chunks: the prefetch of the next chunk runs across the spin boundary
  constexpr int CHA = NP / 2 / R;                    // chunks per spin and thread half
  // this CTA of a cluster walks chunks [ch_lo, ch_lo + CH) (any cluster size)
  const int ch_lo = SINGLE ? 0 : crank * CHA / cs, CH = SINGLE ? CHA : (crank + 1) * CHA / cs - ch_lo;
  // the walk (spin, local chunk) is kept as running counters: no integer division in the loop
  auto chunk_ptr = [&](int spin, int lc) -> double* {
    return Gc + (size_t)spin * NP * NP + (size_t)(rbase + (ch_lo + lc) * R) * NP + jj;
  };
  double nxt[R][2];
  {
    const double* c0 = chunk_ptr(0, 0);
#pragma unroll
    for (int r = 0; r < R; ++r) { nxt[r][0] = c0[(size_t)r * NP]; nxt[r][1] = c0[(size_t)r * NP + 128]; }
  }
  TmemQuad ca, cb;
  tmem_ld_quad_issue(tm_lane, tm_lane + 4 * L2_KDT, ca);
  int spin = 0, lc = 0;
  for (int ch = 0; ch < 2 * CH; ++ch) {
    const int r0 = (ch_lo + lc) * R;
    double* const col = chunk_ptr(spin, lc);
    const double* const Us = U3 + (size_t)spin * L2_KDT * NP + rbase + r0;
    const uint32_t tm0 = tm_lane + 2 * (spin * L2_KDT), tm1 = tm0 + 4 * L2_KDT;
    // the chunk after this one (possibly the first of the other spin): its G0 is prefetched, its history starts at update 0
    int spin_n = spin, lc_n = lc + 1;
    if (lc_n == CH) { lc_n = 0; spin_n = spin + 1; }
    const int spin_h = spin_n < 2 ? spin_n : 1;
    const uint32_t tn0 = tm_lane + 2 * (spin_h * L2_KDT), tn1 = tn0 + 4 * L2_KDT;
    double g[R][2];
#pragma unroll
    for (int r = 0; r < R; ++r) { g[r][0] = nxt[r][0]; g[r][1] = nxt[r][1]; }
    if (ch + 1 < 2 * CH) {
      const double* cn = chunk_ptr(spin_n, lc_n);
#pragma unroll
      for (int r = 0; r < R; ++r) { nxt[r][0] = cn[(size_t)r * NP]; nxt[r][1] = cn[(size_t)r * NP + 128]; }
    }
    for (int m0 = 0; m0 < nd8; m0 += 8) {
      const bool last = m0 + 8 >= nd8;
      tmem_ld_quad_wait(ca);
      tmem_ld_quad_issue(tm0 + 2 * (m0 + 4), tm1 + 2 * (m0 + 4), cb);
      l2_flush_apply4<EXACT, R>(g, ca, Us + (size_t)m0 * NP, NP);
      tmem_ld_quad_wait(cb);
      tmem_ld_quad_issue(last ? tn0 : tm0 + 2 * (m0 + 8), last ? tn1 : tm1 + 2 * (m0 + 8), ca);
      l2_flush_apply4<EXACT, R>(g, cb, Us + (size_t)(m0 + 4) * NP, NP);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) { col[(size_t)r * NP] = g[r][0]; col[(size_t)r * NP + 128] = g[r][1]; }
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c = jj + 128 * cc;
      if ((unsigned)(c - wlo) < (unsigned)L2_COLWIN) {
        double* dst = Tc + (size_t)spin * NP * NP + (size_t)c * NP + rbase + r0;
#pragma unroll
        for (int r = 0; r < R; r += 2) *reinterpret_cast<double2*>(dst + r) = make_double2(g[r][cc], g[r + 1][cc]);
      }
    }
    spin = spin_n; lc = lc_n;
  }
  tmem_ld_quad_wait(ca);                               // drain the last (unused) history prefetch
  __syncthreads();
}

// columns [c_lo, c_lo + L2_COLWIN) of both spins, transposed into Tc (the window of a slice's first flips, and the window after a
// tensor-core flush, whose register layout does not hold 4 consecutive rows of a column); NP == 256
template <bool CL>
__device__ __forceinline__ void l2_colwin_init(const double* __restrict__ Gc, double* __restrict__ Tc, int c_lo = 0, int cs = 1, int crank = 0) {
  constexpr int NP = 256;
  const int tid = threadIdx.x + L2_THREADS * crank;         // pieces dealt out over all threads of the cluster
  // thread -> (spin, column c, 32-row block): 2 x 64 x 8 = 1024 pieces of 32 rows, four per thread; a warp reads 32 consecutive
  // columns of one row at a time (256-byte segments) and every thread stores 32 contiguous rows (256 bytes)
  for (int piece = tid; piece < 2 * L2_COLWIN * (NP / 32); piece += L2_THREADS * cs) {
    const int c = c_lo + piece % L2_COLWIN, rb = (piece / L2_COLWIN) % (NP / 32), spin = piece / (L2_COLWIN * (NP / 32));
    if (c >= NP) continue;
    const double* src = Gc + (size_t)spin * NP * NP + (size_t)(32 * rb) * NP + c;
    double* dst = Tc + (size_t)spin * NP * NP + (size_t)c * NP + 32 * rb;
#pragma unroll 4
    for (int r = 0; r < 32; r += 4) {
      const double a0 = l2_ld_peer<CL>(src + (size_t)r * NP), a1 = l2_ld_peer<CL>(src + (size_t)(r + 1) * NP),
                   a2 = l2_ld_peer<CL>(src + (size_t)(r + 2) * NP), a3 = l2_ld_peer<CL>(src + (size_t)(r + 3) * NP);
      *reinterpret_cast<double2*>(dst + r) = make_double2(a0, a1);
      *reinterpret_cast<double2*>(dst + r + 2) = make_double2(a2, a3);
    }
  }
}

// ---- fma arithmetic: the flush as an FP64 tensor-core GEMM (north_star item 1) -------------------------------------------------
// Where the reference's roundings are not claimed (LQMC_ARITH_FMA: one FMA per element update) the block update
// G0 <- G0 - U W^T is a k = 24 GEMM and runs on DMMA fragments:
//   * A = -e: fragments straight from the shared-memory history U3[k][row] (one LDS.64 per 4 DMMAs);
//   * B = c: the c history lives in tensor memory, one lane per G column.  tcgen05.ld in the 16x256b shape hands thread
//     (lr = t / 4, lk = t % 4) the two words 2 lk, 2 lk + 1 of lane lr (and of lane lr + 8) - the double c_{4s + lk} of column lr,
//     i.e. exactly the mma.m8n8k4 B fragment (k = lk, column = lr).  Tensor memory feeds the tensor-core operand with no
//     shuffle and no shared-memory round trip (layout checked on the device by tools/tmem_frag_test.cu);
//   * C = G0: accumulator fragments loaded from and stored to global memory (row lr, two adjacent columns per lane).
// Warp w owns the 32 columns whose histories sit in its lane quarter (window w / 4) and walks all 256 rows, two 8-row tiles at a
// time: 8 independent accumulators per k step.  2 x 256 x 256 x 24 FMAs = 49 K pipe clocks per flush, half the exact flush, and
// 1 / 8 of its shared-memory wavefronts - the scalar flush is LSU-bound in fma arithmetic (same time as the exact one).
// Measured (profiles/r02_cfg4_summary.md): cfg4 in fma arithmetic 269 -> 250 ms per sweep.  Prefetching the next pair of tiles
// makes the flush itself faster (11.3 K -> 9.6 K clocks per flip) but the sweep SLOWER (266 ms): back-to-back 16-clock DMMAs starve
// the co-resident chain's latency-bound build (5.1 K -> 8.8 K clocks per flip), so the walk is left un-prefetched.
__device__ __forceinline__ void tmem_ld_bfrag(uint32_t taddr, double& b0, double& b1) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  b0 = __hiloint2double((int)r1, (int)r0);
  b1 = __hiloint2double((int)r3, (int)r2);
}
template <bool CL>
__device__ __noinline__ void l2_flush_tmem2_dmma(double* __restrict__ Gc, int nd, double* __restrict__ U3, uint32_t tm_base,
                                                 double* __restrict__ Tc, int wlo, int cs_, int crank_) {
  const int cs = CL ? cs_ : 1, crank = CL ? crank_ : 0;
  U3 = as_shared(U3); Gc = as_global(Gc); Tc = as_global(Tc);
  constexpr int NP = 256;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lk = lane & 3;
  const int q = warp & 3, win = warp >> 2;
  const uint32_t tm_quarter = tm_base + ((uint32_t)(32 * q) << 16);
  const int nd4 = (nd + 3) & ~3;
  if (nd4 != nd) {                                   // zero-pad the last k step (thread t: entry t of the e vectors, history of column t)
    const uint32_t tm_my = tm_quarter + (uint32_t)(win * 4 * L2_KDT);
    for (int m = nd; m < nd4; ++m)
#pragma unroll
      for (int spin = 0; spin < 2; ++spin) {
        U3[((size_t)spin * L2_KDT + m) * NP + tid] = 0.0;
        tmem_st_f64(tm_my + 2 * (spin * L2_KDT + m), 0.0);
      }
    tmem_wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int c0 = 32 * q + 128 * win;                 // this warp's 32 columns
  const int nsteps = nd4 / 4;
  for (int spin = 0; spin < 2; ++spin) {
    double b[L2_KDT / 4][4];
#pragma unroll
    for (int s4 = 0; s4 < L2_KDT / 4; ++s4) {
      if (s4 < nsteps) {
        const uint32_t col = (uint32_t)(win * 4 * L2_KDT + 2 * (spin * L2_KDT + 4 * s4));
        tmem_ld_bfrag(tm_quarter + col, b[s4][0], b[s4][1]);                          // lanes 0-7 / 8-15 of the quarter: column tiles 0, 1
        tmem_ld_bfrag(tm_quarter + (16u << 16) + col, b[s4][2], b[s4][3]);            // lanes 16-23 / 24-31: column tiles 2, 3
      } else {
        b[s4][0] = b[s4][1] = b[s4][2] = b[s4][3] = 0.0;
      }
    }
    double* const Gs = Gc + (size_t)spin * NP * NP + c0 + 2 * lk;
    const double* const Us = U3 + (size_t)spin * L2_KDT * NP + (size_t)lk * NP + lr;
    const int mi_lo = 2 * (crank * (NP / 16) / cs), mi_hi = 2 * ((crank + 1) * (NP / 16) / cs);     // pairs of 8-row tiles of this CTA of a cluster
    for (int mi = mi_lo; mi < mi_hi; mi += 2) {
      double acc[2][4][2];
#pragma unroll
      for (int mm = 0; mm < 2; ++mm)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const double2 v = *reinterpret_cast<const double2*>(Gs + (size_t)(8 * (mi + mm) + lr) * NP + 8 * n);
          acc[mm][n][0] = v.x; acc[mm][n][1] = v.y;
        }
#pragma unroll
      for (int s4 = 0; s4 < L2_KDT / 4; ++s4) {
        if (s4 < nsteps) {
          double a[2];
#pragma unroll
          for (int mm = 0; mm < 2; ++mm) a[mm] = -Us[(size_t)(4 * s4) * NP + 8 * (mi + mm)];
#pragma unroll
          for (int mm = 0; mm < 2; ++mm)
#pragma unroll
            for (int n = 0; n < 4; ++n) dmma884(acc[mm][n], a[mm], b[s4][n]);
        }
      }
#pragma unroll
      for (int mm = 0; mm < 2; ++mm)
#pragma unroll
        for (int n = 0; n < 4; ++n)
          *reinterpret_cast<double2*>(Gs + (size_t)(8 * (mi + mm) + lr) * NP + 8 * n) = make_double2(acc[mm][n][0], acc[mm][n][1]);
    }
  }
  if (CL) l2_cluster_sync(cs); else __syncthreads();
  if (wlo < NP) {                                    // the builder's transposed column window, from the freshly written G0
    l2_colwin_init<CL>(Gc, Tc, wlo, cs, crank);
    if (cs == 1) __syncthreads();                    // (in a cluster the caller's barrier after the flush orders these stores)
  }
}

// The build of one accepted flip (row and column of the current G, Sherman-Morrison vectors) is a chain of dependent
// latencies - the FP64 work in it is small - and with the flush at the pipe floor it is what the co-resident CTA's flush has to
// hide.  Three things keep the chain short (r01: 3.5 K clocks per flip alone on an SM, 6.9 K next to a flushing CTA):
//   * G0 row / column of the next LQMC_L2_PF candidate sites are loaded into registers while the current flip is being built (the
//     next accepted site is the very next one with probability ~0.6, within three with ~0.93), so the strided 256-sector
//     column gather of a flip is no longer on its critical path; G0 only changes at a flush, after which the slots are reloaded;
//   * the c history of the flipped site, which every thread needs for the column rebuild, used to be published by the warp that
//     owns the site (tcgen05.ld, shared-memory store, one extra barrier per flip).  Now the threads of the next L2_PUB candidate
//     sites publish their own history while they apply it (they hold it in registers at that moment anyway) into a small ring
//     (slot = site & 7); the per-flip barrier orders it.  The owner-publish path remains for the rare longer jump;
//   * all divisions' common reciprocal depends only on the flipped site's diagonal and is computed by every thread while the
//     tensor-memory loads are in flight.
#ifndef LQMC_L2_PF
#define LQMC_L2_PF 0
#endif
#ifndef LQMC_L2_PFD
#define LQMC_L2_PFD 0
#endif
#ifndef LQMC_L2_SPEC
#define LQMC_L2_SPEC 0
#endif
// Two chains share an SM (2 CTAs) and one FP64 pipe.  A flush wants the whole pipe for ~100 K clocks, a build hardly any of it:
// the pair runs fastest in anti-phase (one chain flushes while the other builds its next 24 flips), but nothing makes two
// independent CTAs fall into that rhythm - whatever offset they start with persists, and when both flush at once both then
// build at once with the pipe idle.  A per-SM token (global memory, indexed by %smid) serialises the flushes of an SM: a CTA
// that finds the token taken waits for the other chain's flush to end, which locks the pair into anti-phase.  The wait is
// bounded (a CTA of another, aborted launch can never wedge the SM) and only the holder releases.
#ifndef LQMC_FLUSH_TOKEN
#define LQMC_FLUSH_TOKEN 0          // measured (profiles/r02_cfg4_summary.md): cfg4 sweep 268.6 ms without, 274.1 ms with the token
#endif
__device__ int g_l2_flush_token[1024];
__device__ __forceinline__ bool l2_flush_token_acquire() {
  bool got = false;
#if LQMC_FLUSH_TOKEN
  if (threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    const long long t0 = clock64();
    while (true) {
      if (atomicCAS(&g_l2_flush_token[smid & 1023], 0, 1) == 0) { got = true; break; }
      if (clock64() - t0 > 400000) break;
      __nanosleep(500);
    }
  }
  __syncthreads();
#endif
  return got;
}
__device__ __forceinline__ void l2_flush_token_release(bool got) {
#if LQMC_FLUSH_TOKEN
  if (threadIdx.x == 0 && got) {
    unsigned smid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    atomicExch(&g_l2_flush_token[smid & 1023], 0);
  }
#endif
}
// Compiled as a subroutine of its own (by-value arguments, like the GEMM): inlined into the sweep kernel the flip loop shared the
// 128-register budget with everything else that is live in the kernel and ptxas kept the prefetch slots in local memory - a
// store of a just-issued load, i.e. the prefetch turned into a blocking load.
struct L2SliceArgs {
  double* Gc; double* Tc;
  double* U; double* d; double* u; double* hist; double* ring; int8_t* h;      // shared memory (L2Smem)
  double* tr_ratio; uint8_t* tr_acc; double* obs;
  double f_p2, f_m2;
  long long trace_base;
  uint32_t tm_base;
  // 16-bit fields keep the struct at 128 bytes: one int more and nvcc stops scalarising the by-value argument - the pointers lose
  // their (deduced) shared / global state space and every history access becomes a generic LD.E / ST.E (build 5 K -> 10 K clocks)
  unsigned short NP, N, cs, crank;
};
static_assert(sizeof(L2SliceArgs) <= 128, "keep L2SliceArgs scalarisable");
struct L2SliceView {                       // the members of L2Smem / SweepParams the slice path touches, under their old names
  double* U; double* d; double* u; double* hist; double* ring; int8_t* h;
  double* tr_ratio; uint8_t* tr_acc; double* obs_sum; double f_p2, f_m2; int n_sites;
};
template <bool EXACT, bool PHYS, bool CL>
__device__ __noinline__ int l2_propose_slice_tmem_sub(const L2SliceArgs a) {
  constexpr int PF = LQMC_L2_PF;
  double* const Gc = as_global(a.Gc); double* const Tc = as_global(a.Tc);
  const int NP = a.NP;
  const long long trace_base = a.trace_base;
  const uint32_t tm_base = a.tm_base;
  L2SliceView sm{as_shared(a.U), as_shared(a.d), as_shared(a.u), as_shared(a.hist), as_shared(a.ring), as_shared(a.h), nullptr, nullptr, nullptr, 0.0, 0.0, 0};
  L2SliceView p{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, a.tr_ratio, a.tr_acc, a.obs, a.f_p2, a.f_m2, a.N};
  const int cs = CL ? a.cs : 1, crank = CL ? a.crank : 0;
  int n_accepted = 0;
  static_assert(PF >= 0 && PF <= 3, "prefetch depth");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.n_sites;
  double* const U3 = sm.U;                       // [2][L2_KDT][NP]: exactly the generic path's U / W region at KD = L2_KDT / 2
  double* const wis = sm.hist;                   // [2][L2_KDT] c history of the site being flipped (owner-publish path)
  const int j = tid;
  const bool act = j < NP;
  const size_t NN = (size_t)NP * NP;
  // this thread's TMEM window: lane quarter of its warp, columns [0, 4 KDT) for warps 0-3, [4 KDT, 8 KDT) for warps 4-7
  const uint32_t tm_my = tm_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 4 * L2_KDT);
  double* const myring = sm.ring + (size_t)(j & (L2_RING - 1)) * 2 * L2_KDT;
  for (int spin = 0; spin < 2; ++spin)
    for (int q = tid; q < NP; q += L2_THREADS) sm.d[spin * NP + q] = l2_ld_peer<CL>(Gc + (size_t)spin * NN + (size_t)q * NP + q);
  int wlo = NP;                                  // first column of the transposed window in Tc (NP: none)
  if (NP == 256) { l2_colwin_init<CL>(Gc, Tc, 0, cs, crank); wlo = 0; }
  if (CL) l2_cluster_sync(cs); else __syncthreads();
  int nd = 0, i0 = 0, cur = 0;
  int pf_base = -(1 << 20);                      // prow / pcol [k] = G0 row / column of site pf_base + k (valid until the next flush)
  int pub_base = -(1 << 20);                     // ring slots of sites pub_base + 1 .. pub_base + L2_PUB hold those sites' c history
  double prow[PF > 0 ? PF : 1][2], pcol[PF > 0 ? PF : 1][2];
  auto load_site = [&](int s, double (&r)[2], double (&c)[2]) {
    if (s > NP - 1) s = NP - 1;                  // past the last site: a harmless in-bounds load (never used)
    if (act) {
      const bool win = (unsigned)(s - wlo) < (unsigned)L2_COLWIN;      // CTA-uniform
      const double* cp = win ? Tc + (size_t)s * NP + j : Gc + (size_t)j * NP + s;
      c[0] = l2_ld_peer<CL>(cp);                     // L2 loads: in a cluster these lines are written by peer CTAs (flush, wrap)
      c[1] = l2_ld_peer<CL>(cp + NN);
      r[0] = l2_ld_peer<CL>(Gc + (size_t)s * NP + j);
      r[1] = l2_ld_peer<CL>(Gc + NN + (size_t)s * NP + j);
    }
  };
  // register-free look-ahead: rows / window columns of the sites up to LQMC_L2_PFD ahead are pulled into L2 (prefetch.global.L2
  // has no destination register), so that the loads of a later flip pay the L2, not the HBM latency
  int l2_hi = 0;                                  // sites below l2_hi have been prefetched since the last flush
  auto l2_prefetch_to = [&](int hi) {
#if LQMC_L2_PFD > 0
    if (hi > NP) hi = NP;
    for (int s = l2_hi; s < hi; ++s) {
      if (act) {
        const bool win = (unsigned)(s - wlo) < (unsigned)L2_COLWIN;
        const double* cp = win ? Tc + (size_t)s * NP + j : Gc + (size_t)j * NP + s;
        const double* rp = Gc + (size_t)s * NP + j;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rp));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + NN));
        if (win) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(cp));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(cp + NN));
        }
      }
    }
    if (hi > l2_hi) l2_hi = hi;
#else
    (void)hi;
#endif
  };
#ifdef LQMC_PHASE_CLOCKS
  long long tk_scan = 0, tk_build = 0, tk_flush = 0, tk0 = clock64();
  long long tk_b1 = 0, tk_b2 = 0, tk_b0 = 0, tk_b3 = 0, tk_b4 = 0; int n_slow = 0, n_miss = 0;
  const long long tk_begin = tk0;
  unsigned long long gt0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt0));
#endif
  while (i0 < N) {
    const double* dcur = sm.d + cur * 2 * NP;
    double* dnxt = sm.d + (cur ^ 1) * 2 * NP;
#if LQMC_L2_SPEC
    // Speculative loads: the very next site is the next accepted one with probability ~0.6.  Its G0 row / column are requested
    // BEFORE the scan decides - the registers live only inside this iteration (the loop-carried prefetch slots of PF > 0 are what
    // ptxas turned into blocking local-memory traffic) - so the scan hides part of their latency; a miss reloads.
    double srow[2], scol[2];
    const int spec = i0;
    load_site(spec, srow, scol);
#endif
    const int i = i0 + lane;
    bool acc = false;
    double gu = 0.0, gd = 0.0, ratio = 0.0;
    int8_t h = 1;
    if (i < N) {
      h = sm.h[i];
      gu = dcur[i];
      gd = dcur[NP + i];
      const double fu = (h > 0) ? p.f_p2 : p.f_m2;
      const double fd = (h > 0) ? p.f_m2 : p.f_p2;
      const double du = __dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, gu), fu));
      const double dd = __dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, gd), fd));
      ratio = __dmul_rn(du, dd);
      acc = sm.u[i] <= ratio;
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, acc);
    const int first = ballot ? (__ffs(ballot) - 1) : 32;
    if (p.tr_ratio != nullptr && tid < 32 && i < N && lane <= first) {
      p.tr_ratio[trace_base + i] = ratio;
      p.tr_acc[trace_base + i] = (lane == first) ? 1 : 0;
    }
    if (!ballot) { i0 += 32; continue; }
#ifdef LQMC_PHASE_CLOCKS
    { const long long tk1 = clock64(); tk_scan += tk1 - tk0; tk0 = tk1; }
#endif
    const int is = i0 + first;
    gu = __shfl_sync(0xffffffffu, gu, first);
    gd = __shfl_sync(0xffffffffu, gd, first);
    const int hs = __shfl_sync(0xffffffffu, (int)h, first);
    const double fu = (hs > 0) ? p.f_p2 : p.f_m2;
    const double fd = (hs > 0) ? p.f_m2 : p.f_p2;
    // G0 row / column of the flipped site, both spins: from the prefetch slots when the site is one of them.  The strided
    // column gather is issued first: it is the longest latency of the flip and is not needed before the column loop below.
    double row[2] = {0.0, 0.0}, col[2] = {0.0, 0.0};
#if LQMC_L2_EARLY_RCP
    // the flip's denominators and their reciprocals depend on the scan only: formed here, their latency (reciprocal + Newton steps,
    // ~30 dependent instructions for both spins) runs under the G0 loads instead of after the history loops
    double den_s[2], rcp_s[2];
#pragma unroll
    for (int spin = 0; spin < 2; ++spin) {
      const double gs = spin ? gd : gu;
      if (!PHYS) {
        const double gamma = spin ? fu : fd;
        den_s[spin] = __dadd_rn(1.0, __dadd_rn(__dmul_rn(-gamma, gs), gamma));
        rcp_s[spin] = __drcp_rn(den_s[spin]);
      } else {
        const double delta = spin ? fd : fu;
        den_s[spin] = __dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, gs), delta));
        rcp_s[spin] = delta / den_s[spin];
      }
    }
#endif
    const int dpf = is - pf_base;                  // CTA-uniform
    if (PF >= 1 && dpf >= 0 && dpf < PF) {
#pragma unroll
      for (int k = 0; k < PF; ++k)
        if (k == dpf) { row[0] = prow[k][0]; row[1] = prow[k][1]; col[0] = pcol[k][0]; col[1] = pcol[k][1]; }
    } else {
#if LQMC_L2_SPEC
      if (is == spec) { row[0] = srow[0]; row[1] = srow[1]; col[0] = scol[0]; col[1] = scol[1]; }
      else
#endif
      {
        load_site(is, row, col);
#ifdef LQMC_PHASE_CLOCKS
        ++n_miss;
#endif
      }
    }
    if (PF >= 1) {
      // slots for sites is + 1 .. is + PF: keep what is already there (shifted), load the rest
#pragma unroll
      for (int k = 0; k < PF; ++k) {
        bool have = false;
#pragma unroll
        for (int k2 = k + 1; k2 < PF; ++k2)
          if (dpf >= 0 && k2 == k + 1 + dpf) { prow[k][0] = prow[k2][0]; prow[k][1] = prow[k2][1]; pcol[k][0] = pcol[k2][0]; pcol[k][1] = pcol[k2][1]; have = true; }
        if (!have) load_site(is + 1 + k, prow[k], pcol[k]);
      }
      pf_base = is + 1;
    }
    if (l2_hi < is + 1) l2_hi = is + 1;
    l2_prefetch_to(is + 1 + LQMC_L2_PFD);
    // c history of the flipped site: in the ring if the site was one of the last flip's next candidates, else the warp that
    // owns the site publishes it (tcgen05.ld is warp-collective) at the price of one more barrier
    const int dpub = is - pub_base;
    const double* wsrc = sm.ring + (size_t)(is & (L2_RING - 1)) * 2 * L2_KDT;
    if (nd > 0 && !(dpub >= 1 && dpub <= L2_PUB)) {
      if (warp == (is >> 5)) {
        for (int m0 = 0; m0 < nd; m0 += 8) {
          double v0[8], v1[8];
          tmem_ld_f64x8_pair(tm_my + 2 * m0, tm_my + 2 * (L2_KDT + m0), v0, v1);
          if (lane == (is & 31)) {
#pragma unroll
            for (int q = 0; q < 8; ++q) { wis[m0 + q] = v0[q]; wis[L2_KDT + m0 + q] = v1[q]; }
          }
        }
      }
      __syncthreads();
      wsrc = wis;
#ifdef LQMC_PHASE_CLOCKS
      ++n_slow;
#endif
    }
    const int dj = j - is;
    const bool pubme = act && dj >= 1 && dj <= L2_PUB;
#ifdef LQMC_PHASE_CLOCKS
    { const long long tk1 = clock64(); tk_b1 += tk1 - tk0; }
#endif
    if (act) {
      // Row rebuild: own c history from tensor memory (four updates of both spins per load: r[0..7] up, r[8..15] down) times
      // e_m[is] (broadcast).  No per-update branch - the loads of a quad are issued together and the four multiplies are
      // independent; updates past nd are computed on stale operands and discarded by a select.
      for (int m0 = 0; m0 < nd; m0 += 4) {
        TmemQuad wq;
        tmem_ld_quad_issue(tm_my + 2 * m0, tm_my + 2 * (L2_KDT + m0), wq);
        double ui[2][4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int spin = 0; spin < 2; ++spin) ui[spin][q] = U3[((size_t)spin * L2_KDT + m0 + q) * NP + is];
#if LQMC_L2_ZPAD && LQMC_L2_MERGE_RC
        double uj[2][4], ws[2][4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int spin = 0; spin < 2; ++spin) {
            uj[spin][q] = U3[((size_t)spin * L2_KDT + m0 + q) * NP + j];
            ws[spin][q] = wsrc[spin * L2_KDT + m0 + q];
          }
#endif
        tmem_ld_quad_wait(wq);
        double wj[2][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          wj[0][q] = __hiloint2double((int)wq.r[2 * q + 1], (int)wq.r[2 * q]);
          wj[1][q] = __hiloint2double((int)wq.r[8 + 2 * q + 1], (int)wq.r[8 + 2 * q]);
        }
        if (pubme) {
#pragma unroll
          for (int q = 0; q < 4; ++q) { myring[m0 + q] = wj[0][q]; myring[L2_KDT + m0 + q] = wj[1][q]; }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#if LQMC_L2_ZPAD
          // slots nd .. of the quad hold e = c = +0 (written with slot nd & ~3, below): applying them changes no bit
#pragma unroll
          for (int spin = 0; spin < 2; ++spin) {
            row[spin] = rank1<EXACT>(row[spin], ui[spin][q], wj[spin][q]);
#if LQMC_L2_MERGE_RC
            col[spin] = rank1<EXACT>(col[spin], uj[spin][q], ws[spin][q]);     // four independent chains per thread
#endif
          }
#else
          const bool valid = m0 + q < nd;
#pragma unroll
          for (int spin = 0; spin < 2; ++spin) {
            const double r = rank1<EXACT>(row[spin], ui[spin][q], wj[spin][q]);
            row[spin] = valid ? r : row[spin];
          }
#endif
        }
      }
      // c vectors need the row only: park them in tensor memory (and the ring) while the column is still being rebuilt
      double cv[2];
#pragma unroll
      for (int spin = 0; spin < 2; ++spin) {
        if (!PHYS) {
          const double gamma = spin ? fu : fd;            // exp(-arg)-1 for up, exp(+arg)-1 for down (lqmc.py:320-323)
          cv[spin] = __dmul_rn(-gamma, row[spin]);
          if (j == is) cv[spin] = __dadd_rn(cv[spin], gamma);
        } else {
          cv[spin] = row[spin];
        }
        tmem_st_f64(tm_my + 2 * (spin * L2_KDT + nd), cv[spin]);
        if (pubme) myring[spin * L2_KDT + nd] = cv[spin];
#if LQMC_L2_ZPAD
        if ((nd & 3) == 0) {       // a new quad of history slots: its other three read as +0 until they are written
#pragma unroll
          for (int q = 1; q < 4; ++q) {
            tmem_st_f64(tm_my + 2 * (spin * L2_KDT + nd + q), 0.0);
            if (pubme) myring[spin * L2_KDT + nd + q] = 0.0;
          }
        }
#endif
      }
      // Column rebuild: e_m[j] (own entry) times the flipped site's c history (ring / owner-published, broadcast)
      for (int m0 = 0; m0 < ((LQMC_L2_ZPAD && LQMC_L2_MERGE_RC) ? 0 : nd); m0 += 4) {
        double uj[2][4], ws[2][4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int spin = 0; spin < 2; ++spin) {
            uj[spin][q] = U3[((size_t)spin * L2_KDT + m0 + q) * NP + j];
            ws[spin][q] = wsrc[spin * L2_KDT + m0 + q];
          }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#if LQMC_L2_ZPAD
#pragma unroll
          for (int spin = 0; spin < 2; ++spin) col[spin] = rank1<EXACT>(col[spin], uj[spin][q], ws[spin][q]);
#else
          const bool valid = m0 + q < nd;
#pragma unroll
          for (int spin = 0; spin < 2; ++spin) {
            const double r = rank1<EXACT>(col[spin], uj[spin][q], ws[spin][q]);
            col[spin] = valid ? r : col[spin];
          }
#endif
        }
      }
#ifdef LQMC_PHASE_CLOCKS
      { double chk = row[0] + col[0] + row[1] + col[1]; if (chk == 1.2345e300) tk_b2 -= 1; const long long tk1 = clock64(); tk_b2 += tk1 - tk0; }
#endif
#pragma unroll
      for (int spin = 0; spin < 2; ++spin) {
        const double gs = spin ? gd : gu;
        double e;
#if LQMC_L2_EARLY_RCP
        (void)gs;
        if (!PHYS) e = EXACT ? div_shared_rcp(col[spin], den_s[spin], rcp_s[spin], div_safe(den_s[spin])) : col[spin] * rcp_s[spin];
        else e = ((j == is) ? (1.0 - col[spin]) : -col[spin]) * rcp_s[spin];
#else
        if (!PHYS) {
          const double gamma = spin ? fu : fd;
          const double ci = __dadd_rn(__dmul_rn(-gamma, gs), gamma);
          const double den = __dadd_rn(1.0, ci);
          const double r = __drcp_rn(den);
          e = EXACT ? div_shared_rcp(col[spin], den, r, div_safe(den)) : col[spin] * r;
        } else {
          const double delta = spin ? fd : fu;
          const double rr = __dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, gs), delta));
          const double fac = delta / rr;
          e = ((j == is) ? (1.0 - col[spin]) : -col[spin]) * fac;
        }
#endif
        U3[((size_t)spin * L2_KDT + nd) * NP + j] = e;
#if LQMC_L2_ZPAD
        if ((nd & 3) == 0) {
#pragma unroll
          for (int q = 1; q < 4; ++q) U3[((size_t)spin * L2_KDT + nd + q) * NP + j] = 0.0;
        }
#endif
#ifdef LQMC_PHASE_CLOCKS
        if (spin == 1) { if (e + cv[1] == 1.2345e300) tk_b3 -= 1; const long long tk1 = clock64(); tk_b3 += tk1 - tk0; }
#endif
        dnxt[spin * NP + j] = rank1<EXACT>(dcur[spin * NP + j], e, cv[spin]);
      }
      tmem_wait_st();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");    // the two-column flush reads the partner warp's window
#ifdef LQMC_PHASE_CLOCKS
      { const long long tk1 = clock64(); tk_b4 += tk1 - tk0; }
#endif
    }
    pub_base = is;
    ++n_accepted;
    ++nd;
    cur ^= 1;
    __syncthreads();
    if (tid == 0) sm.h[is] = (int8_t)(-hs);     // after the barrier: no warp is still scanning site `is`
#ifdef LQMC_PHASE_CLOCKS
    { const long long tk1 = clock64(); tk_build += tk1 - tk0; tk0 = tk1; }
#endif
    if (nd == L2_KDT) {
      const bool tok = l2_flush_token_acquire();
      if (CL && cs > 1) l2_cluster_sync(cs);       // every CTA of the cluster has read what it needs of the un-flushed G0
      if (NP == 256) {
        wlo = is + 1;
        if (EXACT || !LQMC_FMA_FLUSH_DMMA) {
          if (!CL || cs == 1) l2_flush_tmem2_single<EXACT>(Gc, nd, U3, tm_base, Tc, wlo); else l2_flush_tmem2<EXACT, false>(Gc, nd, U3, tm_base, Tc, wlo, cs, crank);
        } else l2_flush_tmem2_dmma<CL>(Gc, nd, U3, tm_base, Tc, wlo, cs, crank);
      } else l2_flush_tmem<EXACT>(Gc, NP, nd, U3, tm_my);
      if (CL && cs > 1) l2_cluster_sync(cs);       // ... and now sees every CTA's share of the flushed one
      l2_flush_token_release(tok);
      nd = 0;
      if (PF > 0) {                               // G0 changed: reload the slots of the next candidates
#pragma unroll
        for (int k = 0; k < PF; ++k) load_site(is + 1 + k, prow[k], pcol[k]);
      }
    }
#ifdef LQMC_PHASE_CLOCKS
    { const long long tk1 = clock64(); tk_flush += tk1 - tk0; tk0 = tk1; }
#endif
    i0 = is + 1;
  }
  if (nd > 0) {
    const bool tok = l2_flush_token_acquire();
    if (CL && cs > 1) l2_cluster_sync(cs);
    if (NP == 256) {
      if (EXACT || !LQMC_FMA_FLUSH_DMMA) {
        if (!CL || cs == 1) l2_flush_tmem2_single<EXACT>(Gc, nd, U3, tm_base, Tc, NP); else l2_flush_tmem2<EXACT, false>(Gc, nd, U3, tm_base, Tc, NP, cs, crank);
      } else l2_flush_tmem2_dmma<CL>(Gc, nd, U3, tm_base, Tc, NP, cs, crank);
    } else l2_flush_tmem<EXACT>(Gc, NP, nd, U3, tm_my);
    l2_flush_token_release(tok);
  }
  if (CL && cs > 1) l2_cluster_sync(cs);           // the slice's last writes of G0 before the wrap GEMMs of the peers read it
#ifdef LQMC_PHASE_CLOCKS
  { const long long tk1 = clock64(); tk_flush += tk1 - tk0;
    if (tid == 0) { double* ob = p.obs_sum + (size_t)blockIdx.x * 3 * N; ob[0] = (double)tk_scan; ob[1] = (double)tk_build; ob[2] = (double)tk_flush; ob[3] = (double)n_accepted;
      ob[4] = (double)tk_b1; ob[5] = (double)tk_b2; ob[6] = (double)n_slow; ob[7] = (double)n_miss;
      unsigned smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
      unsigned long long gt1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt1));
      ob[8] = (double)(gt0 & 0xffffffffffffull); ob[9] = (double)(gt1 & 0xffffffffffffull); ob[10] = (double)smid; ob[11] = (double)(tk1 - tk_begin);
      ob[12] = (double)tk_b0; ob[13] = (double)tk_b3; ob[14] = (double)tk_b4; } }
#endif
  return n_accepted;
}

template <bool EXACT, bool PHYS, bool CL>
__device__ __forceinline__ void l2_propose_slice_tmem(double* __restrict__ Gc, double* __restrict__ Tc, int NP, L2Smem& sm, const SweepParams& p,
                                                      long long trace_base, int& n_accepted, uint32_t tm_base) {
  L2SliceArgs a;
  a.Gc = Gc; a.Tc = Tc; a.U = sm.U; a.d = sm.d; a.u = sm.u; a.hist = sm.hist; a.ring = sm.ring; a.h = sm.h;
  a.tr_ratio = p.tr_ratio; a.tr_acc = p.tr_acc; a.obs = p.obs_sum;
  a.f_p2 = p.f_p2; a.f_m2 = p.f_m2; a.trace_base = trace_base; a.tm_base = tm_base; a.NP = (unsigned short)NP; a.N = (unsigned short)p.n_sites;
  a.cs = (unsigned short)sm.cs; a.crank = (unsigned short)sm.crank;
  n_accepted += l2_propose_slice_tmem_sub<EXACT, PHYS, CL>(a);
}

// ---- 256 < NP <= 640: the tensor-memory slice path with several columns per thread ------------------------------------------
// At N = 576 (BASELINE configs[4]) the generic path's delay depth is 9 (U and W in 220 KB of shared memory) and the flush -
// 13 MB of G per 9 flips and chain - is 83 % of the slice phase and HBM-bound (5 TB/s with 148 chains, clock64 split in
// profiles/r01e_cfg4_summary.md).  Same remedy as for NP <= 256: only U stays in shared memory, thread t parks the c history of
// ITS columns t, t + 256, ... (CPT of them) in tensor memory, which buys delay depth KDX = 16 (three columns) / 24 (two columns).  One
// CTA per SM at these sizes, so the whole 512-column TMEM is this CTA's: window of warp w = [4 CPT KDX (w / 4), ...), inside it
// column-set q, spin s, update m at 32-bit column 2 ((2 q + s) KDX + m).  Same roundings in the same order as the generic path.
constexpr int L2_TMEMX_COLS = 512;
__device__ __forceinline__ uint32_t tmem_alloc_cta_x(uint32_t* slot) {
  if ((threadIdx.x >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(L2_TMEMX_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return *slot;
}
__device__ __forceinline__ void tmem_free_cta_x(uint32_t base) {
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(L2_TMEMX_COLS) : "memory");
}

// All CPT columns of a thread in one pass per spin: with one CTA (8 warps) per SM the one-column walk cannot fill the FP64
// pipe (clock64 split at N = 576: 53.8 K clocks per flip alone on an SM against a floor of 25.6 K); CPT columns per e value
// loaded give CPT x 8 independent multiply-subtract chains per thread and 1 / CPT of the shared-memory loads per update.
// The c histories of all columns (CPT x KDX doubles) live in registers for the pass - these kernels run one CTA per SM and
// may use 255 registers.
// (r2) Branch-free like the two-column flush: updates past nd are applied with e = 0 (rows nd .. KDX-1 of U3 are zeroed before a partial
// flush) and c = 0 (selected at load time), so the fully unrolled update loop has no per-update branch and the compiler schedules
// the shared-memory loads of e ahead of the multiply-subtract chains.  The threads that own the L2_COLWIN columns after the flush
// point also store them transposed into Tc (8 consecutive rows = 64 contiguous bytes per thread and chunk): the builder's column of
// G0 becomes a coalesced read instead of an NP-line gather.
template <bool EXACT, int CPT, int KDX>
__device__ void l2_flush_tmemx(double* __restrict__ Gc, int NP, int nd, double* __restrict__ U3, uint32_t tm_my, double* __restrict__ Tc, int wlo) {
  const int tid = threadIdx.x;
  // NP is a multiple of 64, so whether a thread has a column in window q is warp-uniform; windows 0 .. CPT-2 are always full and
  // only the last one is partial (N = 576: warps 0 and 1 own a third column, warps 2 - 7 do not).  The warps without it run the
  // CPT - 1 instantiation of the pass instead of multiplying zeros (a third of their FP64 instructions at N = 576).
  const bool last_col = tid + L2_THREADS * (CPT - 1) < NP;
  if (nd < KDX) {
    for (int spin = 0; spin < 2; ++spin)
      for (int m = nd; m < KDX; ++m)
        for (int r = tid; r < NP; r += L2_THREADS) U3[((size_t)spin * KDX + m) * NP + r] = 0.0;
    __syncthreads();
  }
  auto pass = [&](auto ca_tag) {
    constexpr int CA = decltype(ca_tag)::value;      // columns of this thread
    bool jw[CA];
#pragma unroll
    for (int q = 0; q < CA; ++q) jw[q] = (unsigned)(tid + L2_THREADS * q - wlo) < (unsigned)L2_COLWIN;
    for (int spin = 0; spin < 2; ++spin) {
      double cj[CA][KDX];
#pragma unroll
      for (int q = 0; q < CA; ++q)
#pragma unroll
        for (int m0 = 0; m0 < KDX; m0 += 8) {
          double v[8];
          tmem_ld_f64x8(tm_my + 2 * ((2 * q + spin) * KDX + m0), v);
#pragma unroll
          for (int t = 0; t < 8; ++t) cj[q][m0 + t] = (m0 + t < nd) ? v[t] : 0.0;
        }
      double* const col = Gc + (size_t)spin * NP * NP + tid;
      double* const colT = Tc + (size_t)spin * NP * NP + (size_t)tid * NP;
      const double* const Us = U3 + (size_t)spin * KDX * NP;
      double nxt[CA][8];
#pragma unroll
      for (int q = 0; q < CA; ++q)
#pragma unroll
        for (int r = 0; r < 8; ++r) nxt[q][r] = col[(size_t)r * NP + L2_THREADS * q];
      for (int r0 = 0; r0 < NP; r0 += 8) {
        double g[CA][8];
#pragma unroll
        for (int q = 0; q < CA; ++q)
#pragma unroll
          for (int r = 0; r < 8; ++r) g[q][r] = nxt[q][r];
        if (r0 + 8 < NP) {
          // the next chunk's G0 is requested before this chunk's arithmetic (volatile: pinned here)
#pragma unroll
          for (int q = 0; q < CA; ++q)
#pragma unroll
            for (int r = 0; r < 8; ++r)
              asm volatile("ld.global.f64 %0, [%1];" : "=d"(nxt[q][r]) : "l"(col + (size_t)(r0 + 8 + r) * NP + L2_THREADS * q) : "memory");
        }
#if LQMC_FLUSHX_L2PF
        // ... and the chunk after that is pulled into L2 (no destination register: more HBM requests in flight than the 24
        // register loads per thread allow); one lane per 128-byte line
        if (r0 + 8 * (1 + LQMC_FLUSHX_L2PF) < NP && (tid & 15) == 0) {
#pragma unroll
          for (int q = 0; q < CA; ++q)
#pragma unroll
            for (int r = 0; r < 8; ++r)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(col + (size_t)(r0 + 8 * (1 + LQMC_FLUSHX_L2PF) + r) * NP + L2_THREADS * q));
        }
#endif
#if LQMC_FLUSHX_EPF
        // e values of update m + 1 requested (volatile: pinned in this order) before the arithmetic of update m: with 8 warps per SM a
        // broadcast LDS.128 issued right before its use is not covered by the 12 FP64 instructions of the previous one
        // (clock64 at N = 576: flush 57.2 K -> 50.3 K clocks per flip)
        double2 ebuf[2][4];
        {
          const uint32_t ua = smem_u32(Us + r0);
#pragma unroll
          for (int r = 0; r < 4; ++r)
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(ebuf[0][r].x), "=d"(ebuf[0][r].y) : "r"(ua + 16 * r) : "memory");
        }
#pragma unroll
        for (int m = 0; m < KDX; ++m) {
          if (m + 1 < KDX) {
            const uint32_t ua = smem_u32(Us + (size_t)(m + 1) * NP + r0);
#pragma unroll
            for (int r = 0; r < 4; ++r)
              asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(ebuf[(m + 1) & 1][r].x), "=d"(ebuf[(m + 1) & 1][r].y) : "r"(ua + 16 * r) : "memory");
          }
#pragma unroll
          for (int r = 0; r < 8; r += 2) {
            const double2 e = ebuf[m & 1][r >> 1];
#pragma unroll
            for (int q = 0; q < CA; ++q) {
              g[q][r] = rank1<EXACT>(g[q][r], e.x, cj[q][m]);
              g[q][r + 1] = rank1<EXACT>(g[q][r + 1], e.y, cj[q][m]);
            }
          }
        }
#else
#pragma unroll
        for (int m = 0; m < KDX; ++m) {
          const double* ur = Us + (size_t)m * NP + r0;
#pragma unroll
          for (int r = 0; r < 8; r += 2) {
            const double2 e = *reinterpret_cast<const double2*>(ur + r);
#pragma unroll
            for (int q = 0; q < CA; ++q) {
              g[q][r] = rank1<EXACT>(g[q][r], e.x, cj[q][m]);
              g[q][r + 1] = rank1<EXACT>(g[q][r + 1], e.y, cj[q][m]);
            }
          }
        }
#endif
#pragma unroll
        for (int q = 0; q < CA; ++q) {
#pragma unroll
          for (int r = 0; r < 8; ++r) col[(size_t)(r0 + r) * NP + L2_THREADS * q] = g[q][r];
          if (jw[q]) {
            double* dst = colT + (size_t)(L2_THREADS * q) * NP + r0;
#pragma unroll
            for (int r = 0; r < 8; r += 2) *reinterpret_cast<double2*>(dst + r) = make_double2(g[q][r], g[q][r + 1]);
          }
        }
      }
    }
  };
  if (last_col) pass(std::integral_constant<int, CPT>{});
  else pass(std::integral_constant<int, CPT - 1>{});
  __syncthreads();
}

// columns [c_lo, c_lo + L2_COLWIN) of both spins, transposed into Tc - any NP (the window of a slice's first flips)
__device__ __forceinline__ void l2_colwin_init_np(const double* __restrict__ Gc, double* __restrict__ Tc, int NP, int c_lo) {
  const int tid = threadIdx.x, nb = NP / 32;
  for (int piece = tid; piece < 2 * L2_COLWIN * nb; piece += L2_THREADS) {
    const int c = c_lo + piece % L2_COLWIN, rb = (piece / L2_COLWIN) % nb, spin = piece / (L2_COLWIN * nb);
    if (c >= NP) continue;
    const double* src = Gc + (size_t)spin * NP * NP + (size_t)(32 * rb) * NP + c;
    double* dst = Tc + (size_t)spin * NP * NP + (size_t)c * NP + 32 * rb;
#pragma unroll 4
    for (int r = 0; r < 32; r += 4) {
      const double a0 = src[(size_t)r * NP], a1 = src[(size_t)(r + 1) * NP], a2 = src[(size_t)(r + 2) * NP], a3 = src[(size_t)(r + 3) * NP];
      *reinterpret_cast<double2*>(dst + r) = make_double2(a0, a1);
      *reinterpret_cast<double2*>(dst + r + 2) = make_double2(a2, a3);
    }
  }
}

template <bool EXACT, bool PHYS, int CPT, int KDX>
__device__ void l2_propose_slice_tmemx(double* __restrict__ Gc, double* __restrict__ Tc, int NP, L2Smem& sm, const SweepParams& p, long long trace_base,
                                       int& n_accepted, uint32_t tm_base) {
  static_assert(KDX % 8 == 0 && 8 * CPT * KDX <= L2_TMEMX_COLS && 2 * KDX <= 62, "TMEM windows / history slots");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.n_sites;
  double* const U3 = sm.U;                       // [2][KDX][NP]
  double* const wis = sm.hist;                   // [2][KDX] c history of the site being flipped
  const uint32_t tm_my = tm_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 4 * CPT * KDX);
  for (int spin = 0; spin < 2; ++spin)
    for (int q = tid; q < NP; q += L2_THREADS) sm.d[spin * NP + q] = Gc[(size_t)spin * NP * NP + (size_t)q * NP + q];
  l2_colwin_init_np(Gc, Tc, NP, 0);              // transposed copy of the first L2_COLWIN columns (see l2_flush_tmemx)
  int wlo = 0;
  __syncthreads();
  int nd = 0, i0 = 0, cur = 0;
#ifdef LQMC_PHASE_CLOCKS
  long long tk_scan = 0, tk_build = 0, tk_flush = 0, tk0 = clock64();
  const long long tk_begin = tk0;
  unsigned long long gt0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt0));
#endif
  while (i0 < N) {
    const double* dcur = sm.d + cur * 2 * NP;
    double* dnxt = sm.d + (cur ^ 1) * 2 * NP;
    const int i = i0 + lane;
    bool acc = false;
    double gu = 0.0, gd = 0.0, ratio = 0.0;
    int8_t h = 1;
    if (i < N) {
      h = sm.h[i];
      gu = dcur[i];
      gd = dcur[NP + i];
      const double fu = (h > 0) ? p.f_p2 : p.f_m2;
      const double fd = (h > 0) ? p.f_m2 : p.f_p2;
      const double du = __dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, gu), fu));
      const double dd = __dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, gd), fd));
      ratio = __dmul_rn(du, dd);
      acc = sm.u[i] <= ratio;
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, acc);
    const int first = ballot ? (__ffs(ballot) - 1) : 32;
    if (p.tr_ratio != nullptr && tid < 32 && i < N && lane <= first) {
      p.tr_ratio[trace_base + i] = ratio;
      p.tr_acc[trace_base + i] = (lane == first) ? 1 : 0;
    }
    if (!ballot) { i0 += 32; continue; }
#ifdef LQMC_PHASE_CLOCKS
    { const long long tk1 = clock64(); tk_scan += tk1 - tk0; tk0 = tk1; }
#endif
    const int is = i0 + first;
    gu = __shfl_sync(0xffffffffu, gu, first);
    gd = __shfl_sync(0xffffffffu, gd, first);
    const int hs = __shfl_sync(0xffffffffu, (int)h, first);
    const double fu = (hs > 0) ? p.f_p2 : p.f_m2;
    const double fd = (hs > 0) ? p.f_m2 : p.f_p2;
    // G0 row / column of the flipped site, both spins, all of this thread's columns, issued together
    double row[CPT][2], col[CPT][2];
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
      const int j = tid + L2_THREADS * q;
#pragma unroll
      for (int spin = 0; spin < 2; ++spin) {
        const double* G = Gc + (size_t)spin * NP * NP;
        row[q][spin] = (j < NP) ? G[(size_t)is * NP + j] : 0.0;
        col[q][spin] = (j < NP) ? (((unsigned)(is - wlo) < (unsigned)L2_COLWIN) ? Tc[(size_t)spin * NP * NP + (size_t)is * NP + j]
                                                                                 : G[(size_t)j * NP + is])
                                : 0.0;
      }
    }
    // the warp that owns site `is` publishes that site's c history
    {
      const int t_is = is & (L2_THREADS - 1), q_is = is / L2_THREADS;
      if (warp == (t_is >> 5)) {
        for (int m0 = 0; m0 < nd; m0 += 8) {
          double v0[8], v1[8];
          tmem_ld_f64x8_pair(tm_my + 2 * ((2 * q_is) * KDX + m0), tm_my + 2 * ((2 * q_is + 1) * KDX + m0), v0, v1);
          if (lane == (t_is & 31)) {
#pragma unroll
            for (int t = 0; t < 8; ++t) { wis[m0 + t] = v0[t]; wis[KDX + m0 + t] = v1[t]; }
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
      const int j = tid + L2_THREADS * q;
      if (j < NP) {
        for (int m0 = 0; m0 < nd; m0 += 8) {
          double wj[2][8];
          tmem_ld_f64x8_pair(tm_my + 2 * ((2 * q) * KDX + m0), tm_my + 2 * ((2 * q + 1) * KDX + m0), wj[0], wj[1]);
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int m = m0 + t;
            if (m < nd) {
#pragma unroll
              for (int spin = 0; spin < 2; ++spin) {
                const double* Um = U3 + ((size_t)spin * KDX + m) * NP;
                row[q][spin] = rank1<EXACT>(row[q][spin], Um[is], wj[spin][t]);
                col[q][spin] = rank1<EXACT>(col[q][spin], Um[j], wis[spin * KDX + m]);
              }
            }
          }
        }
#pragma unroll
        for (int spin = 0; spin < 2; ++spin) {
          const double gs = spin ? gd : gu;
          double e, c;
          if (!PHYS) {
            const double gamma = spin ? fu : fd;            // exp(-arg)-1 for up, exp(+arg)-1 for down (lqmc.py:320-323)
            const double ci = __dadd_rn(__dmul_rn(-gamma, gs), gamma);
            const double den = __dadd_rn(1.0, ci);
            const double r = __drcp_rn(den);
            c = __dmul_rn(-gamma, row[q][spin]);
            if (j == is) c = __dadd_rn(c, gamma);
            e = EXACT ? div_shared_rcp(col[q][spin], den, r, div_safe(den)) : col[q][spin] * r;
          } else {
            const double delta = spin ? fd : fu;
            const double rr = __dadd_rn(1.0, __dmul_rn(__dsub_rn(1.0, gs), delta));
            const double fac = delta / rr;
            e = ((j == is) ? (1.0 - col[q][spin]) : -col[q][spin]) * fac;
            c = row[q][spin];
          }
          U3[((size_t)spin * KDX + nd) * NP + j] = e;
          tmem_st_f64(tm_my + 2 * ((2 * q + spin) * KDX + nd), c);
          dnxt[spin * NP + j] = rank1<EXACT>(dcur[spin * NP + j], e, c);
        }
      }
    }
    tmem_wait_st();
    ++n_accepted;
    ++nd;
    cur ^= 1;
    __syncthreads();
    if (tid == 0) sm.h[is] = (int8_t)(-hs);     // after the barrier: no warp is still scanning site `is`
#ifdef LQMC_PHASE_CLOCKS
    { const long long tk1 = clock64(); tk_build += tk1 - tk0; tk0 = tk1; }
#endif
    if (nd == KDX) { wlo = is + 1; l2_flush_tmemx<EXACT, CPT, KDX>(Gc, NP, nd, U3, tm_my, Tc, wlo); nd = 0; }
#ifdef LQMC_PHASE_CLOCKS
    { const long long tk1 = clock64(); tk_flush += tk1 - tk0; tk0 = tk1; }
#endif
    i0 = is + 1;
  }
  if (nd > 0) l2_flush_tmemx<EXACT, CPT, KDX>(Gc, NP, nd, U3, tm_my, Tc, NP);
#ifdef LQMC_PHASE_CLOCKS
  { const long long tk1 = clock64(); tk_flush += tk1 - tk0;
    if (tid == 0) { double* ob = p.obs_sum + (size_t)blockIdx.x * 3 * N; ob[0] = (double)tk_scan; ob[1] = (double)tk_build; ob[2] = (double)tk_flush; ob[3] = (double)n_accepted;
      for (int q = 4; q < 15; ++q) ob[q] = 0.0;
      unsigned smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
      unsigned long long gt1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt1));
      ob[8] = (double)(gt0 & 0xffffffffffffull); ob[9] = (double)(gt1 & 0xffffffffffffull); ob[10] = (double)smid; ob[11] = (double)(tk1 - tk_begin); } }
#endif
}

// ---- Gauss-Jordan inverse in memory (both spins in lockstep, in place), partial pivoting, delayed updates ----------
// np.linalg.inv of the sweep-start matrix (lqmc.py:306-307).  A Gauss-Jordan step is a rank-1 update of the whole
// matrix, so it is delayed exactly like the flips: the pivot column and pivot row of the *current* matrix are rebuilt
// from M0 and the pending (U, W) pairs, the step is appended, and M0 <- M0 - U W^T is applied once per KD pivots
// (KD x less memory traffic than one full pass per pivot).  The pivot row and pivot column are written to M0 in
// final form and detached from the pending updates (their U / W entries are zeroed), so they carry no
// cancellation error.  Same pivot rule as LAPACK's getrf (first entry of largest magnitude); rows are swapped
// physically, columns are un-permuted at the end.
template <int NS = 2>
__device__ void l2_gj_inverse(double* __restrict__ Gc, int NP, int KD, L2Smem& sm, int* piv_global) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int HALF = L2_THREADS / NS, WPS = HALF / 32;     // threads / warps per spin in the pivot search
  double* colk = sm.d;                                       // [2][NP] (the diagonal cache is idle here)
  int nd = 0;
  for (int k = 0; k < NP; ++k) {
    // 1. column k of the current matrices
    for (int spin = 0; spin < NS; ++spin) {
      const double* U = sm.U + (size_t)spin * KD * NP;
      const double* W = sm.W + (size_t)spin * KD * NP;
      for (int r = tid; r < NP; r += L2_THREADS) {
        double v = Gc[(size_t)spin * NP * NP + (size_t)r * NP + k];
        for (int m = 0; m < nd; ++m) v = fma(-U[(size_t)m * NP + r], W[(size_t)m * NP + k], v);
        colk[spin * NP + r] = v;
      }
    }
    __syncthreads();
    // 2. pivot search, one half of the CTA per spin
    {
      const int spin = tid / HALF, t = tid % HALF;
      double pv = 0.0, av = -1.0;
      int idx = NP;
      for (int r = k + t; r < NP; r += HALF) {
        const double v = colk[spin * NP + r];
        const double a = fabs(v);
        if (a > av) { av = a; pv = v; idx = r; }
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const double oa = __shfl_down_sync(0xffffffffu, av, off);
        const double op = __shfl_down_sync(0xffffffffu, pv, off);
        const int oi = __shfl_down_sync(0xffffffffu, idx, off);
        if (oa > av || (oa == av && oi < idx)) { av = oa; pv = op; idx = oi; }
      }
      if (lane == 0) { sm.red_v[warp] = pv; sm.red_i[warp] = idx; }
    }
    __syncthreads();
    int pidx[NS];
    double ppv[NS];
#pragma unroll
    for (int spin = 0; spin < NS; ++spin) {
      double pv = sm.red_v[spin * WPS];
      int idx = sm.red_i[spin * WPS];
#pragma unroll
      for (int w = 1; w < WPS; ++w) {
        const double ov = sm.red_v[spin * WPS + w];
        const int oi = sm.red_i[spin * WPS + w];
        if (oi < NP && (idx >= NP || fabs(ov) > fabs(pv) || (fabs(ov) == fabs(pv) && oi < idx))) { pv = ov; idx = oi; }
      }
      if (idx >= NP) { idx = k; pv = colk[spin * NP + k]; }
      pidx[spin] = idx; ppv[spin] = pv;
    }
    if (tid < NS) piv_global[tid * NP + k] = pidx[tid];
    // 3. current pivot row (row p before the swap), scaled with the 1-injection; physical swap in M0
    for (int spin = 0; spin < NS; ++spin)
    for (int c = tid; c < NP; c += L2_THREADS) {
      const int pi = pidx[spin];
      const double* U = sm.U + (size_t)spin * KD * NP;
      const double* W = sm.W + (size_t)spin * KD * NP;
      double* M = Gc + (size_t)spin * NP * NP;
      const double oldk = M[(size_t)k * NP + c];
      double v = (pi == k) ? oldk : M[(size_t)pi * NP + c];
      for (int m = 0; m < nd; ++m) v = fma(-U[(size_t)m * NP + pi], W[(size_t)m * NP + c], v);
      const double rk = ((c == k) ? 1.0 : v) / ppv[spin];
      M[(size_t)k * NP + c] = rk;                                   // row k in final form
      if (pi != k) M[(size_t)pi * NP + c] = (c == k) ? 0.0 : oldk;      // old row k moves to row p (its column-k entry is eliminated)
      sm.W[((size_t)spin * KD + nd) * NP + c] = rk;
    }
    __syncthreads();
    // 4. multipliers, detach row k / column k from the pending updates, eliminate column k in M0
    for (int spin = 0; spin < NS; ++spin)
    for (int r = tid; r < NP; r += L2_THREADS) {
      const int pi = pidx[spin];
      double f;
      if (r == k) f = 0.0;
      else if (r == pi) f = colk[spin * NP + k];
      else f = colk[spin * NP + r];
      sm.U[((size_t)spin * KD + nd) * NP + r] = f;
      if (r != k && r != pi) Gc[(size_t)spin * NP * NP + (size_t)r * NP + k] = 0.0;
    }
    for (int spin = 0; spin < NS; ++spin)
    for (int m = tid; m < nd; m += L2_THREADS) {
      const int pi = pidx[spin];
      double* U = sm.U + ((size_t)spin * KD + m) * NP;
      double* W = sm.W + ((size_t)spin * KD + m) * NP;
      if (pi != k) U[pi] = U[k];
      U[k] = 0.0;
      W[k] = 0.0;
    }
    ++nd;
    __syncthreads();
    if (nd == KD) { l2_flush<false, NS>(Gc, NP, nd, sm, KD); nd = 0; }
  }
  if (nd > 0) l2_flush<false, NS>(Gc, NP, nd, sm, KD);
  // undo the row interchanges on the columns (last pivot first).  The swap sequence is composed once into a gather index
  // in shared memory (idx[c] = where column c of the result comes from); then every row is staged through a per-warp
  // shared buffer and written back permuted, coalesced both ways.  (A per-thread chain of NP dependent global swaps cost
  // 15 % of the inverse kernel.)
  {
    int* idx = reinterpret_cast<int*>(sm.U);                          // [NS][NP]; U / W are dead after the last flush
    double* rowbuf = sm.U + (size_t)NS * NP / 2 + (size_t)warp * NP;    // 8 x [NP] doubles behind it
    for (int q = tid; q < NS * NP; q += L2_THREADS) idx[q] = q % NP;
    __syncthreads();
    if (tid < NS) {
      int* ix = idx + tid * NP;
      const int* piv = piv_global + tid * NP;
      for (int k = NP - 1; k >= 0; --k) {
        const int pk = piv[k];
        if (pk != k) { const int t = ix[k]; ix[k] = ix[pk]; ix[pk] = t; }
      }
    }
    __syncthreads();
    for (int spin = 0; spin < NS; ++spin) {
      const int* ix = idx + spin * NP;
      for (int r = warp; r < NP; r += L2_THREADS / 32) {
        double* rp = Gc + (size_t)spin * NP * NP + (size_t)r * NP;
        for (int c = lane; c < NP; c += 32) rowbuf[c] = rp[c];
        __syncwarp();
        for (int c = lane; c < NP; c += 32) rp[c] = rowbuf[ix[c]];
        __syncwarp();
      }
    }
  }
  __syncthreads();
}

// ---- sweep-start G = inv(I + prod B) in memory ----------------------------------------------------------------
template <bool CL, int NT = 4>
__device__ void l2_recompute(double* __restrict__ Gc, double* __restrict__ Tc, int NP, int KD, const int8_t* field, int l0,
                             const SweepParams& p, L2Smem& sm, int* piv_global) {
  const int L = p.n_slices;
  const int tid = threadIdx.x;
  for (int spin = 0; spin < 2; ++spin) {
    double* G = Gc + (size_t)spin * NP * NP;
    double* T = Tc + (size_t)spin * NP * NP;
    // buffers alternate so that the last product reads T and writes row-major into G
    double* cur = (L % 2 == 0) ? T : G;
    double* oth = (L % 2 == 0) ? G : T;
    int l = (l0 - 1 + L) % L;
    const int8_t* hl = field + (size_t)l * NP;
    if (L == 1) {
      for (int r = 0; r < NP; ++r)
        for (int c = tid; c < NP; c += L2_THREADS)
          G[(size_t)r * NP + c] = p.E[(size_t)r * NP + c] * hs_v(hl[c], spin, p) + (r == c ? 1.0 : 0.0);
    } else {
      // first factor, stored k-major (transposed): cur[c][r] = E[r][c] * v_c   (columns split over the cluster)
      // eight independent loads per thread in flight (one column per iteration was a chain of NP dependent load -> store round
      // trips: 3.7 % of the stall samples of the N = 576 launch, profiles/r02l_cfg5_ncu_lines.txt)
      const int c_first = CL ? sm.crank : 0, c_step = CL ? sm.cs : 1;
      for (int c0 = c_first; c0 < NP; c0 += 8 * c_step) {
        for (int r = tid; r < NP; r += L2_THREADS) {
          double x[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) { const int c = c0 + q * c_step; x[q] = (c < NP) ? p.Et[(size_t)c * NP + r] : 0.0; }
#pragma unroll
          for (int q = 0; q < 8; ++q) { const int c = c0 + q * c_step; if (c < NP) cur[(size_t)c * NP + r] = x[q] * hs_v(hl[c], spin, p); }
        }
      }
    }
    if (CL) l2_cluster_sync(sm.cs); else __syncthreads();
    for (int m = 1; m < L; ++m) {
      l = (l0 - 1 - m + 2 * L) % L;
      L2Epilogue ep;
      ep.hcol = field + (size_t)l * NP;
      ep.transposed_out = (m != L - 1);
      ep.add_identity = (m == L - 1);
      l2_gemm<CL, NT>(cur, p.E, oth, NP, spin, ep, p, sm);
      double* t = cur; cur = oth; oth = t;
    }
  }
  // the Gauss-Jordan inverse is a serial chain of pivots: rank 0 of a cluster runs it alone
  if (!CL || sm.crank == 0) l2_gj_inverse<2>(Gc, NP, KD, sm, piv_global);
  if (CL && sm.cs > 1) l2_cluster_sync(sm.cs);
}

// ---- wrap from slice l to l-1 ------------------------------------------------------------------------------
template <bool PHYS, bool CL, int NT = 4>
__device__ void l2_wrap(double* __restrict__ Gc, double* __restrict__ Tc, int NP, const int8_t* hprev, const SweepParams& p, L2Smem& sm) {
  for (int spin = 0; spin < 2; ++spin) {
    double* G = Gc + (size_t)spin * NP * NP;
    double* T = Tc + (size_t)spin * NP * NP;
    L2Epilogue e1;
    e1.transposed_out = true;
    e1.col_outer = LQMC_L2_COL_OUTER;
    l2_gemm<CL, NT>(PHYS ? p.Eit : p.Et, G, T, NP, spin, e1, p, sm);           // T^T = (E G)^T   (or E^-1 G)
    L2Epilogue e2;
    e2.hrow = hprev; e2.hcol = hprev;
    e2.row_inv = PHYS; e2.col_inv = !PHYS;
    l2_gemm<CL, NT>(T, PHYS ? p.E : p.Ei, G, NP, spin, e2, p, sm);              // G = D (E G E^-1) D^-1
  }
}

struct L2Params {
  L2TmaMaps maps;
  SweepParams p;
  double* T;
  int use_tmem;   // slice path: 0 shared memory, 1 tensor memory (NP <= 256), 2 / 3 tensor memory with 2 / 3 columns per thread
  int* piv;       // [chain][2][NP] scratch
  int NP, KD;
  int cluster;    // CTAs per chain (thread-block cluster size; 1 = one CTA per chain)
  int gemm_stages;   // ring depth of the 64 x 192 GEMMs (one-CTA-per-SM kernels)
};

// TMEM: 0 shared-memory slice path; 1 tensor-memory path, one column per thread (NP <= 256); 2 / 3 several columns per thread
// (384 < NP <= 512: 2 x depth 24; 512 < NP <= 768: 3 x depth 16 where it fits; one CTA per SM at those sizes)
template <bool EXACT, bool PHYS, int TMEM, bool CL = false>
__global__ void __launch_bounds__(L2_THREADS, TMEM >= 2 ? 1 : 2) sweep_l2_kernel(const __grid_constant__ L2Params lp) {
  // GEMM warp tile: 32 x 48 (block tile 64 x 192) in the one-CTA-per-SM instantiations, which may use 255 registers
  constexpr int GNT = (TMEM >= 2 && LQMC_L2_WIDE_TILE) ? 6 : 4;
  extern __shared__ __align__(1024) unsigned char smem_raw[];       // 128-byte-swizzled TMA boxes need 1 KB-aligned stages
  const SweepParams& p = lp.p;
  const int NP = lp.NP, KD = lp.KD;
  L2Smem sm(smem_raw, NP, KD);
  sm.maps = &lp.maps;        // the launcher refuses to launch without valid tensor maps; the 1 KB alignment is checked below
  if (CL) { sm.cs = lp.cluster; sm.crank = (int)(blockIdx.x % (unsigned)lp.cluster); }      // cluster dimension = (size, 1, 1): rank = blockIdx.x % size
  const int chain = CL ? (int)(blockIdx.x / (unsigned)lp.cluster) : (int)blockIdx.x, tid = threadIdx.x;
  const int N = p.n_sites, L = p.n_slices;
  int8_t* field = p.field + (size_t)chain * L * NP;
  double* Gc = p.G + (size_t)chain * 2 * NP * NP;
  double* Tc = lp.T + (size_t)chain * 2 * NP * NP;
  int* piv = lp.piv + (size_t)chain * 2 * NP;
  int n_accepted = 0;
  uint32_t tm_base = 0;
  if (TMEM == 1) tm_base = tmem_alloc_cta(reinterpret_cast<uint32_t*>(sm.hist + 63));
  if (TMEM >= 2) tm_base = tmem_alloc_cta_x(reinterpret_cast<uint32_t*>(sm.hist + 63));
#if LQMC_L2_STAGING_TMA
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();          // the swizzled TMA boxes assume 1 KB-aligned stages
#endif
  if (tid == 0) {
    for (int s0 = 0; s0 < L2_TMA_STAGES_MAX; ++s0) { mbar_init(sm.full + s0, 1); mbar_init(sm.full + L2_TMA_STAGES_MAX + s0, L2_THREADS / 32); }
    reinterpret_cast<uint32_t*>(sm.hist + 62)[0] = 0;        // panels consumed so far (TMA staging: stage / phase parity)
    reinterpret_cast<uint32_t*>(sm.hist + 62)[1] = (uint32_t)lp.gemm_stages;   // ring depth of the wide-tile GEMMs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  for (int sweep = 0; sweep < p.n_sweeps; ++sweep) {
#ifdef LQMC_PHASE_CLOCKS
    long long kt_rec = 0, kt_slice = 0, kt_wrap = 0, kt0 = clock64();
#define LQMC_KT(acc) { const long long kt1 = clock64(); acc += kt1 - kt0; kt0 = kt1; }
#else
#define LQMC_KT(acc)
#endif
    if (p.do_recompute) l2_recompute<CL, GNT>(Gc, Tc, NP, KD, field, p.recompute_l0, p, sm, piv);
    LQMC_KT(kt_rec)
    for (int step = p.step_lo; step < p.step_hi; ++step) {
      const int l = L - 1 - step;
      const long long base = (((long long)chain * p.buf_sweeps + p.buf_sweep0 + sweep) * p.buf_steps + (step - p.buf_step0)) * N;
      if (p.wrap_first && step == p.step_lo) {
        __syncthreads();
        l2_wrap<PHYS, CL, GNT>(Gc, Tc, NP, field + (size_t)l * NP, p, sm);
      }
      if (p.do_propose) {
        __syncthreads();
        for (int j = tid; j < NP; j += L2_THREADS) {
          sm.h[j] = field[(size_t)l * NP + j];
          double u = 2.0;
          if (j < N)
            u = (p.uniforms != nullptr)
                    ? p.uniforms[base + j]
                    : lqmc_philox_uniform(p.seed, (uint64_t)(p.chain0 + chain), (uint64_t)(p.sweep0 + sweep), (uint32_t)(step * N + j));
          sm.u[j] = u;
        }
        LQMC_KT(kt_wrap)
        if (TMEM == 1) l2_propose_slice_tmem<EXACT, PHYS, CL>(Gc, Tc, NP, sm, p, base, n_accepted, tm_base);
        else if (TMEM == 2) l2_propose_slice_tmemx<EXACT, PHYS, 2, 24>(Gc, Tc, NP, sm, p, base, n_accepted, tm_base);
        else if (TMEM == 3) l2_propose_slice_tmemx<EXACT, PHYS, 3, 16>(Gc, Tc, NP, sm, p, base, n_accepted, tm_base);
        else l2_propose_slice<EXACT, PHYS, L2_MAXQ>(Gc, NP, KD, sm, p, base, n_accepted);
        __syncthreads();
        for (int j = tid; j < NP; j += L2_THREADS) field[(size_t)l * NP + j] = sm.h[j];
        LQMC_KT(kt_slice)
      }
      if (p.do_wrap && l > 0 && !(p.skip_last_wrap && step == p.step_hi - 1)) {
        __syncthreads();
        l2_wrap<PHYS, CL, GNT>(Gc, Tc, NP, field + (size_t)(l - 1) * NP, p, sm);
        LQMC_KT(kt_wrap)
      }
    }
#ifdef LQMC_PHASE_CLOCKS
    if (tid == 0 && p.do_recompute) { double* ob = p.obs_sum + (size_t)chain * 3 * N; ob[16] = (double)kt_rec; ob[17] = (double)kt_slice; ob[18] = (double)kt_wrap; }
#endif
    if (p.measure && (!CL || sm.crank == 0)) {
      __syncthreads();
      for (int spin = 0; spin < 2; ++spin) {
        const double* G = Gc + (size_t)spin * NP * NP;
        double* gs = p.g_sum + ((size_t)chain * 2 + spin) * N * N;
        for (int r = 0; r < N; ++r)
          for (int c = tid; c < N; c += L2_THREADS) gs[(size_t)r * N + c] += G[(size_t)r * NP + c];
      }
      for (int j = tid; j < N; j += L2_THREADS) {
        const double nu = 1.0 - Gc[(size_t)j * NP + j], nd = 1.0 - Gc[(size_t)NP * NP + (size_t)j * NP + j];
        double* ob = p.obs_sum + (size_t)chain * 3 * N;
        ob[j] += nu; ob[N + j] += nd; ob[2 * N + j] += nu * nd;
      }
      if (tid == 0) p.n_meas[chain] += 1;
      __syncthreads();
    }
    if (CL && sm.cs > 1) l2_cluster_sync(sm.cs);  // rank 0 has read G for the measurement before the next sweep's product overwrites it
  }
  if (tid == 0 && n_accepted && (!CL || sm.crank == 0)) p.n_acc[chain] += n_accepted;
  if (TMEM == 1) tmem_free_cta(tm_base);
  if (TMEM >= 2) tmem_free_cta_x(tm_base);
}

// ---- host side ------------------------------------------------------------------------------------------------
struct L2Host { int* piv = nullptr; };

inline int l2_alloc(L2Workspace& w, int n_sites, int np, int n_slices, int n_chains, char* err, size_t errlen) {
  (void)n_sites; (void)n_slices;
  const size_t bytes = (size_t)n_chains * 2 * np * np * sizeof(double) + (size_t)n_chains * 2 * np * sizeof(int);
  if (cudaMalloc(&w.T, bytes) != cudaSuccess) {
    snprintf(err, errlen, "cudaMalloc of the %zu-byte GEMM workspace failed", bytes);
    return 4;
  }
  // two CTAs per SM up to NP = 384 (228 KiB per SM, 1 KiB reserved per CTA: 113 KiB each), one above
  const size_t budget = (np <= 384) ? 113 * 1024 : 220 * 1024;
  int kd = 1;
  while (kd < 64 && l2_smem_bytes(np, kd + 1) <= budget) ++kd;
  if (const char* env = getenv("LQMC_L2_KD")) { const int v = atoi(env); if (v >= 1 && v <= kd) kd = v; }   // experiments
  w.kd = kd;
  w.smem = l2_smem_bytes(np, kd);
  if (np <= L2_THREADS) {
    // the tensor-memory path keeps U [2][L2_KDT][NP] in the region the generic path sizes for U and W
    size_t region = (size_t)4 * kd * np;
    const size_t gemm = L2_GEMM_DOUBLES;
    if (region < gemm) region = gemm;
    w.tmem_mode = ((size_t)2 * L2_KDT * np <= region) ? 1 : 0;
  } else if (np > 384 && np <= 3 * L2_THREADS) {
    // several columns per thread, one CTA per SM: U [2][24][NP] (two columns, NP <= 512) / [2][16][NP] (three columns) in the
    // generic path's U / W region
    const size_t region = (size_t)4 * kd * np;
    const int mode = (np <= 2 * L2_THREADS) ? 2 : 3;
    const size_t need = (size_t)2 * (mode == 2 ? 24 : 16) * np;
    // these instantiations stage 64 x 192 GEMM tiles: 4 stages x 32 KB alias the U / W region
    const size_t wide_stage = (size_t)L2_BK * (L2_BM + 192);        // doubles per 32 KB stage
    w.tmem_mode = (need <= region && L2_TMA_STAGES * wide_stage <= region) ? mode : 0;
    int st = (int)(region / wide_stage);
    if (const char* env = getenv("LQMC_L2_GEMM_STAGES")) { const int v = atoi(env); if (v >= 2 && v <= st) st = v; }   // experiments
    w.gemm_stages = st > L2_TMA_STAGES_MAX ? L2_TMA_STAGES_MAX : st;
  }
  return 0;
}
inline void l2_free(L2Workspace& w) { if (w.T) cudaFree(w.T); w.T = nullptr; }

// One 2-D f64 tensor map over a buffer of `rows` rows of NP doubles: box 16 x 16 (128-byte rows), 128-byte swizzle, zero fill
// outside.  cuTensorMapEncodeTiled comes from the driver through the runtime (no link against libcuda).
inline bool l2_encode_map(CUtensorMap* out, const double* base, int np, size_t rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || !sym) return false;
    fn = reinterpret_cast<EncodeFn>(sym);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)np, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)np * sizeof(double)};
  const cuuint32_t box[2] = {16, 16};
  const cuuint32_t estr[2] = {1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline void l2_build_maps(L2Workspace& w, const SweepParams& p, int np) {
  L2TmaMaps& m = w.maps;
  if (m.valid && m.pG == p.G && m.pT == w.T && m.pE == p.E && m.pEt == p.Et && m.pEi == p.Ei && m.pEit == p.Eit) return;
  memset(&m, 0, sizeof(m));
#if LQMC_L2_STAGING_TMA
  const size_t rows = (size_t)p.n_chains * 2 * np;
  bool ok = l2_encode_map(&m.G, p.G, np, rows) && l2_encode_map(&m.T, w.T, np, rows) && l2_encode_map(&m.E, p.E, np, np) &&
            l2_encode_map(&m.Et, p.Et, np, np) && l2_encode_map(&m.Ei, p.Ei, np, np) && l2_encode_map(&m.Eit, p.Eit, np, np);
  m.pG = p.G; m.pT = w.T; m.pE = p.E; m.pEt = p.Et; m.pEi = p.Ei; m.pEit = p.Eit;
  m.n_elems = (long long)rows * np;
  m.valid = ok ? 1 : 0;
#endif
}

inline int launch_l2(L2Workspace& w, const SweepParams& p, int np, uint32_t flags, cudaStream_t s, long long* launches, char* err,
                     size_t errlen) {
  L2Params lp;
  l2_build_maps(w, p, np);
#if LQMC_L2_STAGING_TMA
  if (!w.maps.valid) { snprintf(err, errlen, "cuTensorMapEncodeTiled failed for the GEMM operands (driver too old for TMA tensor maps?)"); return 2; }
#endif
  lp.maps = w.maps;
  lp.p = p; lp.T = w.T; lp.NP = np; lp.KD = w.kd;
  lp.cluster = 1;
  lp.gemm_stages = w.gemm_stages;
  lp.use_tmem = p.do_propose ? w.tmem_mode : 0;
  if (const char* env = getenv("LQMC_L2_SLICE_PATH")) { if (strcmp(env, "smem") == 0) lp.use_tmem = 0; }     // experiments
  lp.piv = reinterpret_cast<int*>(w.T + (size_t)p.n_chains * 2 * np * np);
  const bool exact = !(flags & 0x2u), phys = (flags & 0x1u) != 0;
  // Strong scaling: with few chains, one chain per cluster of CTAs (one CTA per SM: the cluster variant asks for more than half an
  // SM's shared memory).  Largest cluster size whose clusters are all co-resident for this chain count; NP = 256 tensor-memory
  // path only (the split flush / tile schedule is written for it).  LQMC_L2_CLUSTER = 1 / 2 / 4 / 8 overrides (experiments).
  const size_t smem_cluster = (w.smem > (size_t)116 * 1024) ? w.smem : (size_t)118 * 1024;
  auto pick_cluster = [&](auto kernel) -> int {
#if LQMC_L2_STAGING_TMA
    if (np != 256 || lp.use_tmem != 1 || w.tmem_mode != 1) return 1;
    int forced = 0;
    if (const char* env = getenv("LQMC_L2_CLUSTER")) forced = atoi(env);
    if (forced == 1) return 1;
    int n_sm = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (!forced && 2 * p.n_chains > n_sm) return 1;              // enough chains to fill the GPU with one CTA each
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cluster) != cudaSuccess) return 1;
    for (int cs = 8; cs >= 2; --cs) {
      if (forced && cs != forced) continue;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(p.n_chains * cs)); cfg.blockDim = dim3(L2_THREADS); cfg.dynamicSmemBytes = smem_cluster;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int max_clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); continue; }
      if (max_clusters >= p.n_chains) return cs;
    }
#endif
    return 1;
  };
  // kernel: one CTA per chain; kernel_cl: the same kernel with the cluster code compiled in (nullptr where it does not exist)
  auto go = [&](auto kernel, auto kernel_cl) -> int {
    int cs = 1;
    if constexpr (!std::is_same_v<decltype(kernel_cl), std::nullptr_t>) cs = pick_cluster(kernel_cl);
    lp.cluster = cs;
    w.last_cluster = cs;
    cudaError_t e = cudaSuccess;
    if (cs > 1) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(p.n_chains * cs)); cfg.blockDim = dim3(L2_THREADS); cfg.dynamicSmemBytes = smem_cluster; cfg.stream = s;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      if constexpr (!std::is_same_v<decltype(kernel_cl), std::nullptr_t>) e = cudaLaunchKernelEx(&cfg, kernel_cl, lp);
    } else {
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.smem);
      if (e == cudaSuccess) {
        kernel<<<p.n_chains, L2_THREADS, w.smem, s>>>(lp);
        e = cudaGetLastError();
      }
    }
    if (e != cudaSuccess) { snprintf(err, errlen, "sweep_l2_kernel launch failed: %s", cudaGetErrorString(e)); return 2; }
    *launches += 1;
    return 0;
  };
  auto pick = [&](auto tm) -> int {
    constexpr int TM = decltype(tm)::value;
    if constexpr (TM == 1) {          // the cluster variant exists for the NP <= 256 tensor-memory path only
      if (exact && !phys) return go(sweep_l2_kernel<true, false, 1, false>, sweep_l2_kernel<true, false, 1, true>);
      if (!exact && !phys) return go(sweep_l2_kernel<false, false, 1, false>, sweep_l2_kernel<false, false, 1, true>);
      if (exact && phys) return go(sweep_l2_kernel<true, true, 1, false>, sweep_l2_kernel<true, true, 1, true>);
      return go(sweep_l2_kernel<false, true, 1, false>, sweep_l2_kernel<false, true, 1, true>);
    } else {
      if (exact && !phys) return go(sweep_l2_kernel<true, false, TM, false>, nullptr);
      if (!exact && !phys) return go(sweep_l2_kernel<false, false, TM, false>, nullptr);
      if (exact && phys) return go(sweep_l2_kernel<true, true, TM, false>, nullptr);
      return go(sweep_l2_kernel<false, true, TM, false>, nullptr);
    }
  };
  switch (lp.use_tmem) {
    case 1: return pick(std::integral_constant<int, 1>());
    case 2: return pick(std::integral_constant<int, 2>());
    case 3: return pick(std::integral_constant<int, 3>());
    default: return pick(std::integral_constant<int, 0>());
  }
}

}  // namespace lqmc
