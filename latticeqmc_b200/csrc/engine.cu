// C-ABI implementation of the B200 sweep engine (include/lqmc_b200.h).
//
// Host side of the drop-in boundary: owns the device state of `n_chains` Markov chains, converts
// between the reference's host layouts (Configuration.config, (gf_up, gf_dn)) and the kernels'
// padded device layouts, and launches the sweep kernels.  No CPU fallback exists: every entry
// point that computes launches a CUDA kernel or fails.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>
#include <vector>
#include <new>

#include "../../include/lqmc_b200.h"
#include "sweep_reg.cuh"
#include "sweep_l2.cuh"
#include "stab.cuh"
#include "sweep_det.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t err__ = (call);                                                                    \
    if (err__ != cudaSuccess)                                                                      \
      return fail(LQMC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
  } while (0)

}  // namespace

struct lqmc_engine {
  int device = 0;
  int N = 0, L = 0, C = 0, NP = 0;
  uint32_t flags = 0;
  bool family_reg = true;
  double lamb = 0;
  double hs[4] = {1, 1, 0, 0};
  cudaStream_t stream = nullptr;       // non-blocking: every engine-side memset / copy below is issued ON it, never on the legacy stream
  cudaEvent_t foreignDone = nullptr;   // recorded on a caller-owned stream after lqmc_sweep_async queued work there
  bool foreignPending = false;
  bool detCarried = false;             // the next lqmc_sweep_det starts from the old_det already on the device (lqmc_set_det)
  // device state
  double *dE = nullptr, *dEt = nullptr, *dEi = nullptr, *dEit = nullptr;
  int8_t* dField = nullptr;
  int8_t* dFieldRaw = nullptr;     // [chain][site][slice]: the host layout, converted on the device
  int* dBad = nullptr;
  double* dG = nullptr;
  double* dGsum = nullptr;
  double* dObs = nullptr;
  long long* dNmeas = nullptr;
  long long* dNacc = nullptr;
  double* dDetOld = nullptr;       // det mode: old_det per chain (lqmc.py:236-259)
  double* dDetWork = nullptr;      // det mode, N > 64: [chain][4][N][N] matrices
  double* dUni = nullptr;   size_t uniCap = 0;     // staged host uniforms
  double* dTrRatio = nullptr; uint8_t* dTrAcc = nullptr; size_t trCap = 0; size_t trCount = 0;
  lqmc::L2Workspace l2;
  long long sweep_counter = 0;
  long long chain0 = 0;
  long long launches = 0;
  int8_t* hostField = nullptr;     // pinned staging for the layout conversions
  double* hostG = nullptr;
  // stabilised recompute (stab.cuh), allocated on first use
  int stab_every = 0;
  struct Stab {
    bool ready = false;
    int NPs = 0, nfrag = 4, kd = 1;
    double* X[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // [2C][NPs^2] work matrices
    // left stack of the two-sided scheme: UDV of C_j C_{j+1} ... C_{nseg-1} for every segment j
    int stack_k = 0, nseg = 0;
    double* stU = nullptr; double* stV = nullptr; double* stD = nullptr;   // [nseg][2C][NPs^2], [nseg][2C][NPs]
    double* dvec = nullptr;       // [2C][NPs]
    double* tfac = nullptr;       // [2C][NPs * ST_NB]
    int* perm = nullptr;          // [2C][NPs]
    int* piv = nullptr;           // [2C][NPs]
    double* E = nullptr; double* Et = nullptr;   // exp(-dtau K) padded to NPs (own copies when NPs != NP)
    bool own_E = false;
    size_t smem_gemm = 0, smem_qr = 0, smem_inv = 0;
  } st;
  std::vector<double> hE;          // host copy of exp_k (N x N)
};

namespace {

// host layout [chain][site][slice]  <->  device layout [chain][slice][NP] (pad = +1); one CTA per (chain, 32-slice block)
__global__ void field_to_device_kernel(const int8_t* __restrict__ raw, int8_t* __restrict__ dev, int N, int L, int NP, int* bad) {
  __shared__ int8_t tile[32][33];
  const int c = blockIdx.y, l0 = blockIdx.x * 32;
  const int8_t* src = raw + (size_t)c * N * L;
  int8_t* dst = dev + (size_t)c * L * NP;
  for (int i0 = 0; i0 < NP; i0 += 32) {
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const int i = i0 + r, l = l0 + threadIdx.x;
      int8_t v = 1;
      if (i < N && l < L) { v = src[(size_t)i * L + l]; if (v != 1 && v != -1) atomicExch(bad, 1 + (int)(((size_t)c * N + i) * L + l)); }
      tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const int l = l0 + r, i = i0 + threadIdx.x;
      if (l < L && i < NP) dst[(size_t)l * NP + i] = tile[threadIdx.x][r];
    }
    __syncthreads();
  }
}
__global__ void field_to_host_kernel(const int8_t* __restrict__ dev, int8_t* __restrict__ raw, int N, int L, int NP) {
  __shared__ int8_t tile[32][33];
  const int c = blockIdx.y, l0 = blockIdx.x * 32;
  const int8_t* src = dev + (size_t)c * L * NP;
  int8_t* dst = raw + (size_t)c * N * L;
  for (int i0 = 0; i0 < N; i0 += 32) {
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const int l = l0 + r, i = i0 + threadIdx.x;
      tile[r][threadIdx.x] = (l < L && i < NP) ? src[(size_t)l * NP + i] : (int8_t)1;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const int i = i0 + r, l = l0 + threadIdx.x;
      if (i < N && l < L) dst[(size_t)i * L + l] = tile[threadIdx.x][r];
    }
    __syncthreads();
  }
}

// Everything this engine has queued: its own stream, and the caller's stream if lqmc_sweep_async was given one (an event
// recorded there at submit time - valid even if the caller has destroyed the stream since).
int drain(lqmc_engine* e) {
  if (e->foreignPending) {
    CU(cudaEventSynchronize(e->foreignDone));
    e->foreignPending = false;
  }
  CU(cudaStreamSynchronize(e->stream));
  return LQMC_OK;
}
// Later work on the engine's stream must not overtake sweeps still running on a caller's stream.
int order_after_foreign(lqmc_engine* e) {
  if (e->foreignPending) CU(cudaStreamWaitEvent(e->stream, e->foreignDone, 0));
  return LQMC_OK;
}

size_t field_bytes(const lqmc_engine* e) { return (size_t)e->C * e->L * e->NP; }
size_t g_elems(const lqmc_engine* e) { return (size_t)e->C * 2 * e->NP * e->NP; }

int ensure_trace(lqmc_engine* e, size_t count) {
  if (!(e->flags & LQMC_TRACE)) { e->trCount = 0; return LQMC_OK; }
  if (count > e->trCap) {
    if (e->dTrRatio) cudaFree(e->dTrRatio);
    if (e->dTrAcc) cudaFree(e->dTrAcc);
    e->dTrRatio = nullptr; e->dTrAcc = nullptr; e->trCap = 0;
    CU(cudaMalloc(&e->dTrRatio, count * sizeof(double)));
    CU(cudaMalloc(&e->dTrAcc, count));
    e->trCap = count;
  }
  e->trCount = count;
  return LQMC_OK;
}

int stage_uniforms(lqmc_engine* e, const double* host, size_t count, cudaStream_t s) {
  if (count > e->uniCap) {
    if (e->dUni) cudaFree(e->dUni);
    e->dUni = nullptr; e->uniCap = 0;
    CU(cudaMalloc(&e->dUni, count * sizeof(double)));
    e->uniCap = count;
  }
  CU(cudaMemcpyAsync(e->dUni, host, count * sizeof(double), cudaMemcpyHostToDevice, s));
  return LQMC_OK;
}

template <class C>
int launch_reg(lqmc_engine* e, const lqmc::SweepParams& p, cudaStream_t s) {
  const bool exact = !(e->flags & LQMC_ARITH_FMA);
  const bool phys = (e->flags & LQMC_MODE_PHYSICS) != 0;
  const size_t smem = C::smem_bytes;
  auto go = [&](auto kernel) -> int {
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<e->C, C::THREADS, smem, s>>>(p);
    CU(cudaGetLastError());
    e->launches += 1;
    return LQMC_OK;
  };
  if (exact && !phys) return go(lqmc::sweep_reg_kernel<C, true, false>);
  if (!exact && !phys) return go(lqmc::sweep_reg_kernel<C, false, false>);
  if (exact && phys) return go(lqmc::sweep_reg_kernel<C, true, true>);
  return go(lqmc::sweep_reg_kernel<C, false, true>);
}

// One entry for every phase combination: n_sweeps x [recompute?] + steps [step_lo, step_hi) x [propose?][wrap?]
struct RunSpec {
  int n_sweeps = 1, step_lo = 0, step_hi = 0;
  bool recompute = false, propose = false, wrap = false, measure = false, skip_last_wrap = false, wrap_first = false;
  int l0 = 0;
  const double* d_uniforms = nullptr;
  uint64_t seed = 0;
  long long sweep0 = -1;                 // -1: the engine's sweep counter
  int buf_sweeps = -1, buf_steps = -1, buf_sweep0 = 0, buf_step0 = -1;   // -1: this launch's own extent
  bool keep_trace_extent = false;        // the caller sized the trace buffers for a segmented sweep
};

int run(lqmc_engine* e, const RunSpec& r, cudaStream_t s) {
  CU(cudaSetDevice(e->device));
  lqmc::SweepParams p;
  memset(&p, 0, sizeof(p));
  p.n_sites = e->N; p.n_slices = e->L; p.n_chains = e->C;
  p.E = e->dE; p.Et = e->dEt; p.Ei = e->dEi; p.Eit = e->dEit;
  p.field = e->dField; p.G = e->dG;
  p.uniforms = r.d_uniforms; p.seed = r.seed;
  p.sweep0 = r.sweep0 >= 0 ? r.sweep0 : e->sweep_counter; p.chain0 = e->chain0;
  p.g_sum = e->dGsum; p.obs_sum = e->dObs; p.n_meas = e->dNmeas; p.n_acc = e->dNacc;
  p.n_sweeps = r.n_sweeps; p.step_lo = r.step_lo; p.step_hi = r.step_hi;
  p.do_recompute = r.recompute; p.do_propose = r.propose; p.do_wrap = r.wrap; p.measure = r.measure; p.recompute_l0 = r.l0;
  p.skip_last_wrap = r.skip_last_wrap; p.wrap_first = r.wrap_first;
  p.buf_sweeps = r.buf_sweeps >= 0 ? r.buf_sweeps : r.n_sweeps;
  p.buf_steps = r.buf_steps >= 0 ? r.buf_steps : (r.step_hi - r.step_lo);
  p.buf_sweep0 = r.buf_sweep0;
  p.buf_step0 = r.buf_step0 >= 0 ? r.buf_step0 : r.step_lo;
  p.exp_pl = e->hs[0]; p.exp_ml = e->hs[1]; p.f_p2 = e->hs[2]; p.f_m2 = e->hs[3];
  if (r.propose) {
    if (!r.keep_trace_extent) {
      const size_t count = (size_t)e->C * r.n_sweeps * (r.step_hi - r.step_lo) * e->N;
      int rc = ensure_trace(e, count);
      if (rc) return rc;
    }
    if (e->flags & LQMC_TRACE) { p.tr_ratio = e->dTrRatio; p.tr_acc = e->dTrAcc; }
  }
  int rc;
  if (e->family_reg) {
    switch (e->NP) {
      case 16: rc = launch_reg<lqmc::RegCfg<16, 8, 8>>(e, p, s); break;
      case 32: rc = launch_reg<lqmc::RegCfg<32, 8, 8>>(e, p, s); break;
      case 64: rc = launch_reg<lqmc::RegCfg<64, 16, 8>>(e, p, s); break;
      default: return fail(LQMC_ERR_UNSUPPORTED, "no register-resident kernel for padded size %d", e->NP);
    }
  } else {
    rc = lqmc::launch_l2(e->l2, p, e->NP, e->flags, s, &e->launches, g_err, sizeof(g_err));
    if (rc) return rc;
  }
  return rc;
}

// ---- stabilised recompute: host orchestration of the stab.cuh kernels ---------------------------------------------
template <class K>
int set_smem(K kernel, size_t bytes) {
  CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return LQMC_OK;
}

void stab_free(lqmc_engine* e);
int stab_init_impl(lqmc_engine* e);
int stab_init(lqmc_engine* e) {
  if (e->st.ready) return LQMC_OK;
  const int rc = stab_init_impl(e);
  if (rc) stab_free(e);          // a failed allocation part-way leaves nothing behind (stab_free resets the struct)
  return rc;
}
int stab_init_impl(lqmc_engine* e) {
  auto& st = e->st;
  CU(cudaSetDevice(e->device));
  st.NPs = lqmc::st_padded_size(e->N);
  if (e->N > 64 && st.NPs != e->NP)
    return fail(LQMC_ERR_UNSUPPORTED, "stabilisation kernels and sweep kernels disagree on the padded size (N = %d: %d vs %d)", e->N,
                st.NPs, e->NP);
  st.nfrag = (st.NPs == 64) ? 2 : 4;
  const size_t nm = (size_t)2 * e->C, mat = (size_t)st.NPs * st.NPs;
  for (int i = 0; i < 6; ++i)
    if (cudaMalloc(&st.X[i], nm * mat * sizeof(double)) != cudaSuccess)
      return fail(LQMC_ERR_NOMEM, "cudaMalloc of the stabilisation workspace failed (%zu bytes x 5)", nm * mat * sizeof(double));
  if (cudaMalloc(&st.dvec, nm * st.NPs * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&st.tfac, nm * st.NPs * lqmc::ST_NB * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&st.perm, nm * st.NPs * sizeof(int)) != cudaSuccess || cudaMalloc(&st.piv, nm * st.NPs * sizeof(int)) != cudaSuccess)
    return fail(LQMC_ERR_NOMEM, "cudaMalloc of the stabilisation vectors failed");
  if (st.NPs == e->NP) {
    st.E = e->dE; st.Et = e->dEt; st.own_E = false;
  } else {
    std::vector<double> m(mat, 0.0), mt(mat, 0.0);
    for (int i = 0; i < st.NPs; ++i)
      for (int j = 0; j < st.NPs; ++j) {
        const double v = (i < e->N && j < e->N) ? e->hE[(size_t)i * e->N + j] : (i == j ? 1.0 : 0.0);
        m[(size_t)i * st.NPs + j] = v;
        mt[(size_t)j * st.NPs + i] = v;
      }
    if (cudaMalloc(&st.E, mat * sizeof(double)) != cudaSuccess || cudaMalloc(&st.Et, mat * sizeof(double)) != cudaSuccess)
      return fail(LQMC_ERR_NOMEM, "cudaMalloc of the padded exp_k failed");
    st.own_E = true;
    CU(cudaMemcpy(st.E, m.data(), mat * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(st.Et, mt.data(), mat * sizeof(double), cudaMemcpyHostToDevice));
  }
  st.smem_gemm = (size_t)lqmc::ST_STAGES * lqmc::ST_BK * (lqmc::ST_LDA + 32 * st.nfrag + 4) * sizeof(double);
  st.smem_qr = (st.NPs == 64) ? (size_t)(2 * 64 * 65 + 6 * 64) * sizeof(double) + 64 * sizeof(int)
                              : lqmc::StQrSmem<lqmc::ST_NB>::bytes(st.NPs);
  int rc = 0;
  if (st.nfrag == 4) {
    int kd = 1;
    while (kd < 32 && lqmc::l2_smem_bytes(st.NPs, kd + 1, 1, false) <= 110 * 1024) ++kd;
    st.kd = kd;
    st.smem_inv = lqmc::l2_smem_bytes(st.NPs, kd, 1, false);
    rc |= set_smem(lqmc::st_chain_kernel<4>, st.smem_gemm);
    rc |= set_smem(lqmc::st_v_kernel<4>, st.smem_gemm);
    rc |= set_smem(lqmc::st_gemm_kernel<4>, st.smem_gemm);
    rc |= set_smem(lqmc::st_inverse_l2_kernel, st.smem_inv);
  } else {
    st.smem_inv = lqmc::RegCfg<64, 16, 8>::smem_bytes;
    rc |= set_smem(lqmc::st_chain_kernel<2>, st.smem_gemm);
    rc |= set_smem(lqmc::st_v_kernel<2>, st.smem_gemm);
    rc |= set_smem(lqmc::st_gemm_kernel<2>, st.smem_gemm);
    rc |= set_smem(lqmc::st_inverse_small_kernel, st.smem_inv);
  }
  if (st.NPs == 64) rc |= set_smem(lqmc::st_qr_small_kernel, st.smem_qr);
  else rc |= set_smem(lqmc::st_qr_kernel<lqmc::ST_NB>, st.smem_qr);
  if (rc) return rc;
  st.ready = true;
  return LQMC_OK;
}

void stab_free(lqmc_engine* e) {
  auto& st = e->st;
  for (int i = 0; i < 6; ++i) if (st.X[i]) cudaFree(st.X[i]);
  if (st.stU) cudaFree(st.stU);
  if (st.stV) cudaFree(st.stV);
  if (st.stD) cudaFree(st.stD);
  if (st.dvec) cudaFree(st.dvec);
  if (st.tfac) cudaFree(st.tfac);
  if (st.perm) cudaFree(st.perm);
  if (st.piv) cudaFree(st.piv);
  if (st.own_E) { if (st.E) cudaFree(st.E); if (st.Et) cudaFree(st.Et); }
  st = lqmc_engine::Stab();
}

lqmc::HsConsts hs_consts(const lqmc_engine* e) { return lqmc::HsConsts{e->hs[0], e->hs[1]}; }

// (U, D, V) <- UDV decomposition of  [B_{s_last} ... B_{s_first}] * U D V   (normal)  or of the transposed chain.
// Usrc / Dsrc / Vold may be nullptr (identity).  Uses X[0], X[1] as scratch; outputs must not alias them or Vold.
int stab_absorb(lqmc_engine* e, const double* Usrc, size_t u_stride, const double* Dsrc, size_t d_stride, const double* Vold,
                size_t vold_stride, int s_start, int s_step, int count, bool transposed, double* Uout, size_t uo_stride,
                double* Dout, size_t do_stride, double* Vout, size_t vo_stride, cudaStream_t s) {
  auto& st = e->st;
  const int grid = 2 * e->C;
  lqmc::StChainArgs ca;
  ca.src = Usrc; ca.src_stride = u_stride; ca.dcol = Dsrc; ca.dcol_stride = d_stride;
  ca.buf0 = st.X[0]; ca.buf1 = st.X[1];
  ca.field = e->dField; ca.Eop = transposed ? st.E : st.Et;
  ca.N = e->N; ca.NPs = st.NPs; ca.NPf = e->NP; ca.L = e->L;
  ca.s_start = s_start; ca.s_step = s_step; ca.count = count; ca.transposed = transposed ? 1 : 0;
  ca.hc = hs_consts(e);
  if (st.nfrag == 4) lqmc::st_chain_kernel<4><<<grid, lqmc::ST_THREADS, st.smem_gemm, s>>>(ca);
  else lqmc::st_chain_kernel<2><<<grid, lqmc::ST_THREADS, st.smem_gemm, s>>>(ca);
  double* Mres = st.X[count & 1];
  double* Moth = st.X[(count + 1) & 1];
  lqmc::StQrArgs qa;
  qa.M = Mres; qa.A = Moth; qa.Q = Uout; qa.q_stride = uo_stride; qa.dvec = Dout; qa.d_stride = do_stride;
  qa.tfac = st.tfac; qa.perm = st.perm; qa.N = e->N; qa.NPs = st.NPs;
  if (st.NPs == 64) lqmc::st_qr_small_kernel<<<grid, lqmc::ST_THREADS, st.smem_qr, s>>>(qa);
  else lqmc::st_qr_kernel<lqmc::ST_NB><<<grid, lqmc::ST_THREADS, st.smem_qr, s>>>(qa);
  lqmc::StVArgs va;
  va.R = Moth; va.dvec = Dout; va.d_stride = do_stride; va.perm = st.perm; va.Vold = Vold; va.vold_stride = vold_stride;
  va.At = Mres; va.Vnew = Vout; va.vnew_stride = vo_stride; va.N = e->N; va.NPs = st.NPs; va.hc = hs_consts(e);
  if (st.nfrag == 4) lqmc::st_v_kernel<4><<<grid, lqmc::ST_THREADS, st.smem_gemm, s>>>(va);
  else lqmc::st_v_kernel<2><<<grid, lqmc::ST_THREADS, st.smem_gemm, s>>>(va);
  CU(cudaGetLastError());
  e->launches += 3;
  return LQMC_OK;
}

int stab_invert(lqmc_engine* e, double* M, cudaStream_t s) {
  auto& st = e->st;
  lqmc::StInvArgs ia;
  ia.M = M; ia.piv = st.piv; ia.NPs = st.NPs; ia.KD = st.kd;
  if (st.nfrag == 4) lqmc::st_inverse_l2_kernel<<<2 * e->C, lqmc::L2_THREADS, st.smem_inv, s>>>(ia);
  else lqmc::st_inverse_small_kernel<<<2 * e->C, 128, st.smem_inv, s>>>(ia);
  CU(cudaGetLastError());
  e->launches += 1;
  return LQMC_OK;
}

struct GemmSpec {
  const double* At = nullptr; const double* B = nullptr; double* out = nullptr;
  const double* rvec = nullptr; int rmode = 0; const double* cvec = nullptr; int cmode = 0;
  const double* addend = nullptr;
  bool transposed_out = false;
  double* G = nullptr;
};

int stab_gemm(lqmc_engine* e, const GemmSpec& g, cudaStream_t s) {
  auto& st = e->st;
  const size_t mat = (size_t)st.NPs * st.NPs;
  lqmc::StGemmArgs a;
  a.At = g.At; a.at_stride = mat; a.B = g.B; a.b_stride = mat; a.out = g.out; a.out_stride = mat;
  a.rvec = g.rvec; a.rvec_stride = st.NPs; a.rmode = g.rmode; a.cvec = g.cvec; a.cvec_stride = st.NPs; a.cmode = g.cmode;
  a.addend = g.addend; a.addend_stride = mat; a.G = g.G;
  a.N = e->N; a.NPs = st.NPs; a.NPg = e->NP; a.transposed_out = g.transposed_out ? 1 : 0; a.hc = hs_consts(e);
  if (st.nfrag == 4) lqmc::st_gemm_kernel<4><<<2 * e->C, lqmc::ST_THREADS, st.smem_gemm, s>>>(a);
  else lqmc::st_gemm_kernel<2><<<2 * e->C, lqmc::ST_THREADS, st.smem_gemm, s>>>(a);
  CU(cudaGetLastError());
  e->launches += 1;
  return LQMC_OK;
}

// G(l0) = inv(I + B_{l0-1} ... B_0 B_{L-1} ... B_{l0}) from scratch, `chunk` factors per QR (one-sided UDV)
int stab_recompute_scratch(lqmc_engine* e, int l0, int chunk, cudaStream_t s) {
  int rc = stab_init(e);
  if (rc) return rc;
  auto& st = e->st;
  const size_t mat = (size_t)st.NPs * st.NPs;
  double* XU = st.X[2];
  double* Vcur = st.X[3];
  double* Vnext = st.X[4];
  bool first = true;
  int sl = l0, remaining = e->L;
  while (remaining > 0) {
    const int cnt = remaining < chunk ? remaining : chunk;
    rc = stab_absorb(e, first ? nullptr : XU, mat, first ? nullptr : st.dvec, st.NPs, first ? nullptr : Vcur, mat, sl, +1, cnt, false,
                     XU, mat, st.dvec, st.NPs, Vnext, mat, s);
    if (rc) return rc;
    double* t = Vcur; Vcur = Vnext; Vnext = t;
    first = false;
    sl = (sl + cnt) % e->L;
    remaining -= cnt;
  }
  const int grid = 2 * e->C;
  lqmc::StFormsArgs fa;
  fa.U = XU; fa.u_stride = mat; fa.V = Vcur; fa.v_stride = mat; fa.dvec = st.dvec; fa.d_stride = st.NPs;
  fa.lhsT = st.X[0]; fa.rhs = st.X[1]; fa.NPs = st.NPs;
  lqmc::st_forms_kernel<<<grid, lqmc::ST_THREADS, 0, s>>>(fa);
  e->launches += 1;
  rc = stab_invert(e, st.X[0], s);
  if (rc) return rc;
  GemmSpec g;
  g.At = st.X[1]; g.B = st.X[0]; g.out = Vnext; g.transposed_out = true; g.G = e->dG;
  return stab_gemm(e, g, s);
}

// ---- two-sided scheme: left stack + running right product ------------------------------------------------------------
// Segment j (j = 0 .. nseg-1) covers slices [lo_j, hi_j], hi_j = L-1 - j k, visited in this order by the sweep.
// C_j = B_{hi_j} ... B_{lo_j}.  At the top of segment j the sweep needs
//     G(hi_j + 1) = inv(I + Left_j Right_j),   Left_j = C_j ... C_{nseg-1} (not yet updated),  Right_j = C_0 ... C_{j-1},
// which the sweep kernel wraps down to G(hi_j).  Left_j = U_L D_L V_L comes from the stack built at the sweep start;
// Right_j^T = U_R D_R V_R grows by one transposed chunk per segment.  With D = D_b D_s split at 1,
//     G = U_R D_Rb^-1 [ D_Lb^-1 (U_L^T U_R) D_Rb^-1 + D_Ls (V_L V_R^T) D_Rs ]^-1 D_Lb^-1 U_L^T.
void seg_bounds(const lqmc_engine* e, int j, int* lo, int* hi) {
  const int k = e->stab_every;
  *hi = e->L - 1 - j * k;
  *lo = (e->L - (j + 1) * k > 0) ? e->L - (j + 1) * k : 0;
}

int stab_stack_alloc(lqmc_engine* e) {
  auto& st = e->st;
  int rc = stab_init(e);
  if (rc) return rc;
  const int k = e->stab_every, nseg = (e->L + k - 1) / k;
  if (st.stack_k == k && st.stU) return LQMC_OK;
  if (st.stU) { cudaFree(st.stU); cudaFree(st.stV); cudaFree(st.stD); st.stU = st.stV = st.stD = nullptr; }
  const size_t nm = (size_t)2 * e->C, mat = (size_t)st.NPs * st.NPs;
  if (cudaMalloc(&st.stU, nseg * nm * mat * sizeof(double)) != cudaSuccess || cudaMalloc(&st.stV, nseg * nm * mat * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&st.stD, nseg * nm * st.NPs * sizeof(double)) != cudaSuccess)
    return fail(LQMC_ERR_NOMEM, "cudaMalloc of the UDV stack failed (%d segments x %zu bytes x 2)", nseg, nm * mat * sizeof(double));
  st.stack_k = k; st.nseg = nseg;
  return LQMC_OK;
}

int stab_build_left_stack(lqmc_engine* e, cudaStream_t s) {
  auto& st = e->st;
  const size_t nm = (size_t)2 * e->C, mat = (size_t)st.NPs * st.NPs;
  for (int j = st.nseg - 1; j >= 0; --j) {
    int lo, hi;
    seg_bounds(e, j, &lo, &hi);
    const bool first = (j == st.nseg - 1);
    const double* Uprev = first ? nullptr : st.stU + (size_t)(j + 1) * nm * mat;
    const double* Dprev = first ? nullptr : st.stD + (size_t)(j + 1) * nm * st.NPs;
    const double* Vprev = first ? nullptr : st.stV + (size_t)(j + 1) * nm * mat;
    int rc = stab_absorb(e, Uprev, mat, Dprev, st.NPs, Vprev, mat, lo, +1, hi - lo + 1, false, st.stU + (size_t)j * nm * mat, mat,
                         st.stD + (size_t)j * nm * st.NPs, st.NPs, st.stV + (size_t)j * nm * mat, mat, s);
    if (rc) return rc;
  }
  return LQMC_OK;
}

// state of the running right product: U_R = X[2], D_R = dvec, V_R = X[3] / X[4] (ping-pong)
struct RightState { double* U; double* V; double* Vother; bool empty; };

int stab_transpose(lqmc_engine* e, const double* in, double* out, const double* vec, int mode, cudaStream_t s) {
  auto& st = e->st;
  const size_t mat = (size_t)st.NPs * st.NPs;
  lqmc::StTransposeArgs ta;
  ta.in = in; ta.in_stride = mat; ta.out = out; ta.out_stride = mat; ta.vec = vec; ta.vec_stride = st.NPs; ta.mode = mode; ta.NPs = st.NPs;
  lqmc::st_transpose_kernel<<<2 * e->C, lqmc::ST_THREADS, 0, s>>>(ta);
  CU(cudaGetLastError());
  e->launches += 1;
  return LQMC_OK;
}

// G(hi_j + 1) from stack entry j and the right state, into the sweep kernels' G
int stab_combine(lqmc_engine* e, int j, const RightState& R, cudaStream_t s) {
  auto& st = e->st;
  const size_t nm = (size_t)2 * e->C, mat = (size_t)st.NPs * st.NPs;
  const double* UL = st.stU + (size_t)j * nm * mat;
  const double* VL = st.stV + (size_t)j * nm * mat;
  const double* DL = st.stD + (size_t)j * nm * st.NPs;
  const double* DR = st.dvec;
  double* inner = st.X[0];
  double* S1 = st.X[1];
  double* S2 = st.X[5];
  int rc;
  // innerT = D_Rb^-1 (U_R^T U_L) D_Lb^-1
  GemmSpec g;
  g.At = R.U; g.B = UL; g.out = inner; g.rvec = DR; g.rmode = lqmc::ST_VEC_INV_BIG; g.cvec = DL; g.cmode = lqmc::ST_VEC_INV_BIG;
  if ((rc = stab_gemm(e, g, s))) return rc;
  // innerT += (D_Ls (V_L V_R^T) D_Rs)^T
  if ((rc = stab_transpose(e, VL, S1, nullptr, 0, s))) return rc;
  if ((rc = stab_transpose(e, R.V, S2, nullptr, 0, s))) return rc;
  g = GemmSpec();
  g.At = S1; g.B = S2; g.out = inner; g.transposed_out = true; g.addend = inner;
  g.rvec = DL; g.rmode = lqmc::ST_VEC_SMALL; g.cvec = DR; g.cmode = lqmc::ST_VEC_SMALL;
  if ((rc = stab_gemm(e, g, s))) return rc;
  // W^T = inv(innerT)
  if ((rc = stab_invert(e, inner, s))) return rc;
  // Z = W (D_Lb^-1 U_L^T)
  if ((rc = stab_transpose(e, UL, S1, DL, lqmc::ST_VEC_INV_BIG, s))) return rc;
  g = GemmSpec();
  g.At = inner; g.B = S1; g.out = S2;
  if ((rc = stab_gemm(e, g, s))) return rc;
  // G^T = Z^T (D_Rb^-1 U_R^T)
  if ((rc = stab_transpose(e, R.U, S1, DR, lqmc::ST_VEC_INV_BIG, s))) return rc;
  g = GemmSpec();
  g.At = S2; g.B = S1; g.out = inner; g.transposed_out = true; g.G = e->dG;
  return stab_gemm(e, g, s);
}

int stab_sweeps(lqmc_engine* e, int n_sweeps, const double* d_uniforms, uint64_t seed, bool measure, cudaStream_t s) {
  int rc = stab_stack_alloc(e);
  if (rc) return rc;
  auto& st = e->st;
  const size_t mat = (size_t)st.NPs * st.NPs;
  rc = ensure_trace(e, (size_t)e->C * n_sweeps * e->L * e->N);
  if (rc) return rc;
  const int k = e->stab_every;
  for (int sw = 0; sw < n_sweeps; ++sw) {
    rc = stab_build_left_stack(e, s);
    if (rc) return rc;
    RightState R{st.X[2], st.X[3], st.X[4], true};
    lqmc::StIdentityArgs ia;
    ia.M0 = R.U; ia.M1 = R.V; ia.d = st.dvec; ia.NPs = st.NPs;
    lqmc::st_identity_kernel<<<2 * e->C, lqmc::ST_THREADS, 0, s>>>(ia);
    e->launches += 1;
    for (int j = 0; j < st.nseg; ++j) {
      if (j > 0) {
        int lo, hi;
        seg_bounds(e, j - 1, &lo, &hi);
        rc = stab_absorb(e, R.U, mat, st.dvec, st.NPs, R.V, mat, hi, -1, hi - lo + 1, true, R.U, mat, st.dvec, st.NPs, R.Vother, mat, s);
        if (rc) return rc;
        double* t = R.V; R.V = R.Vother; R.Vother = t;
      }
      rc = stab_combine(e, j, R, s);
      if (rc) return rc;
      RunSpec r;
      r.step_lo = j * k; r.step_hi = (j + 1) * k < e->L ? (j + 1) * k : e->L;
      r.propose = true; r.wrap = true; r.skip_last_wrap = true; r.wrap_first = true;
      r.measure = measure && r.step_hi == e->L;
      r.d_uniforms = d_uniforms; r.seed = seed; r.sweep0 = e->sweep_counter + sw;
      r.buf_sweeps = n_sweeps; r.buf_steps = e->L; r.buf_sweep0 = sw; r.buf_step0 = 0; r.keep_trace_extent = true;
      rc = run(e, r, s);
      if (rc) return rc;
    }
  }
  e->sweep_counter += n_sweeps;
  return LQMC_OK;
}

}  // namespace

extern "C" {

const char* lqmc_last_error(void) { return g_err; }
const char* lqmc_version(void) { return "lqmc_b200 0.1 (sm_100a)"; }

int lqmc_create(lqmc_engine** out, int device, int n_sites, int n_slices, int n_chains, const double* exp_k,
                const double* exp_k_inv, double lamb, const double hs_consts[4], uint32_t flags) {
  if (!out) return fail(LQMC_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (n_sites < 1 || n_slices < 1 || n_chains < 1) return fail(LQMC_ERR_INVALID, "n_sites, n_slices, n_chains must be >= 1");
  if (!exp_k || !exp_k_inv || !hs_consts) return fail(LQMC_ERR_INVALID, "exp_k / exp_k_inv / hs_consts is NULL");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(LQMC_ERR_INVALID, "device %d out of range (%d visible)", device, ndev);
  CU(cudaSetDevice(device));
  lqmc_engine* e = new (std::nothrow) lqmc_engine();
  if (!e) return fail(LQMC_ERR_NOMEM, "out of host memory");
  e->device = device; e->N = n_sites; e->L = n_slices; e->C = n_chains; e->flags = flags; e->lamb = lamb;
  memcpy(e->hs, hs_consts, sizeof(e->hs));
  e->hE.assign(exp_k, exp_k + (size_t)n_sites * n_sites);
  if (n_sites <= 64) {
    e->family_reg = true;
    e->NP = n_sites <= 16 ? 16 : (n_sites <= 32 ? 32 : 64);
  } else {
    e->family_reg = false;
    e->NP = lqmc::l2_padded_size(n_sites);
    if (e->NP <= 0) { delete e; return fail(LQMC_ERR_UNSUPPORTED, "n_sites = %d exceeds the largest supported lattice", n_sites); }
  }
  const int NP = e->NP, N = e->N;
  auto cleanup = [&](int code) { lqmc_destroy(e); return code; };
  if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) return cleanup(fail(LQMC_ERR_CUDA, "cudaStreamCreate failed"));
  if (cudaEventCreateWithFlags(&e->foreignDone, cudaEventDisableTiming) != cudaSuccess) return cleanup(fail(LQMC_ERR_CUDA, "cudaEventCreate failed"));
  // identity-padded operator matrices and their transposes
  std::vector<double> m((size_t)NP * NP), mt((size_t)NP * NP);
  double** dst[4] = {&e->dE, &e->dEt, &e->dEi, &e->dEit};
  for (int which = 0; which < 2; ++which) {
    const double* src = which ? exp_k_inv : exp_k;
    for (int i = 0; i < NP; ++i)
      for (int j = 0; j < NP; ++j) {
        const double v = (i < N && j < N) ? src[(size_t)i * N + j] : (i == j ? 1.0 : 0.0);
        m[(size_t)i * NP + j] = v;
        mt[(size_t)j * NP + i] = v;
      }
    for (int tr = 0; tr < 2; ++tr) {
      double** d = dst[2 * which + tr];
      if (cudaMalloc(d, sizeof(double) * NP * NP) != cudaSuccess) return cleanup(fail(LQMC_ERR_NOMEM, "cudaMalloc(E) failed"));
      if (cudaMemcpy(*d, tr ? mt.data() : m.data(), sizeof(double) * NP * NP, cudaMemcpyHostToDevice) != cudaSuccess)
        return cleanup(fail(LQMC_ERR_CUDA, "upload of exp_k failed"));
    }
  }
  const size_t nG = g_elems(e);
  if (cudaMalloc(&e->dField, field_bytes(e)) != cudaSuccess || cudaMalloc(&e->dG, nG * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&e->dFieldRaw, (size_t)e->C * N * e->L) != cudaSuccess || cudaMalloc(&e->dBad, sizeof(int)) != cudaSuccess ||
      cudaMalloc(&e->dGsum, (size_t)e->C * 2 * N * N * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&e->dObs, (size_t)e->C * 3 * N * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&e->dNmeas, (size_t)e->C * sizeof(long long)) != cudaSuccess ||
      cudaMalloc(&e->dNacc, (size_t)e->C * sizeof(long long)) != cudaSuccess)
    return cleanup(fail(LQMC_ERR_NOMEM, "cudaMalloc of chain state failed (%d chains, N=%d)", e->C, N));
  if (cudaMallocHost(&e->hostField, field_bytes(e)) != cudaSuccess || cudaMallocHost(&e->hostG, nG * sizeof(double)) != cudaSuccess)
    return cleanup(fail(LQMC_ERR_NOMEM, "cudaMallocHost of the staging buffers failed"));
  if (cudaMemsetAsync(e->dField, 1, field_bytes(e), e->stream) != cudaSuccess ||
      cudaMemsetAsync(e->dG, 0, nG * sizeof(double), e->stream) != cudaSuccess)
    return cleanup(fail(LQMC_ERR_CUDA, "initialising the chain state failed"));
  if (!e->family_reg) {
    int rc = lqmc::l2_alloc(e->l2, N, NP, e->L, e->C, g_err, sizeof(g_err));
    if (rc) return cleanup(rc);
  }
  int rc = lqmc_reset_measurements(e);
  if (rc) return cleanup(rc);
  if (cudaStreamSynchronize(e->stream) != cudaSuccess) return cleanup(fail(LQMC_ERR_CUDA, "engine initialisation failed"));
  *out = e;
  return LQMC_OK;
}

void lqmc_destroy(lqmc_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  void* ptrs[] = {e->dE, e->dEt, e->dEi, e->dEit, e->dField, e->dFieldRaw, e->dBad, e->dG, e->dGsum, e->dObs, e->dNmeas, e->dNacc, e->dUni, e->dTrRatio, e->dTrAcc, e->dDetOld, e->dDetWork};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  lqmc::l2_free(e->l2);
  stab_free(e);
  if (e->hostField) cudaFreeHost(e->hostField);
  if (e->hostG) cudaFreeHost(e->hostG);
  if (e->foreignDone) cudaEventDestroy(e->foreignDone);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int lqmc_set_field(lqmc_engine* e, const int8_t* field) {
  if (!e || !field) return fail(LQMC_ERR_INVALID, "engine or field is NULL");
  CU(cudaSetDevice(e->device));
  { int rc = order_after_foreign(e); if (rc) return rc; }
  const int N = e->N, L = e->L, NP = e->NP;
  // raw upload, layout conversion and the +-1 check on the device (a scalar host loop over C*N*L bytes cost more than the copy)
  const size_t raw_bytes = (size_t)e->C * N * L;
  memcpy(e->hostField, field, raw_bytes);
  CU(cudaMemcpyAsync(e->dFieldRaw, e->hostField, raw_bytes, cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemsetAsync(e->dBad, 0, sizeof(int), e->stream));
  field_to_device_kernel<<<dim3((L + 31) / 32, e->C), dim3(32, 8), 0, e->stream>>>(e->dFieldRaw, e->dField, N, L, NP, e->dBad);
  CU(cudaGetLastError());
  int bad = 0;
  CU(cudaMemcpyAsync(&bad, e->dBad, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  if (bad) {
    const size_t idx = (size_t)bad - 1;
    return fail(LQMC_ERR_INVALID, "field[%zu][%zu][%zu] = %d is not +-1", idx / ((size_t)N * L), (idx / L) % N, idx % L,
                (int)field[idx]);
  }
  return LQMC_OK;
}

int lqmc_get_field(lqmc_engine* e, int8_t* field) {
  if (!e || !field) return fail(LQMC_ERR_INVALID, "engine or field is NULL");
  CU(cudaSetDevice(e->device));
  { int rc = order_after_foreign(e); if (rc) return rc; }
  const int N = e->N, L = e->L, NP = e->NP;
  const size_t raw_bytes = (size_t)e->C * N * L;
  field_to_host_kernel<<<dim3((L + 31) / 32, e->C), dim3(32, 8), 0, e->stream>>>(e->dField, e->dFieldRaw, N, L, NP);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(e->hostField, e->dFieldRaw, raw_bytes, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  memcpy(field, e->hostField, raw_bytes);
  return LQMC_OK;
}

int lqmc_set_g(lqmc_engine* e, const double* g) {
  if (!e || !g) return fail(LQMC_ERR_INVALID, "engine or g is NULL");
  CU(cudaSetDevice(e->device));
  { int rc = order_after_foreign(e); if (rc) return rc; }
  const int N = e->N, NP = e->NP;
  const size_t mats = (size_t)e->C * 2;
  if (N != NP) {
    // identity padding first, then the N x N blocks (strided copy straight from the caller's buffer)
    memset(e->hostG, 0, g_elems(e) * sizeof(double));
    for (size_t m = 0; m < mats; ++m)
      for (int i = N; i < NP; ++i) e->hostG[m * NP * NP + (size_t)i * NP + i] = 1.0;
    CU(cudaMemcpyAsync(e->dG, e->hostG, g_elems(e) * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    for (size_t m = 0; m < mats; ++m)
      CU(cudaMemcpy2DAsync(e->dG + m * NP * NP, (size_t)NP * sizeof(double), g + m * N * N, (size_t)N * sizeof(double),
                           (size_t)N * sizeof(double), N, cudaMemcpyHostToDevice, e->stream));
  } else {
    CU(cudaMemcpyAsync(e->dG, g, g_elems(e) * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  }
  CU(cudaStreamSynchronize(e->stream));
  return LQMC_OK;
}

int lqmc_get_g(lqmc_engine* e, double* g) {
  if (!e || !g) return fail(LQMC_ERR_INVALID, "engine or g is NULL");
  CU(cudaSetDevice(e->device));
  { int rc = order_after_foreign(e); if (rc) return rc; }
  const int N = e->N, NP = e->NP;
  const size_t mats = (size_t)e->C * 2;
  if (N != NP) {
    for (size_t m = 0; m < mats; ++m)
      CU(cudaMemcpy2DAsync(g + m * N * N, (size_t)N * sizeof(double), e->dG + m * NP * NP, (size_t)NP * sizeof(double),
                           (size_t)N * sizeof(double), N, cudaMemcpyDeviceToHost, e->stream));
  } else {
    // no padding: one copy straight into the caller's buffer (no staging pass on the host)
    CU(cudaMemcpyAsync(g, e->dG, g_elems(e) * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  }
  CU(cudaStreamSynchronize(e->stream));
  return LQMC_OK;
}

int lqmc_recompute(lqmc_engine* e, int l0) {
  if (!e) return fail(LQMC_ERR_INVALID, "engine is NULL");
  if (l0 < 0 || l0 >= e->L) return fail(LQMC_ERR_INVALID, "l0 = %d outside [0, %d)", l0, e->L);
  RunSpec r;
  r.recompute = true; r.l0 = l0;
  int rc = run(e, r, e->stream);
  if (rc) return rc;
  CU(cudaStreamSynchronize(e->stream));
  return LQMC_OK;
}

int lqmc_recompute_stable(lqmc_engine* e, int l0, int chunk) {
  if (!e) return fail(LQMC_ERR_INVALID, "engine is NULL");
  if (l0 < 0 || l0 >= e->L) return fail(LQMC_ERR_INVALID, "l0 = %d outside [0, %d)", l0, e->L);
  if (chunk < 1) return fail(LQMC_ERR_INVALID, "chunk = %d must be >= 1", chunk);
  CU(cudaSetDevice(e->device));
  int rc = stab_recompute_scratch(e, l0, chunk, e->stream);
  if (rc) return rc;
  CU(cudaStreamSynchronize(e->stream));
  return LQMC_OK;
}

int lqmc_set_stabilization(lqmc_engine* e, int stab_every) {
  if (!e) return fail(LQMC_ERR_INVALID, "engine is NULL");
  if (stab_every < 0) return fail(LQMC_ERR_INVALID, "stab_every = %d is negative", stab_every);
  if (stab_every > 0 && !(e->flags & LQMC_MODE_PHYSICS))
    return fail(LQMC_ERR_INVALID, "stabilisation belongs to physics mode: the reference recurrence recomputes G exactly once per "
                                  "sweep, unstabilised (lqmc.py:303-307)");
  e->stab_every = stab_every;
  return LQMC_OK;
}

int lqmc_slice(lqmc_engine* e, int l, const double* uniforms, uint64_t seed) {
  if (!e) return fail(LQMC_ERR_INVALID, "engine is NULL");
  if (l < 0 || l >= e->L) return fail(LQMC_ERR_INVALID, "slice %d outside [0, %d)", l, e->L);
  const int step = e->L - 1 - l;
  const double* d_u = nullptr;
  if (uniforms) {
    int rc = stage_uniforms(e, uniforms, (size_t)e->C * e->N, e->stream);
    if (rc) return rc;
    d_u = e->dUni;
  }
  RunSpec r;
  r.step_lo = step; r.step_hi = step + 1; r.propose = true; r.d_uniforms = d_u; r.seed = seed;
  int rc = run(e, r, e->stream);
  if (rc) return rc;
  CU(cudaStreamSynchronize(e->stream));
  return LQMC_OK;
}

int lqmc_wrap(lqmc_engine* e, int l) {
  if (!e) return fail(LQMC_ERR_INVALID, "engine is NULL");
  if (l < 1 || l >= e->L) return fail(LQMC_ERR_INVALID, "wrap needs 1 <= l < %d, got %d", e->L, l);
  const int step = e->L - 1 - l;
  RunSpec r;
  r.step_lo = step; r.step_hi = step + 1; r.wrap = true;
  int rc = run(e, r, e->stream);
  if (rc) return rc;
  CU(cudaStreamSynchronize(e->stream));
  return LQMC_OK;
}

int lqmc_sweep_async(lqmc_engine* e, int n_sweeps, const double* d_uniforms, uint64_t seed, int measure, void* stream) {
  if (!e) return fail(LQMC_ERR_INVALID, "engine is NULL");
  if (n_sweeps < 0) return fail(LQMC_ERR_INVALID, "n_sweeps = %d is negative", n_sweeps);
  if (n_sweeps == 0) return LQMC_OK;
  cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
  const bool foreign = (s != e->stream);
  CU(cudaSetDevice(e->device));
  if (!foreign) { int rc0 = order_after_foreign(e); if (rc0) return rc0; }
  const bool phys = (e->flags & LQMC_MODE_PHYSICS) != 0;
  int rc;
  if (phys && e->stab_every > 0) {
    // stabilised schedule (stab.cuh): G is rebuilt from the field by QR/UDV at the top of every segment of stab_every
    // slices and propagated by wraps inside it
    rc = stab_sweeps(e, n_sweeps, d_uniforms, seed, measure != 0, s);
  } else {
    RunSpec r;
    r.n_sweeps = n_sweeps; r.step_lo = 0; r.step_hi = e->L; r.recompute = true; r.propose = true; r.wrap = true;
    r.measure = measure != 0; r.l0 = phys ? e->L - 1 : 0; r.d_uniforms = d_uniforms; r.seed = seed;
    rc = run(e, r, s);
    if (!rc) e->sweep_counter += n_sweeps;
  }
  if (foreign) {            // whatever was queued (even a partial schedule on failure) must be waited for by the getters
    if (cudaEventRecord(e->foreignDone, s) == cudaSuccess) e->foreignPending = true;
  }
  return rc;
}

int lqmc_sync(lqmc_engine* e) {
  if (!e) return fail(LQMC_ERR_INVALID, "engine is NULL");
  CU(cudaSetDevice(e->device));
  return drain(e);
}

int lqmc_sweep(lqmc_engine* e, int n_sweeps, const double* uniforms, uint64_t seed, int measure) {
  if (!e) return fail(LQMC_ERR_INVALID, "engine is NULL");
  if (n_sweeps < 0) return fail(LQMC_ERR_INVALID, "n_sweeps = %d is negative", n_sweeps);
  if (n_sweeps == 0) return LQMC_OK;
  CU(cudaSetDevice(e->device));
  const double* d_u = nullptr;
  if (uniforms) {
    int rc = stage_uniforms(e, uniforms, (size_t)e->C * n_sweeps * e->L * e->N, e->stream);
    if (rc) return rc;
    d_u = e->dUni;
  }
  int rc = lqmc_sweep_async(e, n_sweeps, d_u, seed, measure, e->stream);
  if (rc) return rc;
  CU(cudaStreamSynchronize(e->stream));
  return LQMC_OK;
}

int lqmc_sweep_det(lqmc_engine* e, int n_sweeps, const double* uniforms, uint64_t seed, int measure) {
  if (!e) return fail(LQMC_ERR_INVALID, "engine is NULL");
  if (n_sweeps < 0) return fail(LQMC_ERR_INVALID, "n_sweeps = %d is negative", n_sweeps);
  if (n_sweeps == 0) return LQMC_OK;
  const size_t smem = lqmc::det_smem_bytes(e->N, e->L);
  if (smem > 227 * 1024) return fail(LQMC_ERR_UNSUPPORTED, "det mode needs %zu bytes of shared memory (N = %d, L = %d)", smem, e->N, e->L);
  CU(cudaSetDevice(e->device));
  cudaStream_t s = e->stream;
  if (!e->dDetOld) CU(cudaMalloc(&e->dDetOld, (size_t)e->C * sizeof(double)));
  if (e->N > lqmc::DET_MAX_N && !e->dDetWork)   // large lattices: the four N x N matrices of a chain live in global memory
    CU(cudaMalloc(&e->dDetWork, (size_t)e->C * 4 * e->N * e->N * sizeof(double)));
  const double* d_u = nullptr;
  if (uniforms) {
    int rc = stage_uniforms(e, uniforms, (size_t)e->C * n_sweeps * e->L * e->N, s);
    if (rc) return rc;
    d_u = e->dUni;
  }
  int rc = ensure_trace(e, (size_t)e->C * n_sweeps * e->L * e->N);
  if (rc) return rc;
  CU(cudaFuncSetAttribute(lqmc::sweep_det_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lqmc::DetParams p;
  memset(&p, 0, sizeof(p));
  p.n_sites = e->N; p.n_slices = e->L; p.NPf = e->NP; p.ldE = e->NP;
  p.E = e->dE; p.field = e->dField; p.uniforms = d_u; p.seed = seed; p.chain0 = e->chain0;
  p.buf_sweeps = n_sweeps; p.det_old = e->dDetOld; p.n_acc = e->dNacc;
  if (e->flags & LQMC_TRACE) { p.tr_ratio = e->dTrRatio; p.tr_acc = e->dTrAcc; }
  p.exp_pl = e->hs[0]; p.exp_ml = e->hs[1];
  p.work = e->dDetWork;
  // measured: one launch per sweep, each followed by G = inv(get_m(0)) + accumulation (the sweep kernels' recompute);
  // unmeasured: all sweeps in one launch
  const int per_launch = measure ? 1 : n_sweeps;
  for (int s0 = 0; s0 < n_sweeps; s0 += per_launch) {
    p.n_sweeps = per_launch; p.buf_sweep0 = s0; p.sweep0 = e->sweep_counter + s0; p.init_det = (s0 == 0) && !e->detCarried;
    lqmc::sweep_det_kernel<<<e->C, lqmc::DET_THREADS, smem, s>>>(p);
    CU(cudaGetLastError());
    e->launches += 1;
    if (measure) {
      RunSpec r;
      r.recompute = true; r.l0 = 0; r.measure = true;
      rc = run(e, r, s);
      if (rc) return rc;
    }
  }
  e->sweep_counter += n_sweeps;
  e->detCarried = false;
  CU(cudaStreamSynchronize(s));
  return LQMC_OK;
}

int lqmc_set_det(lqmc_engine* e, const double* det_old) {
  if (!e) return fail(LQMC_ERR_INVALID, "engine is NULL");
  CU(cudaSetDevice(e->device));
  if (!det_old && !e->dDetOld) return fail(LQMC_ERR_INVALID, "no det-mode sweep has run on this engine: there is no old_det to carry");
  if (!e->dDetOld) CU(cudaMalloc(&e->dDetOld, (size_t)e->C * sizeof(double)));
  if (det_old) {
    { int rc = order_after_foreign(e); if (rc) return rc; }
    CU(cudaMemcpyAsync(e->dDetOld, det_old, (size_t)e->C * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
  }
  e->detCarried = true;
  return LQMC_OK;
}

int lqmc_sweep_submit(lqmc_engine* e, int n_sweeps, const double* uniforms, uint64_t seed, int measure) {
  if (!e) return fail(LQMC_ERR_INVALID, "engine is NULL");
  if (n_sweeps < 0) return fail(LQMC_ERR_INVALID, "n_sweeps = %d is negative", n_sweeps);
  if (n_sweeps == 0) return LQMC_OK;
  CU(cudaSetDevice(e->device));
  const double* d_u = nullptr;
  if (uniforms) {
    int rc = stage_uniforms(e, uniforms, (size_t)e->C * n_sweeps * e->L * e->N, e->stream);
    if (rc) return rc;
    d_u = e->dUni;
  }
  return lqmc_sweep_async(e, n_sweeps, d_u, seed, measure, e->stream);
}

int lqmc_get_det(lqmc_engine* e, double* det_old) {
  if (!e || !det_old) return fail(LQMC_ERR_INVALID, "engine or det_old is NULL");
  if (!e->dDetOld) return fail(LQMC_ERR_INVALID, "no det-mode sweep has run on this engine");
  CU(cudaSetDevice(e->device));
  { int rc = drain(e); if (rc) return rc; }
  CU(cudaMemcpyAsync(det_old, e->dDetOld, (size_t)e->C * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return LQMC_OK;
}

int lqmc_get_trace(lqmc_engine* e, uint8_t* acc, double* ratio) {
  if (!e) return fail(LQMC_ERR_INVALID, "engine is NULL");
  if (!(e->flags & LQMC_TRACE)) return fail(LQMC_ERR_INVALID, "engine was created without LQMC_TRACE");
  CU(cudaSetDevice(e->device));
  { int rc = drain(e); if (rc) return rc; }                 // the sweep that writes the trace may still be in flight (lqmc_sweep_submit)
  if (acc) CU(cudaMemcpyAsync(acc, e->dTrAcc, e->trCount, cudaMemcpyDeviceToHost, e->stream));
  if (ratio) CU(cudaMemcpyAsync(ratio, e->dTrRatio, e->trCount * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return LQMC_OK;
}

int lqmc_get_measurements(lqmc_engine* e, double* g_sum, double* obs_sum, int64_t* n_meas, int64_t* n_accepted) {
  if (!e) return fail(LQMC_ERR_INVALID, "engine is NULL");
  CU(cudaSetDevice(e->device));
  { int rc = drain(e); if (rc) return rc; }
  const size_t C = e->C, N = e->N;
  if (g_sum) CU(cudaMemcpyAsync(g_sum, e->dGsum, C * 2 * N * N * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  if (obs_sum) CU(cudaMemcpyAsync(obs_sum, e->dObs, C * 3 * N * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  if (n_meas) CU(cudaMemcpyAsync(n_meas, e->dNmeas, C * sizeof(long long), cudaMemcpyDeviceToHost, e->stream));
  if (n_accepted) CU(cudaMemcpyAsync(n_accepted, e->dNacc, C * sizeof(long long), cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return LQMC_OK;
}

int lqmc_set_measurements(lqmc_engine* e, const double* g_sum, const double* obs_sum, const int64_t* n_meas,
                          const int64_t* n_accepted) {
  if (!e) return fail(LQMC_ERR_INVALID, "engine is NULL");
  CU(cudaSetDevice(e->device));
  { int rc = drain(e); if (rc) return rc; }
  const size_t C = e->C, N = e->N;
  if (g_sum) CU(cudaMemcpyAsync(e->dGsum, g_sum, C * 2 * N * N * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  if (obs_sum) CU(cudaMemcpyAsync(e->dObs, obs_sum, C * 3 * N * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  if (n_meas) CU(cudaMemcpyAsync(e->dNmeas, n_meas, C * sizeof(long long), cudaMemcpyHostToDevice, e->stream));
  if (n_accepted) CU(cudaMemcpyAsync(e->dNacc, n_accepted, C * sizeof(long long), cudaMemcpyHostToDevice, e->stream));
  CU(cudaStreamSynchronize(e->stream));        // the caller's buffers are pageable: do not return before they have been read
  return LQMC_OK;
}

int lqmc_reset_measurements(lqmc_engine* e) {
  if (!e) return fail(LQMC_ERR_INVALID, "engine is NULL");
  CU(cudaSetDevice(e->device));
  const size_t C = e->C, N = e->N;
  // on the engine's stream: ordered after sweeps still in flight there (lqmc_sweep_submit) and before the next one queued
  { int rc = order_after_foreign(e); if (rc) return rc; }
  CU(cudaMemsetAsync(e->dGsum, 0, C * 2 * N * N * sizeof(double), e->stream));
  CU(cudaMemsetAsync(e->dObs, 0, C * 3 * N * sizeof(double), e->stream));
  CU(cudaMemsetAsync(e->dNmeas, 0, C * sizeof(long long), e->stream));
  CU(cudaMemsetAsync(e->dNacc, 0, C * sizeof(long long), e->stream));
  return LQMC_OK;
}

int lqmc_device_ptr(lqmc_engine* e, int which, void** ptr, uint64_t* n_bytes) {
  if (!e || !ptr) return fail(LQMC_ERR_INVALID, "engine or ptr is NULL");
  const size_t C = e->C, N = e->N;
  size_t bytes = 0;
  switch (which) {
    case 0: *ptr = e->dField; bytes = field_bytes(e); break;
    case 1: *ptr = e->dG; bytes = g_elems(e) * sizeof(double); break;
    case 2: *ptr = e->dGsum; bytes = C * 2 * N * N * sizeof(double); break;
    case 3: *ptr = e->dObs; bytes = C * 3 * N * sizeof(double); break;
    case 4: *ptr = e->dNmeas; bytes = C * sizeof(long long); break;
    case 5: *ptr = e->dNacc; bytes = C * sizeof(long long); break;
    default: return fail(LQMC_ERR_INVALID, "unknown buffer id %d", which);
  }
  if (n_bytes) *n_bytes = bytes;
  return LQMC_OK;
}

int lqmc_info(lqmc_engine* e, int* n_pad, int64_t* sweep_counter, int64_t* launches, char family[8]) {
  if (!e) return fail(LQMC_ERR_INVALID, "engine is NULL");
  if (n_pad) *n_pad = e->NP;
  if (sweep_counter) *sweep_counter = e->sweep_counter;
  if (launches) *launches = e->launches;
  if (family) { memset(family, 0, 8); strncpy(family, e->family_reg ? "reg" : "l2", 7); }
  return LQMC_OK;
}

int lqmc_get_cluster(lqmc_engine* e, int* ctas_per_chain) {
  if (!e || !ctas_per_chain) return fail(LQMC_ERR_INVALID, "engine or ctas_per_chain is NULL");
  *ctas_per_chain = e->family_reg ? 1 : e->l2.last_cluster;
  return LQMC_OK;
}

int lqmc_set_sweep_counter(lqmc_engine* e, int64_t counter) {
  if (!e) return fail(LQMC_ERR_INVALID, "engine is NULL");
  e->sweep_counter = counter;
  return LQMC_OK;
}

int lqmc_set_chain_offset(lqmc_engine* e, int64_t chain0) {
  if (!e) return fail(LQMC_ERR_INVALID, "engine is NULL");
  e->chain0 = chain0;
  return LQMC_OK;
}

int lqmc_selftest_division(int device, uint64_t n_samples, uint64_t seed, uint64_t* mismatches) {
  if (!mismatches) return fail(LQMC_ERR_INVALID, "mismatches is NULL");
  CU(cudaSetDevice(device));
  unsigned long long* d_bad = nullptr;
  CU(cudaMalloc(&d_bad, sizeof(unsigned long long)));
  CU(cudaMemset(d_bad, 0, sizeof(unsigned long long)));
  const int blocks = 148 * 8, threads = 256;
  const unsigned long long per_thread = (n_samples + (uint64_t)blocks * threads - 1) / ((uint64_t)blocks * threads);
  lqmc::division_selftest_kernel<<<blocks, threads>>>(per_thread, seed, d_bad);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  unsigned long long bad = 0;
  CU(cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost));
  cudaFree(d_bad);
  *mismatches = bad;
  return LQMC_OK;
}

void lqmc_philox_uniforms(uint64_t seed, uint64_t chain, uint64_t sweep, int n_proposals, double* out) {
  for (int p = 0; p < n_proposals; ++p) out[p] = lqmc_philox_uniform(seed, chain, sweep, (uint32_t)p);
}

}  // extern "C"
