/* lqmc_b200.h - C ABI of the B200 determinant-QMC sweep engine.
 *
 * Drop-in boundary for ONE path of KieDani/LatticeQMC: the Metropolis sweep over the
 * Hubbard-Stratonovich Ising field.  The reference has no FFI or plugin interface (it is
 * pure Python); the seam is the method boundary of `LatticeQMC` in lqmc/lqmc.py.  Each entry
 * point below names the reference method (file:line) whose work it replaces.  The binding a
 * maintainer adds on the reference side is the ctypes stub in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns LQMC_OK (0) or an error code; lqmc_last_error() gives the text of the
 *     most recent failure on the calling thread.  Nothing throws, no Python objects cross.
 *   - host buffers are caller-owned, dense, C order; device memory is engine-owned.
 *   - one engine drives `n_chains` independent Markov chains on one CUDA device; calls on one
 *     engine must be serialised by the caller (like the reference: one LatticeQMC per process).
 *   - field   : int8  [chain][site][slice]        (= a stack of reference Configuration.config arrays,
 *                                                   configuration.py:89,123-124)
 *   - G       : f64   [chain][spin(0=up,1=dn)][N][N]   (= (gf_up, gf_dn) of lqmc.py:347, per chain)
 *   - uniforms: f64   [chain][sweep][step][site], step s visits slice l = L-1-s, one number per
 *                     proposal whether accepted or not (lqmc.py:309-317)
 *   - trace   : ratio f64 / acc u8, same indexing as uniforms.
 */
#ifndef LQMC_B200_H
#define LQMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lqmc_engine lqmc_engine;

enum {
  LQMC_OK = 0,
  LQMC_ERR_INVALID = 1,     /* bad argument (the reference would raise ValueError / IndexError) */
  LQMC_ERR_CUDA = 2,        /* CUDA runtime failure, text in lqmc_last_error() */
  LQMC_ERR_UNSUPPORTED = 3, /* size / mode combination no kernel covers */
  LQMC_ERR_NOMEM = 4
};

/* flags for lqmc_create */
#define LQMC_MODE_PARITY   0x0u /* the reference recurrence, quirks included (lqmc.py:301-347)          */
#define LQMC_MODE_PHYSICS  0x1u /* textbook DQMC in the get_m convention         */
#define LQMC_ARITH_EXACT   0x0u /* ratio / rank-1 with the reference's roundings: separate multiply and
                                   subtract, true division (lqmc.py:314-331)                            */
#define LQMC_ARITH_FMA     0x2u /* contracted FMA and one reciprocal per flip (<= 1 ulp per operation)  */
#define LQMC_TRACE         0x4u /* keep per-proposal ratio / acc of the last lqmc_sweep / lqmc_slice    */

/* Build an engine.  Replaces LatticeQMC.__init__ + set_beta device-side state (lqmc.py:16-117): the
 * host computes dtau, lamb = arccosh(exp(U dtau/2)) and exp_k = expm(-dtau K) exactly as the reference
 * does (lqmc.py:102-106) and hands them over; exp_k_inv = expm(+dtau K).  Matrices are N x N row-major.
 * hs_consts = { exp(+lamb), exp(-lamb), exp(+2 lamb) - 1, exp(-2 lamb) - 1 } evaluated by the host with
 * the same libm/NumPy calls the reference makes (lqmc.py:149,154,314-323), so the device never
 * evaluates exp() and the constants are bit-identical to the reference's. */
int lqmc_create(lqmc_engine** out, int device, int n_sites, int n_slices, int n_chains,
                const double* exp_k, const double* exp_k_inv, double lamb, const double hs_consts[4],
                uint32_t flags);
void lqmc_destroy(lqmc_engine* e);

/* Configuration.config <-> device (configuration.py:86-136).  Values must be +-1. */
int lqmc_set_field(lqmc_engine* e, const int8_t* field);
int lqmc_get_field(lqmc_engine* e, int8_t* field);

/* Teacher forcing / inspection of the Green's functions the sweep carries (locals gf_up, gf_dn of
 * lqmc.py:306-347). */
int lqmc_set_g(lqmc_engine* e, const double* g);
int lqmc_get_g(lqmc_engine* e, double* g);

/* Sweep-start G from the field: get_m(l0, +-1) and np.linalg.inv (lqmc.py:156-185,303-307).
 * Parity mode uses l0 = 0 as the reference does; physics mode passes the slice the sweep is at. */
int lqmc_recompute(lqmc_engine* e, int l0);

/* Numerically stabilised G(l0) = inv(I + B_{l0-1} ... B_0 B_{L-1} ... B_{l0}) in the get_m convention
 * (lqmc.py:156-185): the product is accumulated `chunk` factors at a time as U D V with a column-norm
 * pre-pivoted Householder QR after every chunk, and G = (D_b^-1 U^T + D_s V)^-1 D_b^-1 U^T.  The reference has
 * no counterpart (it inverts the raw product, lqmc.py:303-307, cond ~1e21 at beta >= 8: SURVEY.md H8); this is
 * what physics mode uses.  Works in either mode as an inspection call (it only reads the field, writes G). */
int lqmc_recompute_stable(lqmc_engine* e, int l0, int chunk);

/* Physics mode only: rebuild G with lqmc_recompute_stable every `stab_every` slices inside lqmc_sweep
 * (0 = off: one unstabilised sweep-start product, like the reference).  LQMC_ERR_INVALID in parity mode. */
int lqmc_set_stabilization(lqmc_engine* e, int stab_every);

/* The N proposals of time slice l (lqmc.py:311-335): ratio, Metropolis test `u <= ratio`,
 * Sherman-Morrison rank-1 update of both G, field flip.  uniforms: host f64 [chain][site], or NULL for
 * the device Philox stream of (seed, chain) at the engine's current sweep counter. */
int lqmc_slice(lqmc_engine* e, int l, const double* uniforms, uint64_t seed);

/* Wrap from slice l to l-1 (lqmc.py:338-345): G <- B G B^-1 with the diagonal exp(V) folded into the
 * GEMM epilogue.  l must be >= 1. */
int lqmc_wrap(lqmc_engine* e, int l);

/* n_sweeps full sweeps = n_sweeps calls of LatticeQMC._update_step (lqmc.py:301-347), in one launch.
 * uniforms: host f64 [chain][sweep][step][site] or NULL (device Philox keyed by (seed, chain), counter
 * = global sweep index and proposal index, so results do not depend on how chains are sharded over
 * GPUs).  measure != 0 adds each end-of-sweep G to the accumulators (measure_loop, lqmc.py:356-375). */
int lqmc_sweep(lqmc_engine* e, int n_sweeps, const double* uniforms, uint64_t seed, int measure);

/* Same, with nothing crossing the host boundary: inputs already resident (field via lqmc_set_field /
 * a previous sweep), uniforms from a DEVICE pointer (or NULL = Philox), asynchronous on `stream`
 * (a cudaStream_t; NULL = the engine's own stream).  Pair with lqmc_sync. */
int lqmc_sweep_async(lqmc_engine* e, int n_sweeps, const double* d_uniforms, uint64_t seed, int measure,
                     void* stream);
int lqmc_sync(lqmc_engine* e);
/* lqmc_sweep without the final wait: host uniforms (or NULL = Philox) are staged and the sweeps are queued on the
 * engine's own stream; the call returns as soon as the work is submitted (`uniforms` may be released on return).  Lets
 * one host thread drive many engines at once - a beta / U scan (measure_betas, lqmc/__init__.py:57-96;
 * SerialProcessManager, multiprocessing.py:292-341) runs one engine per parameter point, all concurrently on one GPU,
 * where the reference runs one OS process per point.  Pair with lqmc_sync. */
int lqmc_sweep_submit(lqmc_engine* e, int n_sweeps, const double* uniforms, uint64_t seed, int measure);

/* det mode: n_sweeps x LatticeQMC._update_step_det (lqmc.py:236-259), the reference's slow validation sampler -
 * every proposal flips h[i,l], rebuilds get_m(l, +-1) from the field, and accepts on
 * u <= det(M_up) det(M_dn) / old_det (un-flipping on reject).  One call is one of the reference's loops:
 * old_det is initialised from get_m(0, +-1) at the start of the call (warmup_loop_det lqmc.py:261-270;
 * measure_loop_det :272-299) and carried through the n_sweeps sweeps; with measure != 0, inv(get_m(0, +-1)) is
 * added to the accumulators after every sweep (lqmc.py:293-297).  uniforms / seed / trace as in lqmc_sweep.
 * Independent of the engine's mode flag.  N <= 64: matrices in shared memory; larger lattices: the same code on a global-memory workspace (a validation tool, O(N^3) per
 * proposal); LQMC_ERR_UNSUPPORTED only if the field of one chain (N L bytes) does not fit shared memory. */
int lqmc_sweep_det(lqmc_engine* e, int n_sweeps, const double* uniforms, uint64_t seed, int measure);
/* old_det of every chain after the last lqmc_sweep_det (what _update_step_det returns, lqmc.py:259): f64 [chain]. */
int lqmc_get_det(lqmc_engine* e, double* det_old);
/* The old_det the NEXT lqmc_sweep_det starts from.  `_update_step_det(old_det)` takes it as an argument and the reference's
 * loops carry it from sweep to sweep (lqmc.py:236-259, 264-270): det_old = f64 [chain], or NULL to carry the value the
 * previous lqmc_sweep_det left on the device (a loop split into several calls).  One-shot; without it a call starts from
 * det(M_up(0)) det(M_dn(0)) of the current field, as the reference's loops do before their first sweep. */
int lqmc_set_det(lqmc_engine* e, const double* det_old);

/* Per-proposal record of the last lqmc_sweep / lqmc_slice call (needs LQMC_TRACE): what the
 * reference exposes as self.ratio / self.acc and logs through _debug (lqmc.py:217-232,316-317).
 * acc: u8 [chain][sweep][step][site]; ratio: f64, same shape.  Either pointer may be NULL. */
int lqmc_get_trace(lqmc_engine* e, uint8_t* acc, double* ratio);

/* Measurement accumulators (measure_loop, lqmc.py:364-375): g_sum f64 [chain][2][N][N] = sum over
 * measured sweeps of the end-of-sweep G; obs_sum f64 [chain][3][N] = per-site sums of n_up, n_dn and
 * the per-configuration product n_up*n_dn; n_meas int64 [chain]; n_accepted int64
 * [chain] counts accepted flips since the last reset.  Any pointer may be NULL. */
int lqmc_get_measurements(lqmc_engine* e, double* g_sum, double* obs_sum, int64_t* n_meas,
                          int64_t* n_accepted);
int lqmc_reset_measurements(lqmc_engine* e);
/* Restore accumulators saved with lqmc_get_measurements (same layouts; any pointer may be NULL): the resume half of
 * checkpoint / resume.  The reference only caches finished beta points (multiprocessing.py:312-333, broken on
 * numpy >= 1.24: SURVEY.md section 5); the Markov state here is field + sweep counter + accumulators - the Philox
 * stream is a pure function of (seed, chain, sweep, proposal), so a resumed run repeats the uninterrupted one bit for bit. */
int lqmc_set_measurements(lqmc_engine* e, const double* g_sum, const double* obs_sum, const int64_t* n_meas,
                          const int64_t* n_accepted);

/* Raw device pointers, for plumbing that must stay on the device (torch.distributed all-reduce of the
 * accumulators over NCCL, device-side uniforms).  which: 0 field (int8 [chain][slice][NP], slice-major,
 * padded), 1 G (f64 [chain][2][NP][NP]), 2 g_sum (f64 [chain][2][N][N]), 3 obs_sum (f64 [chain][3][N]),
 * 4 n_meas (int64 [chain]), 5 n_accepted (int64 [chain]).  *n_bytes receives the allocation size. */
int lqmc_device_ptr(lqmc_engine* e, int which, void** ptr, uint64_t* n_bytes);

/* CTAs per Markov chain of the most recent sweep / slice / wrap launch: 1, or the size of the thread-block cluster one chain ran
 * on (large-lattice kernel at 16x16 with fewer chains than half the SMs: the reference's analogue is a fixed job split over more
 * worker processes, multiprocessing.py:260-267).  Results do not depend on it (bit-identical). */
int lqmc_get_cluster(lqmc_engine* e, int* ctas_per_chain);

/* Engine facts: padded size NP, global sweep counter, kernels launched so far, which kernel family
 * serves this size ("reg" = register-resident G, one CTA per chain; "l2" = G in HBM/L2 with delayed
 * rank-k updates). */
int lqmc_info(lqmc_engine* e, int* n_pad, int64_t* sweep_counter, int64_t* launches, char family[8]);
int lqmc_set_sweep_counter(lqmc_engine* e, int64_t counter);
/* Global index of this engine's chain 0 (rank * chains_per_rank under torchrun): the Philox stream of
 * chain c is keyed by chain_offset + c, so an N-GPU run equals the concatenation of 1-GPU runs. */
int lqmc_set_chain_offset(lqmc_engine* e, int64_t chain0);

/* The device Philox4x32-10 stream, evaluated on the host: fills out[n_sites * n_slices] with the
 * uniforms chain `chain` consumes in global sweep `sweep` under `seed` (visiting order). */
void lqmc_philox_uniforms(uint64_t seed, uint64_t chain, uint64_t sweep, int n_proposals, double* out);

/* Device self-test: the sweep divides a whole column by one denominator through a shared reciprocal
 * and two FMA corrections; this checks n_samples random (x, d) pairs, including all-ones / sparse
 * mantissas, for bit-equality with IEEE division (what np.divide does in lqmc.py:326-327). */
int lqmc_selftest_division(int device, uint64_t n_samples, uint64_t seed, uint64_t* mismatches);

const char* lqmc_last_error(void);
const char* lqmc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LQMC_B200_H */
