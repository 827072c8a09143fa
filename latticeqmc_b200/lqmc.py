"""`LatticeQMC`: the reference's sampler class (lqmc/lqmc.py:14-408) with the sweep delegated to
the B200 engine.

Same constructor, attributes and methods as the reference; `_update_step`, `warmup_loop`,
`measure_loop`, `run_lqmc` and `run` keep their meaning and return types.  What changes is where
the work happens: a sweep is one kernel launch through the C ABI (`engine.SweepEngine`), and the
warm-up / measurement loops hand whole batches of sweeps to the device instead of iterating in
Python.

RNG contract (lqmc.py:309-317).  With `rng="numpy"` (default) every sweep consumes
exactly `N*L` numbers from the global legacy NumPy stream - `np.random.rand(N*L)` yields the same
numbers as N*L successive `np.random.rand()` calls - so a seeded drop-in run walks through the
same accept/reject sequence as the reference.  `rng="philox"` uses the device counter-based stream
instead (no host traffic).

`det_mode=True` (the reference's validation sampler, lqmc.py:236-299: every proposal rebuilds `get_m(l, +-1)` and
takes two determinants) runs on the device too (`SweepEngine.sweep_det`, `csrc/sweep_det.cuh`).  One
`warmup_loop_det` / `measure_loop_det` call is one engine call, so `old_det` is initialised from `get_m(0, +-1)` at the
start of each loop and carried through it exactly as in the reference (loops longer than 256 MiB of uniforms are split,
and `old_det` is re-initialised from the field at the split).
"""
import time

import numpy as np
from scipy.linalg import expm

from .configuration import Configuration
from .engine import SweepEngine
from .logging import get_logger, DEBUG

_UNIFORM_CHUNK_BYTES = 256 << 20


class LatticeQMC:

    def __init__(self, model, beta, time_steps, warmup=300, sweeps=2000, det_mode=False, log_lvl=DEBUG,
                 *, mode="parity", arith="exact", rng="numpy", seed=0, device=0, trace=None, stab_every=0, chain_offset=0):
        if log_lvl is not None:
            self.logger = get_logger()
            self.logger.setLevel(log_lvl)
            self._log_debug("INIT")
        else:
            self.logger = None
        if rng not in ("numpy", "philox"):
            raise ValueError("rng must be 'numpy' or 'philox'")
        self.model = model
        self.n_sites = model.n_sites
        self.time_steps = time_steps
        self.config = Configuration(self.n_sites, time_steps)
        self.warm_sweeps = warmup
        self.meas_sweeps = sweeps

        self.det_mode = det_mode
        self.status = ""
        self.it = 0
        self.ratio = 0.0
        self.acc = False

        self.ham_kin = self.model.ham_kinetic()
        self.beta = 0.
        self.dtau = 0.
        self.lamb = 0.
        self.exp_k = None

        # engine options (additions; reference defaults unchanged)
        # trace=True keeps the per-proposal (ratio, acc) of the last call in `last_trace` and refreshes `self.ratio` / `self.acc`
        # (what the reference sets per proposal and `_debug` logs, lqmc.py:217-232,316-317).  It costs 9 bytes per proposal on the
        # device, copied back after every call, so the default (None) turns it on only while a whole loop's record stays under
        # 16 MB - small drop-in runs behave like the reference, production-size runs pay nothing.  chain_offset keys the device
        # Philox stream (rng="philox"): distinct jobs of a scan must not share one (seed, chain, sweep) stream.
        if trace is None:
            trace = 9 * self.n_sites * time_steps * max(warmup, sweeps, 1) <= (16 << 20)
        self.mode, self.arith, self.rng, self.seed, self.device, self.trace = mode, arith, rng, seed, device, bool(trace)
        self.chain_offset = chain_offset
        self.last_trace = None
        self._sweeps_done = 0               # Philox sweep counter carried across engine re-creation (set_beta)
        self.stab_every = stab_every        # physics mode: QR/UDV-stabilised G every this many slices (0 = off)
        self._engine = None

        self._log_debug(f"u=          {self.model.u}")
        self._log_debug(f"t=          {self.model.t}")
        self._log_debug(f"mu=         {self.model.mu}")
        self._log_debug(f"sites=      {self.n_sites}")
        self._log_debug(f"time_steps= {self.time_steps}")
        self._log_debug(f"det_mode=   {self.det_mode}")
        self._log_info(f"Warmup=     {self.warm_sweeps}")
        self._log_info(f"Measurement={self.meas_sweeps}")
        self._log_debug("END INIT")

        self.set_beta(beta)

    # ------------------------------------------------------------------ logging helpers
    def _log_info(self, msg, *args, **kwargs):
        if self.logger is not None:
            self.logger.info(msg, *args, **kwargs)

    def _log_debug(self, msg, *args, **kwargs):
        if self.logger is not None:
            self.logger.debug(msg, *args, **kwargs)

    def _log_warning(self, msg, *args, **kwargs):
        if self.logger is not None:
            self.logger.warning(msg, *args, **kwargs)

    # ------------------------------------------------------------------ setup (lqmc.py:93-128)
    def set_beta(self, beta):
        self._log_debug("SETUP")
        self.dtau = beta / self.time_steps
        self.beta = beta
        self.lamb = np.arccosh(np.exp(self.model.u * self.dtau / 2.)) if self.model.u else 0
        self.exp_k = expm(-1 * self.dtau * self.ham_kin)
        self.exp_k_inv = expm(self.dtau * self.ham_kin)
        if self._engine is not None:
            self._engine.close()
            self._engine = None
        self._log_debug(f"beta=       {self.beta}")
        self._log_debug(f"dtau=       {self.dtau}")
        self._log_debug(f"lambda=     {self.lamb}")
        check_val = self.model.u * self.model.t * self.dtau ** 2
        if check_val < 0.1:
            self._log_info(f"Check-value {check_val:.2} is smaller than 0.1!")
        else:
            self._log_warning(f"Check-value {check_val:.2} should be smaller than 0.1!")
        self._log_debug("END SETUP")

    def set_temperature(self, temp):
        self.set_beta(1 / temp)

    @property
    def engine(self):
        """The device engine for the current beta (created on first use; raises without CUDA)."""
        if self._engine is None:
            self._engine = SweepEngine(self.exp_k, self.lamb, self.time_steps, n_chains=1, exp_k_inv=self.exp_k_inv,
                                       device=self.device, mode=self.mode, arith=self.arith, trace=self.trace,
                                       stab_every=self.stab_every, chain_offset=self.chain_offset)
            if self._sweeps_done:
                self._engine.set_sweep_counter(self._sweeps_done)       # a new beta continues the stream, it does not replay it
        return self._engine

    # ------------------------------------------------------------------ inspection helpers (host, O(N^2))
    def get_exp_v(self, l, sigma):
        """Dense `exp(V_sigma(l))` (lqmc.py:132-154); the kernels use it as a row/column scale."""
        return np.diagflat(np.exp(-1 * sigma * self.lamb * self.config[:, l]))

    def get_m(self, l0, sigma):
        """`M_sigma(l0) = I + prod B_l` in the cyclic order of lqmc.py:175-185 (host helper for
        inspection; the sweep recomputes on the device, see `SweepEngine.recompute`)."""
        lt = self.time_steps
        order = [(l0 % lt - 1 - m) % lt for m in range(lt)]
        prod = 1
        for l in order:
            prod = np.dot(prod, np.dot(self.exp_k, self.get_exp_v(l, sigma)))
        return np.eye(self.n_sites) + prod

    def iter_sweeps(self, n, console_updates=200):
        """Progress generator (lqmc.py:187-215)."""
        self._log_debug(self.status.upper())
        every = max(1, int(n / console_updates)) if console_updates else n + 1
        for it in range(n):
            if (it + 1) % every == 0:
                line = f"{self.status} Sweep {it + 1} ({100 * (it + 1) / n:.1f}%)"
                line += f" [Mean: {self.config.mean():5.2f}, Var: {self.config.var():5.2f}]"
                print("\r" + line, end="", flush=True)
            self.it = it
            yield it
        print(f"\r{self.status} Sweep {n} (100.0%)")
        self._log_debug("END " + self.status.upper())

    def _debug(self, i, l):
        """One log line per proposal in the reference (lqmc.py:217-232).  The engine records the
        same information in its device trace buffer; this formats one entry of it."""
        self._log_debug(f"{self.status} {self.it + 1} -- {l:>2} {i:>2} -- {self.ratio:.1f} ({self.acc})"
                        f" -- {self.config.mean():.3f} {self.config.var():.3f}")

    # ------------------------------------------------------------------ the hot path
    def _draw_uniforms(self, n_sweeps):
        if self.rng != "numpy":
            return None
        return np.random.rand(n_sweeps * self.time_steps * self.n_sites).reshape(1, n_sweeps, self.time_steps, self.n_sites)

    def _run_sweeps(self, n_sweeps, measure, det=False, old_det=None):
        eng = self.engine
        eng.set_field(self.config.config[None])
        per_sweep = self.time_steps * self.n_sites * 8
        chunk = max(1, _UNIFORM_CHUNK_BYTES // per_sweep)
        done = 0
        while done < n_sweeps:
            k = min(chunk, n_sweeps - done)
            if det:
                # the first chunk starts from the caller's old_det (or the field's); later chunks carry the device value on
                eng.sweep_det(k, self._draw_uniforms(k), seed=self.seed, measure=measure,
                              old_det=(old_det if done == 0 else "carry"))
            else:
                eng.sweep(k, self._draw_uniforms(k), seed=self.seed, measure=measure)
            done += k
        self._sweeps_done += n_sweeps
        self.config.config[...] = eng.get_field()[0]
        if self.trace and n_sweeps:
            acc, ratio = eng.get_trace()
            self.ratio = float(ratio[0, -1, -1, -1])
            self.acc = bool(acc[0, -1, -1, -1])
            self.last_trace = (acc[0], ratio[0])

    def _update_step(self):
        """One sweep over all `N*L` HS spins on the device (lqmc.py:301-347).  Mutates
        `self.config.config` in place, returns fresh `(gf_up, gf_dn)` arrays."""
        self._run_sweeps(1, measure=False)
        g = self.engine.get_g()[0]
        return g[0].copy(), g[1].copy()

    def warmup_loop(self):
        self.status = "Warmup"
        self._log_debug(self.status.upper())
        self._run_sweeps(self.warm_sweeps, measure=False)
        self.it = max(self.warm_sweeps - 1, 0)
        self._log_debug("END " + self.status.upper())

    def measure_loop(self):
        """Mean of the end-of-sweep G over `meas_sweeps` sweeps, `(2, N, N)` (lqmc.py:356-375).  The
        accumulation `gf_total += gf` happens on the device, in the same order."""
        self.status = "Measurement"
        self._log_debug(self.status.upper())
        eng = self.engine
        eng.reset_measurements()
        self._run_sweeps(self.meas_sweeps, measure=True)
        self.it = max(self.meas_sweeps - 1, 0)
        m = eng.get_measurements()
        self.observables = self._observables(m)
        self._log_debug("END " + self.status.upper())
        return m["g_sum"][0] / self.meas_sweeps

    @staticmethod
    def _observables(m):
        n = np.maximum(m["n_meas"], 1)[:, None]
        return dict(n_up=m["obs_sum"][:, 0] / n, n_dn=m["obs_sum"][:, 1] / n, docc=m["obs_sum"][:, 2] / n,
                    n_meas=m["n_meas"].copy(), n_accepted=m["n_accepted"].copy())

    # ------------------------------------------------------------------ det mode (lqmc.py:236-299)
    def _update_step_det(self, old_det=None):
        """One det-mode sweep on the device (lqmc.py:236-259); returns the new `old_det`.  `old_det` is the value the ratios
        of this sweep are taken against - the reference carries it from sweep to sweep (lqmc.py:264-270); `None` starts from
        `det(get_m(0, +1)) * det(get_m(0, -1))` of the current field (what the reference's loops compute before their first
        sweep)."""
        self._run_sweeps(1, measure=False, det=True, old_det=old_det)
        return float(self.engine.get_det()[0])

    def warmup_loop_det(self):
        self.status = "Warmup"
        self._log_debug(self.status.upper())
        self._run_sweeps(self.warm_sweeps, measure=False, det=True)
        self.it = max(self.warm_sweeps - 1, 0)
        self._log_debug("END " + self.status.upper())

    def measure_loop_det(self):
        """Mean of `inv(get_m(0, +-1))` after each of `meas_sweeps` det-mode sweeps, `(2, N, N)` (lqmc.py:272-299)."""
        self.status = "Measurement"
        self._log_debug(self.status.upper())
        eng = self.engine
        eng.reset_measurements()
        self._run_sweeps(self.meas_sweeps, measure=True, det=True)
        self.it = max(self.meas_sweeps - 1, 0)
        m = eng.get_measurements()
        self.observables = self._observables(m)
        self._log_debug("END " + self.status.upper())
        return m["g_sum"][0] / self.meas_sweeps

    def run_lqmc(self):
        if self.det_mode:
            self.warmup_loop_det()
            return self.measure_loop_det()
        self.warmup_loop()
        return self.measure_loop()

    def run(self):
        t0 = time.time()
        gf_tau = self.run_lqmc()
        mins, secs = divmod(time.time() - t0, 60)
        self._log_info(f"Total time: {int(mins):0>2}:{int(secs):0>2} min")
        return gf_tau
