"""ctypes binding of the C-ABI sweep engine (`include/lqmc_b200.h`) - the only compute path.

There is no CPU fallback: if `liblqmc_b200.so` is missing or no CUDA device is visible, creating a
`SweepEngine` raises.  (The NumPy restatement under `oracle/` is test infrastructure and is never
imported from here.)
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LQMC_B200_LIB") or os.path.join(_HERE, "liblqmc_b200.so")     # override: kernel experiments

MODE_PARITY = 0x0
MODE_PHYSICS = 0x1
ARITH_EXACT = 0x0
ARITH_FMA = 0x2
TRACE = 0x4

EXPORTS = (
    "lqmc_create", "lqmc_destroy", "lqmc_set_field", "lqmc_get_field", "lqmc_set_g", "lqmc_get_g",
    "lqmc_recompute", "lqmc_slice", "lqmc_wrap", "lqmc_sweep", "lqmc_sweep_async", "lqmc_sync",
    "lqmc_get_trace", "lqmc_get_measurements", "lqmc_reset_measurements", "lqmc_device_ptr", "lqmc_info",
    "lqmc_set_sweep_counter", "lqmc_set_chain_offset", "lqmc_philox_uniforms", "lqmc_last_error",
    "lqmc_version", "lqmc_selftest_division", "lqmc_recompute_stable", "lqmc_set_stabilization",
    "lqmc_set_measurements", "lqmc_sweep_det", "lqmc_get_det", "lqmc_set_det", "lqmc_sweep_submit", "lqmc_get_cluster",
)


class EngineError(RuntimeError):
    """A C-ABI call returned a non-zero status.  LQMC_ERR_INVALID maps to ValueError instead,
    mirroring the plain Python exceptions the reference raises on bad arguments."""


_lib = None


def load_library(path=None):
    """dlopen the engine and declare every prototype of include/lqmc_b200.h."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.isfile(path):
        raise EngineError(f"{path} not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"(nvcc, sm_100a); there is no CPU fallback")
    lib = ctypes.CDLL(path)
    c_dp = ctypes.POINTER(ctypes.c_double)
    vp = ctypes.c_void_p
    lib.lqmc_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                c_dp, c_dp, ctypes.c_double, c_dp, ctypes.c_uint32]
    lib.lqmc_destroy.argtypes = [vp]
    lib.lqmc_destroy.restype = None
    lib.lqmc_set_field.argtypes = [vp, vp]
    lib.lqmc_get_field.argtypes = [vp, vp]
    lib.lqmc_set_g.argtypes = [vp, vp]
    lib.lqmc_get_g.argtypes = [vp, vp]
    lib.lqmc_recompute.argtypes = [vp, ctypes.c_int]
    lib.lqmc_recompute_stable.argtypes = [vp, ctypes.c_int, ctypes.c_int]
    lib.lqmc_set_stabilization.argtypes = [vp, ctypes.c_int]
    lib.lqmc_slice.argtypes = [vp, ctypes.c_int, vp, ctypes.c_uint64]
    lib.lqmc_wrap.argtypes = [vp, ctypes.c_int]
    lib.lqmc_sweep.argtypes = [vp, ctypes.c_int, vp, ctypes.c_uint64, ctypes.c_int]
    lib.lqmc_sweep_det.argtypes = [vp, ctypes.c_int, vp, ctypes.c_uint64, ctypes.c_int]
    lib.lqmc_get_det.argtypes = [vp, vp]
    lib.lqmc_set_det.argtypes = [vp, vp]
    lib.lqmc_get_cluster.argtypes = [vp, ctypes.POINTER(ctypes.c_int)]
    lib.lqmc_sweep_submit.argtypes = [vp, ctypes.c_int, vp, ctypes.c_uint64, ctypes.c_int]
    lib.lqmc_sweep_async.argtypes = [vp, ctypes.c_int, vp, ctypes.c_uint64, ctypes.c_int, vp]
    lib.lqmc_sync.argtypes = [vp]
    lib.lqmc_get_trace.argtypes = [vp, vp, vp]
    lib.lqmc_get_measurements.argtypes = [vp, vp, vp, vp, vp]
    lib.lqmc_reset_measurements.argtypes = [vp]
    lib.lqmc_set_measurements.argtypes = [vp, vp, vp, vp, vp]
    lib.lqmc_device_ptr.argtypes = [vp, ctypes.c_int, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_uint64)]
    lib.lqmc_info.argtypes = [vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int64),
                              ctypes.POINTER(ctypes.c_int64), ctypes.c_char_p]
    lib.lqmc_set_sweep_counter.argtypes = [vp, ctypes.c_int64]
    lib.lqmc_set_chain_offset.argtypes = [vp, ctypes.c_int64]
    lib.lqmc_philox_uniforms.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, c_dp]
    lib.lqmc_philox_uniforms.restype = None
    lib.lqmc_selftest_division.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]
    lib.lqmc_last_error.restype = ctypes.c_char_p
    lib.lqmc_version.restype = ctypes.c_char_p
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is ctypes.c_int and name not in ("lqmc_destroy", "lqmc_philox_uniforms"):
            fn.restype = ctypes.c_int
    if path == LIB_PATH:
        _lib = lib
    return lib


def philox_uniforms(seed, chain, sweep, n_proposals):
    """Host evaluation of the device RNG stream (visiting order) - lets tests feed the oracle the
    very numbers a Philox-mode sweep consumed."""
    out = np.empty(n_proposals, dtype=np.float64)
    load_library().lqmc_philox_uniforms(seed, chain, sweep, n_proposals, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out


def selftest_division(n_samples=1 << 28, seed=1, device=0):
    """Number of (x, d) pairs for which the kernel's shared-reciprocal division differs from IEEE."""
    bad = ctypes.c_uint64()
    lib = load_library()
    rc = lib.lqmc_selftest_division(device, n_samples, seed, ctypes.byref(bad))
    if rc:
        raise EngineError(lib.lqmc_last_error().decode())
    return bad.value


def hs_constants(lamb):
    """The four exponentials the sweep needs, evaluated with the NumPy calls the reference makes:
    `np.exp(-1 * sigma * lamb * config[:, l])` on an int8 column for exp(-+lamb) (lqmc.py:149,154) and
    `np.exp(+-arg) - 1` with the scalar `arg = 2 * lamb * config[i, l]` (lqmc.py:313-323)."""
    col = np.array([-1, 1], dtype=np.int8)
    ev = np.exp(-1 * 1 * lamb * col)          # [exp(+lamb), exp(-lamb)]
    arg = 2 * lamb * col[1]
    return np.array([ev[0], ev[1], np.exp(+arg) - 1, np.exp(-arg) - 1], dtype=np.float64)


class SweepEngine:
    """`n_chains` independent Markov chains of one model on one CUDA device.

    Parameters mirror what `LatticeQMC.set_beta` caches (lqmc.py:93-117): `exp_k = expm(-dtau K)`,
    `lamb`.  `exp_k_inv` defaults to `np.linalg.inv(exp_k)`.
    """

    def __init__(self, exp_k, lamb, n_slices, n_chains=1, exp_k_inv=None, device=0, mode="parity",
                 arith="exact", trace=False, chain_offset=0, stab_every=0):
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        exp_k = np.ascontiguousarray(exp_k, dtype=np.float64)
        if exp_k.ndim != 2 or exp_k.shape[0] != exp_k.shape[1]:
            raise ValueError("exp_k must be a square matrix")
        if exp_k_inv is None:
            exp_k_inv = np.linalg.inv(exp_k)
        exp_k_inv = np.ascontiguousarray(exp_k_inv, dtype=np.float64)
        if mode not in ("parity", "physics"):
            raise ValueError("mode must be 'parity' or 'physics'")
        if arith not in ("exact", "fma"):
            raise ValueError("arith must be 'exact' or 'fma'")
        flags = (MODE_PHYSICS if mode == "physics" else MODE_PARITY) | (ARITH_FMA if arith == "fma" else ARITH_EXACT)
        if trace:
            flags |= TRACE
        self.n_sites = int(exp_k.shape[0])
        self.n_slices = int(n_slices)
        self.n_chains = int(n_chains)
        self.mode, self.arith, self.trace = mode, arith, bool(trace)
        self.lamb = float(lamb)
        import hashlib
        self._model_hash = hashlib.sha256(exp_k.tobytes()).hexdigest()[:16]          # checkpoint compatibility check
        self.device = int(device)
        hs = hs_constants(self.lamb)
        dp = ctypes.POINTER(ctypes.c_double)
        self._check(self._lib.lqmc_create(ctypes.byref(self._h), self.device, self.n_sites, self.n_slices,
                                          self.n_chains, exp_k.ctypes.data_as(dp), exp_k_inv.ctypes.data_as(dp),
                                          self.lamb, hs.ctypes.data_as(dp), flags))
        self._chain_offset = 0
        if chain_offset:
            self.set_chain_offset(chain_offset)
        self.stab_every = 0
        if stab_every:
            self.set_stabilization(stab_every)
        self._last_trace_shape = None

    # -- plumbing ------------------------------------------------------------------------------
    def _check(self, rc):
        if rc == 0:
            return
        msg = self._lib.lqmc_last_error().decode()
        if rc == 1:
            raise ValueError(msg)
        raise EngineError(f"lqmc_b200 error {rc}: {msg}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.lqmc_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- state ---------------------------------------------------------------------------------
    def set_field(self, field):
        """`field`: int8 `(n_chains, N, L)` (or `(N, L)` for one chain) - stacked
        `Configuration.config` arrays."""
        f = np.ascontiguousarray(field, dtype=np.int8).reshape(self.n_chains, self.n_sites, self.n_slices)
        self._check(self._lib.lqmc_set_field(self._h, f.ctypes.data))

    def get_field(self, out=None):
        """`out`: optional preallocated C-contiguous int8 `(n_chains, N, L)` array (e.g. a view of pinned memory)."""
        shape = (self.n_chains, self.n_sites, self.n_slices)
        if out is None:
            out = np.empty(shape, dtype=np.int8)
        elif out.dtype != np.int8 or out.shape != shape or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous int8 array of shape (n_chains, N, L)")
        self._check(self._lib.lqmc_get_field(self._h, out.ctypes.data))
        return out

    def set_g(self, g):
        g = np.ascontiguousarray(g, dtype=np.float64).reshape(self.n_chains, 2, self.n_sites, self.n_sites)
        self._check(self._lib.lqmc_set_g(self._h, g.ctypes.data))

    def get_g(self, out=None):
        """`(gf_up, gf_dn)` of every chain, `(n_chains, 2, N, N)`.  `out`: optional preallocated C-contiguous float64
        array of that shape; page-locked memory makes the device->host copy a straight DMA."""
        shape = (self.n_chains, 2, self.n_sites, self.n_sites)
        if out is None:
            out = np.empty(shape, dtype=np.float64)
        elif out.dtype != np.float64 or out.shape != shape or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array of shape (n_chains, 2, N, N)")
        self._check(self._lib.lqmc_get_g(self._h, out.ctypes.data))
        return out

    # -- phases --------------------------------------------------------------------------------
    def recompute(self, l0=0):
        self._check(self._lib.lqmc_recompute(self._h, int(l0)))

    def recompute_stable(self, l0=0, chunk=8):
        """QR/UDV-stabilised `G(l0) = inv(get_m(l0))`, `chunk` B factors per factorization."""
        self._check(self._lib.lqmc_recompute_stable(self._h, int(l0), int(chunk)))

    def set_stabilization(self, stab_every):
        """Physics mode: rebuild G through `recompute_stable` every `stab_every` slices of a sweep (0 = off)."""
        self._check(self._lib.lqmc_set_stabilization(self._h, int(stab_every)))
        self.stab_every = int(stab_every)

    def slice(self, l, uniforms=None, seed=0):
        ptr = None
        if uniforms is not None:
            u = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(self.n_chains, self.n_sites)
            ptr = u.ctypes.data
        self._check(self._lib.lqmc_slice(self._h, int(l), ptr, int(seed)))
        self._last_trace_shape = (self.n_chains, 1, 1, self.n_sites)

    def wrap(self, l):
        self._check(self._lib.lqmc_wrap(self._h, int(l)))

    def sweep(self, n_sweeps=1, uniforms=None, seed=0, measure=False):
        """`n_sweeps` x `LatticeQMC._update_step`.  `uniforms`: `(n_chains, n_sweeps, L, N)` in
        visiting order, or None for the device Philox stream."""
        ptr = None
        if uniforms is not None:
            u = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(self.n_chains, n_sweeps, self.n_slices, self.n_sites)
            ptr = u.ctypes.data
        self._check(self._lib.lqmc_sweep(self._h, int(n_sweeps), ptr, int(seed), int(bool(measure))))
        self._last_trace_shape = (self.n_chains, int(n_sweeps), self.n_slices, self.n_sites)

    def sweep_det(self, n_sweeps=1, uniforms=None, seed=0, measure=False, old_det=None):
        """`n_sweeps` x `LatticeQMC._update_step_det` (lqmc.py:236-259) as one of the reference's det-mode loops:
        `old_det` from `get_m(0, +-1)` at the start of the call (`old_det=None`), a caller-supplied value per chain, or
        `"carry"` = the value the previous call left on the device; `measure` adds `inv(get_m(0, +-1))` after every sweep
        (lqmc.py:293-297).  `uniforms` as in `sweep`.  Any N (N > 64: matrices in global memory, slow)."""
        if isinstance(old_det, str):
            if old_det != "carry":
                raise ValueError("old_det must be None, 'carry' or per-chain values")
            self._check(self._lib.lqmc_set_det(self._h, None))
        elif old_det is not None:
            d = np.ascontiguousarray(np.broadcast_to(np.asarray(old_det, dtype=np.float64), (self.n_chains,)))
            self._check(self._lib.lqmc_set_det(self._h, d.ctypes.data))
        ptr = None
        if uniforms is not None:
            u = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(self.n_chains, n_sweeps, self.n_slices, self.n_sites)
            ptr = u.ctypes.data
        self._check(self._lib.lqmc_sweep_det(self._h, int(n_sweeps), ptr, int(seed), int(bool(measure))))
        self._last_trace_shape = (self.n_chains, int(n_sweeps), self.n_slices, self.n_sites)

    def get_det(self):
        """`old_det` of every chain after the last `sweep_det` (the return value of `_update_step_det`, lqmc.py:259)."""
        out = np.empty(self.n_chains, dtype=np.float64)
        self._check(self._lib.lqmc_get_det(self._h, out.ctypes.data))
        return out

    def sweep_async(self, n_sweeps=1, d_uniforms=0, seed=0, measure=False, stream=0):
        """Device-resident variant: `d_uniforms` is a raw device pointer (0 = Philox), `stream` a raw
        cudaStream_t (0 = the engine's stream).  Returns immediately; call `sync()`."""
        self._check(self._lib.lqmc_sweep_async(self._h, int(n_sweeps), ctypes.c_void_p(d_uniforms or None), int(seed),
                                               int(bool(measure)), ctypes.c_void_p(stream or None)))
        self._last_trace_shape = (self.n_chains, int(n_sweeps), self.n_slices, self.n_sites)

    def sweep_submit(self, n_sweeps=1, uniforms=None, seed=0, measure=False):
        """`sweep` without the final wait: queued on the engine's stream, returns at once (`sync()` waits).  Several
        engines driven this way run concurrently on one GPU (beta scans)."""
        ptr = None
        if uniforms is not None:
            u = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(self.n_chains, n_sweeps, self.n_slices, self.n_sites)
            ptr = u.ctypes.data
        self._check(self._lib.lqmc_sweep_submit(self._h, int(n_sweeps), ptr, int(seed), int(bool(measure))))
        self._last_trace_shape = (self.n_chains, int(n_sweeps), self.n_slices, self.n_sites)

    def sync(self):
        self._check(self._lib.lqmc_sync(self._h))

    def get_trace(self):
        """`(acc bool, ratio float64)` of the last sweep/slice call, shape `(chains, sweeps, steps, N)`."""
        shape = self._last_trace_shape
        if shape is None:
            raise ValueError("no sweep or slice has run yet")
        acc = np.empty(shape, dtype=np.uint8)
        ratio = np.empty(shape, dtype=np.float64)
        self._check(self._lib.lqmc_get_trace(self._h, acc.ctypes.data, ratio.ctypes.data))
        return acc.astype(bool), ratio

    # -- measurements --------------------------------------------------------------------------
    def get_measurements(self):
        """dict with `g_sum (C,2,N,N)`, `obs_sum (C,3,N)`, `n_meas (C,)`, `n_accepted (C,)`."""
        c, n = self.n_chains, self.n_sites
        g_sum = np.empty((c, 2, n, n), dtype=np.float64)
        obs = np.empty((c, 3, n), dtype=np.float64)
        n_meas = np.empty(c, dtype=np.int64)
        n_acc = np.empty(c, dtype=np.int64)
        self._check(self._lib.lqmc_get_measurements(self._h, g_sum.ctypes.data, obs.ctypes.data, n_meas.ctypes.data,
                                                    n_acc.ctypes.data))
        return dict(g_sum=g_sum, obs_sum=obs, n_meas=n_meas, n_accepted=n_acc)

    def reset_measurements(self):
        self._check(self._lib.lqmc_reset_measurements(self._h))

    def set_measurements(self, m):
        """Restore accumulators from a `get_measurements()` dict."""
        c, n = self.n_chains, self.n_sites
        g_sum = np.ascontiguousarray(m["g_sum"], dtype=np.float64).reshape(c, 2, n, n)
        obs = np.ascontiguousarray(m["obs_sum"], dtype=np.float64).reshape(c, 3, n)
        n_meas = np.ascontiguousarray(m["n_meas"], dtype=np.int64).reshape(c)
        n_acc = np.ascontiguousarray(m["n_accepted"], dtype=np.int64).reshape(c)
        self._check(self._lib.lqmc_set_measurements(self._h, g_sum.ctypes.data, obs.ctypes.data, n_meas.ctypes.data,
                                                    n_acc.ctypes.data))

    # -- checkpoint / resume -------------------------------------------------------------------
    @staticmethod
    def _checkpoint_path(path):
        path = str(path)
        return path if path.endswith(".npz") else path + ".npz"          # np.savez appends the suffix: do the same on load

    def save_checkpoint(self, path, seed=0, numpy_rng=False):
        """Markov state of every chain as one `.npz`: HS field (int8), global sweep counter, chain offset, RNG seed and
        the measurement accumulators, plus what identifies the simulation (lamb, a hash of exp_k, mode, arithmetic,
        stab_every).  With the device Philox stream (a pure function of seed / chain / sweep / proposal) this is
        everything: G is rebuilt from the field at the next sweep start.  `numpy_rng=True` also stores the state of NumPy's
        global MT19937 stream, which is what feeds the sweeps in the drop-in `rng="numpy"` mode."""
        m = self.get_measurements()
        info = self.info()
        extra = {}
        if numpy_rng:
            st = np.random.get_state()
            extra = dict(np_rng_keys=st[1], np_rng_pos=np.int64(st[2]), np_rng_has_gauss=np.int64(st[3]), np_rng_gauss=np.float64(st[4]))
        np.savez_compressed(self._checkpoint_path(path), field=self.get_field(), sweep_counter=np.int64(info["sweep_counter"]),
                            chain_offset=np.int64(self._chain_offset), seed=np.uint64(seed), n_sites=self.n_sites,
                            n_slices=self.n_slices, n_chains=self.n_chains, mode=self.mode, arith=self.arith,
                            stab_every=self.stab_every, lamb=np.float64(self.lamb), model_hash=self._model_hash,
                            g_sum=m["g_sum"], obs_sum=m["obs_sum"], n_meas=m["n_meas"], n_accepted=m["n_accepted"], **extra)

    def load_checkpoint(self, path):
        """Restore a `save_checkpoint` file into this engine.  The file must come from the same simulation: lattice size,
        slices, chain count, lamb, exp_k, mode, arithmetic and stab_every are checked (ValueError otherwise).  Restores
        NumPy's global stream if it was stored.  Returns the RNG seed stored with it."""
        with np.load(self._checkpoint_path(path)) as z:
            if (int(z["n_sites"]), int(z["n_slices"]), int(z["n_chains"])) != (self.n_sites, self.n_slices, self.n_chains):
                raise ValueError("checkpoint does not match this engine's (n_sites, n_slices, n_chains)")
            if "model_hash" in z.files:           # files written before these fields existed carry no identity to check
                same = (str(z["model_hash"]) == self._model_hash and float(z["lamb"]) == self.lamb and str(z["mode"]) == self.mode
                        and str(z["arith"]) == self.arith and int(z["stab_every"]) == self.stab_every)
                if not same:
                    raise ValueError("checkpoint was written by a different simulation (exp_k / lamb / mode / arith / stab_every differ)")
            self.set_field(z["field"])
            self.set_sweep_counter(int(z["sweep_counter"]))
            self.set_chain_offset(int(z["chain_offset"]))
            self.set_measurements(dict(g_sum=z["g_sum"], obs_sum=z["obs_sum"], n_meas=z["n_meas"], n_accepted=z["n_accepted"]))
            if "np_rng_keys" in z.files:
                np.random.set_state(("MT19937", z["np_rng_keys"], int(z["np_rng_pos"]), int(z["np_rng_has_gauss"]), float(z["np_rng_gauss"])))
            return int(z["seed"])

    def device_ptr(self, which):
        """Raw `(pointer, n_bytes)` of an engine buffer (see lqmc_device_ptr)."""
        ptr = ctypes.c_void_p()
        nbytes = ctypes.c_uint64()
        self._check(self._lib.lqmc_device_ptr(self._h, int(which), ctypes.byref(ptr), ctypes.byref(nbytes)))
        return ptr.value, nbytes.value

    def info(self):
        n_pad = ctypes.c_int()
        counter = ctypes.c_int64()
        launches = ctypes.c_int64()
        fam = ctypes.create_string_buffer(8)
        self._check(self._lib.lqmc_info(self._h, ctypes.byref(n_pad), ctypes.byref(counter), ctypes.byref(launches), fam))
        return dict(n_pad=n_pad.value, sweep_counter=counter.value, launches=launches.value, family=fam.value.decode())

    def ctas_per_chain(self):
        """CTAs one chain ran on in the most recent launch (1, or the thread-block cluster size chosen for few chains)."""
        v = ctypes.c_int()
        self._check(self._lib.lqmc_get_cluster(self._h, ctypes.byref(v)))
        return v.value

    def set_sweep_counter(self, counter):
        self._check(self._lib.lqmc_set_sweep_counter(self._h, int(counter)))

    def set_chain_offset(self, chain0):
        self._check(self._lib.lqmc_set_chain_offset(self._h, int(chain0)))
        self._chain_offset = int(chain0)
