"""Chain drivers: the reference's process managers (lqmc/multiprocessing.py:16-341) with
"one OS process per chain" replaced by "one chain index in the GPU batch".

`ParallelProcessManager(procs=C)` runs C independent Markov chains of the same model as ONE
engine (one CTA per chain, no communication during the sweeps); `SerialProcessManager` runs one
chain per beta, each on its own engine / CUDA stream, ALL AT ONCE (`_run_batch`: beta is a per-chain
parameter, the kernels of the different points run concurrently on the GPU).  Public methods, job splitting
(`sweeps/procs` each, remainder to chain 0, multiprocessing.py:260-263) and the unweighted mean over
chains (`:265-267`) are the reference's.  Differences, all deliberate:

* seeding: the reference child seeds the legacy NumPy stream with its OS pid
  (multiprocessing.py:47), then draws the initial field and every uniform from it.  Here chain c
  uses `RandomState(seeds[c])` the same way (field first, then `N*L` uniforms per sweep), so a
  chain reproduces exactly what a reference process with that pid would compute; `seeds` defaults
  to `os.getpid() + c`.  `rng="philox"` switches the uniforms to the device stream.
* no Pipe: the reference deadlocks once the `(2,N,N)` result exceeds the 64 KiB pipe buffer
  (N >~ 64); results are read straight from the device accumulators.
* the `.npz` beta-scan cache (multiprocessing.py:312-341) stores a float array plus a done-mask
  instead of an object array holding None, which NumPy >= 1.24 rejects.

With torch.distributed initialised, `ParallelProcessManager` shards its chains over the ranks
(contiguous blocks) and all-reduces the per-chain means once at the end (NCCL on GPUs).
"""
import os
import time
import multiprocessing

import numpy as np
from scipy.linalg import expm

from .engine import SweepEngine
from .lqmc import LatticeQMC, _UNIFORM_CHUNK_BYTES


def timestr(seconds):
    mins, secs = divmod(seconds, 60)
    if mins >= 60:
        hours, mins = divmod(mins, 60)
        return f"{int(hours):0>2}:{int(mins):0>2} h"
    return f"{int(mins):0>2}:{int(secs):0>2} min"


def _dist():
    """(rank, world, module) of an initialised torch.distributed job, else (0, 1, None)."""
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return 0, 1, None
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size(), dist
    return 0, 1, None


def shard_range(total, rank, world):
    """Contiguous block of `total` chains owned by `rank`."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class LqmcProcess(LatticeQMC):
    """One chain slot.  Kept for API compatibility (multiprocessing.py:25-52): it is a `LatticeQMC`
    with an index, a shared progress counter and a result sink instead of an OS process."""

    def __init__(self, i, iters, pipe, *args, seed=None, **kwargs):
        kwargs.pop("log_lvl", None)
        # device Philox streams are keyed by (seed, chain offset, sweep): every job slot gets its own offset, or all points of
        # a scan would consume the very same uniforms and their statistical errors would be fully correlated
        kwargs.setdefault("chain_offset", i)
        super().__init__(*args, **kwargs, log_lvl=None)
        self.idx = i
        self.iters = iters
        self.pipe = pipe
        self.pid = os.getpid() + i if seed is None else seed
        self._alive = False
        self.result = None

    def is_alive(self):
        return self._alive

    def is_done(self):
        return not self._alive

    def start(self):
        self._alive = True
        self.run()
        self._alive = False

    def join(self):
        return None

    def terminate(self):
        self._alive = False

    def run(self):
        state = np.random.get_state()
        np.random.seed(self.pid)
        self.config.initialize()
        gf = self.run_lqmc()
        np.random.set_state(state)
        if self.iters is not None:
            self.iters[self.idx] = self.warm_sweeps + self.meas_sweeps
        self.result = gf
        if self.pipe is not None:
            self.pipe.append(gf)


class ProcessManager:

    def __init__(self, procs=None, **default_kwargs):
        if procs is None or procs == 0:
            n_procs = multiprocessing.cpu_count()
        elif procs < 0:
            n_procs = multiprocessing.cpu_count() + procs
        else:
            n_procs = procs
        self.max_procs = n_procs
        self.processes = list()
        self.lock = multiprocessing.Lock()
        self.iters = [0] * self.max_procs
        self.idx = 0
        self.total = 0
        self.result = None
        self.default_kwargs = default_kwargs
        self.var_kwargs = dict()
        self.t0 = 0

    @property
    def model(self):
        return self.default_kwargs["model"]

    def set_jobs(self, **kwargs):
        lengths = {len(v) for v in kwargs.values()}
        if len(lengths) != 1:
            raise ValueError("All variable lists must have the same length!")
        self.var_kwargs = kwargs
        self.idx = 0
        self.total = lengths.pop()
        self.result = [None for _ in range(self.total)]
        self.iters = [0] * self.total

    def __str__(self):
        return f"{self.__class__.__name__}(Vars: {list(self.var_kwargs)}, Jobs: {self.total}, Processes: {self.max_procs})"

    @property
    def all_done(self):
        return self.idx == self.total and not self.processes

    @property
    def jobs_running(self):
        return len(self.processes)

    @property
    def jobs_pending(self):
        return self.total - self.idx

    @property
    def jobs_done(self):
        return self.idx - self.jobs_running

    @property
    def free_processes(self):
        return self.max_procs - len(self.processes)

    @property
    def time(self):
        return time.time() - self.t0

    def get_result(self):
        return np.array(self.result)

    def get_progress(self):
        return self.jobs_done / self.total if self.total else 1.0

    def get_eta(self):
        p = self.get_progress()
        return (1 / p - 1) * self.time if p else 0.0

    def join(self):
        return None

    def terminate(self):
        self.processes.clear()

    def job_kwargs(self, idx):
        kwargs = self.default_kwargs.copy()
        for key, values in self.var_kwargs.items():
            kwargs[key] = values[idx]
        return kwargs

    def start_process(self):
        idx = self.idx
        self.idx += 1
        sink = []
        p = LqmcProcess(idx, self.iters, sink, **self.job_kwargs(idx))
        self.processes.append((p, sink))
        p.start()

    def end_process(self, item):
        p, sink = item
        self.result[p.idx] = np.array(sink[0])
        self.processes.remove(item)

    def handle_processes(self):
        for item in list(self.processes):
            if item[0].is_done():
                self.end_process(item)
        while self.jobs_pending and self.free_processes:
            self.start_process()
            self.end_process(self.processes[-1])

    def start_str(self, *args, **kwargs):
        return f"Starting {self}"

    def start(self):
        text = self.start_str()
        if text:
            print(text)
        return []

    def update_str(self, *args, **kwargs):
        return (f"Progress: {100 * self.get_progress():5.1f}%, eta: {timestr(self.get_eta())}"
                f" (Alive: {self.jobs_running}, Pending: {self.jobs_pending}, Done: {self.jobs_done})")

    def update(self, *args):
        print(f"\r{self.update_str():<80}", end="", flush=True)
        return args

    def end(self, *args):
        print()
        print(f"Total time: {timestr(self.time)}")

    def run(self, sleep=0.5):
        self.t0 = time.time()
        args = self.start()
        while not self.all_done:
            self.handle_processes()
            args = self.update(*args)
        self.end(args)


class ParallelProcessManager(ProcessManager):

    CORE_COUNT = multiprocessing.cpu_count()

    def __init__(self, model, beta, time_steps, warmup=300, det_mode=False, procs=None, *, seeds=None,
                 rng="numpy", mode="parity", arith="exact", device=None, seed=0, stab_every=0):
        super().__init__(procs, model=model, beta=beta, time_steps=time_steps, warmup=warmup, det_mode=det_mode)
        self.stab_every = stab_every
        self.seeds = list(seeds) if seeds is not None else [os.getpid() + c for c in range(self.max_procs)]
        if len(self.seeds) != self.max_procs:
            raise ValueError("need one seed per chain")
        self.rng, self.mode, self.arith, self.device, self.seed = rng, mode, arith, device, seed
        self.observables = None

    def set_jobs(self, sweeps):
        sweeplist = np.full(self.max_procs, sweeps / self.max_procs, dtype="int")
        sweeplist[0] += sweeps - np.sum(sweeplist)
        super().set_jobs(sweeps=sweeplist)

    def get_result(self):
        return np.sum(super().get_result(), axis=0) / self.max_procs

    def run(self, sleep=0.5):
        """All chains at once: warm-up, `sweeps/procs` measured sweeps each, the remainder on
        chain 0; then (under torch.distributed) one all-reduce of the per-chain means."""
        self.t0 = time.time()
        print(self.start_str())
        kw = self.default_kwargs
        model, lt, warm = kw["model"], kw["time_steps"], kw["warmup"]
        n = model.n_sites
        sweeplist = [int(s) for s in self.var_kwargs["sweeps"]]
        base = min(sweeplist)
        dtau = kw["beta"] / lt
        ham = model.ham_kinetic()
        lamb = np.arccosh(np.exp(model.u * dtau / 2.)) if model.u else 0
        exp_k, exp_k_inv = expm(-1 * dtau * ham), expm(dtau * ham)
        rank, world, dist = _dist()
        lo, hi = shard_range(self.max_procs, rank, world)
        device = self.device if self.device is not None else int(os.environ.get("LOCAL_RANK", 0))
        means = np.zeros((self.max_procs, 2, n, n))
        obs = np.zeros((self.max_procs, 3, n))
        if hi > lo:
            streams = [np.random.RandomState(self.seeds[c]) for c in range(lo, hi)]
            fields = np.stack([(2 * rs.randint(0, 2, size=(n, lt)) - 1).astype(np.int8) for rs in streams])

            def uniforms(k):
                if self.rng != "numpy":
                    return None
                return np.stack([rs.rand(k * lt * n).reshape(k, lt, n) for rs in streams])

            def batched(engine, count, measure, us=None):
                chunk = max(1, _UNIFORM_CHUNK_BYTES // (engine.n_chains * lt * n * 8))
                done = 0
                while done < count:
                    k = min(chunk, count - done)
                    step = engine.sweep_det if kw["det_mode"] else engine.sweep
                    step(k, us(k) if us else uniforms(k), seed=self.seed, measure=measure)
                    done += k

            with SweepEngine(exp_k, lamb, lt, n_chains=hi - lo, exp_k_inv=exp_k_inv, device=device, mode=self.mode,
                             arith=self.arith, chain_offset=lo, stab_every=self.stab_every) as eng:
                eng.set_field(fields)
                batched(eng, warm, False)
                batched(eng, base, True)
                m = eng.get_measurements()
                extra = sweeplist[0] - base
                if extra and lo == 0:
                    with SweepEngine(exp_k, lamb, lt, n_chains=1, exp_k_inv=exp_k_inv, device=device, mode=self.mode,
                                     arith=self.arith, stab_every=self.stab_every) as tail:
                        tail.set_field(eng.get_field()[:1])
                        tail.set_sweep_counter(warm + base)
                        batched(tail, extra, True, us=lambda k: None if self.rng != "numpy" else
                                streams[0].rand(k * lt * n).reshape(1, k, lt, n))
                        mt = tail.get_measurements()
                    m["g_sum"][0] += mt["g_sum"][0]
                    m["obs_sum"][0] += mt["obs_sum"][0]
                    m["n_meas"][0] += mt["n_meas"][0]
            counts = np.maximum(m["n_meas"], 1).astype(np.float64)
            means[lo:hi] = m["g_sum"] / counts[:, None, None, None]
            obs[lo:hi] = m["obs_sum"] / counts[:, None, None]
        if dist is not None:
            import torch
            use_cuda = dist.get_backend() == "nccl"
            buf = torch.from_numpy(np.concatenate([means.ravel(), obs.ravel()]))
            if use_cuda:
                buf = buf.cuda(device)
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            flat = buf.cpu().numpy()
            means = flat[:means.size].reshape(means.shape)
            obs = flat[means.size:].reshape(obs.shape)
        self.result = [means[c] for c in range(self.max_procs)]
        self.observables = dict(n_up=obs[:, 0], n_dn=obs[:, 1], docc=obs[:, 2])
        self.iters = [warm + s for s in sweeplist]
        self.idx = self.total
        self.end()

    @staticmethod
    def _frmt_items(items, delim, width):
        return delim.join(f"{item:>{width}}" for item in items)


class SerialProcessManager(ProcessManager):

    CORE_COUNT = multiprocessing.cpu_count()
    TMP_FILE = "tmp_data.npz"

    MAX_CONCURRENT = 128      # resident grids per device (hardware limit of concurrently running kernels)

    def __init__(self, model, time_steps, warmup=300, sweeps=2000, det_mode=False, procs=None, caching=True,
                 concurrent=None, **engine_kwargs):
        engine_kwargs.setdefault("trace", False)
        super().__init__(procs, model=model, time_steps=time_steps, warmup=warmup, sweeps=sweeps,
                         det_mode=det_mode, **engine_kwargs)
        self.caching = caching
        self._tmp_file = "tmp_gf_series.npz"
        # parameter points in flight at once on the GPU (the reference: `procs` OS processes, multiprocessing.py:55-75;
        # here the limit is the device's, not the host's core count)
        self.concurrent = self.MAX_CONCURRENT if concurrent is None else max(1, int(concurrent))
        self.observables = None

    def set_jobs(self, betas):
        super().set_jobs(beta=betas)

    def start_str(self, *args, **kwargs):
        return (super().start_str() + f"\nWarmup     ={self.default_kwargs['warmup']}"
                + f"\nMeasurement={self.default_kwargs['sweeps']}")

    def start(self):
        args = super().start()
        if self.caching and os.path.isfile(self._tmp_file):
            with np.load(self._tmp_file) as data:
                done, values = data["done"], data["data"]
            if len(done) == self.total:
                for i in range(self.total):
                    if not done[i]:
                        break
                    self.result[i] = values[i]
                    self.idx = i + 1
                print(f"Found temporary data. Continuing at job {self.idx}...")
        return args

    def _save_cache(self):
        if self.caching:
            done = np.array([r is not None for r in self.result])
            shape = next(r.shape for r in self.result if r is not None)
            values = np.stack([r if r is not None else np.zeros(shape) for r in self.result])
            np.savez(self._tmp_file, beta=np.asarray(self.var_kwargs["beta"]), data=values, done=done)

    def end_process(self, item):
        super().end_process(item)
        self._save_cache()

    def _run_batch(self, batch):
        """One engine (own CUDA stream) per parameter point of `batch`, all in flight at once: beta is a per-chain
        parameter (own `dtau`, `lamb`, `exp_k`), sweeps are submitted round-robin with `SweepEngine.sweep_submit` and
        the kernels of the different points run concurrently.  Every chain sees exactly the numbers its sequential
        `LqmcProcess.run` would (multiprocessing.py:45-52): stream seeded with its pid, field first, then `N*L`
        uniforms per sweep - so the results are bit-identical to running the points one after the other."""
        procs = []
        for idx in batch:
            p = LqmcProcess(idx, self.iters, None, **self.job_kwargs(idx))
            state = np.random.get_state()
            np.random.seed(p.pid)
            p.config.initialize()
            p._rs = np.random.RandomState()
            p._rs.set_state(np.random.get_state())
            np.random.set_state(state)
            p.engine.set_field(p.config.config[None])
            procs.append(p)

        def phase(count_of, measure):
            done = [0] * len(procs)
            budget = max(1, _UNIFORM_CHUNK_BYTES // len(procs))
            while True:
                busy = False
                for j, p in enumerate(procs):
                    left = count_of(p) - done[j]
                    if left <= 0:
                        continue
                    busy = True
                    per_sweep = p.time_steps * p.n_sites
                    k = min(left, max(1, budget // (per_sweep * 8)))
                    u = p._rs.rand(k * per_sweep).reshape(1, k, p.time_steps, p.n_sites) if p.rng == "numpy" else None
                    p.engine.sweep_submit(k, u, seed=p.seed, measure=measure)
                    done[j] += k
                if not busy:
                    break
                for p in procs:
                    p.engine.sync()

        phase(lambda p: p.warm_sweeps, False)
        for p in procs:
            p.engine.reset_measurements()
        phase(lambda p: p.meas_sweeps, True)
        for p in procs:
            m = p.engine.get_measurements()
            p.observables = p._observables(m)
            p.config.config[...] = p.engine.get_field()[0]
            p.result = m["g_sum"][0] / max(p.meas_sweeps, 1)
            self.result[p.idx] = np.array(p.result)
            self.observables[p.idx] = p.observables
            self.iters[p.idx] = p.warm_sweeps + p.meas_sweeps
            p.engine.close()

    def run(self, sleep=0.5):
        if self.default_kwargs.get("det_mode"):
            return super().run(sleep)                 # det-mode sweeps are synchronous engine calls: one point at a time
        self.t0 = time.time()
        args = self.start()
        self.observables = [None] * self.total
        while self.idx < self.total:
            batch = list(range(self.idx, min(self.total, self.idx + self.concurrent)))
            self._run_batch(batch)
            self.idx = batch[-1] + 1
            self._save_cache()
            args = self.update(*args)
        self.end(args)

    def delete_cache(self):
        if os.path.isfile(self._tmp_file):
            os.remove(self._tmp_file)

    def end(self, *args):
        super().end()
        self.delete_cache()
