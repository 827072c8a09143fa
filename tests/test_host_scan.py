"""Host logic of the beta scan (`SerialProcessManager._run_batch`, lqmc/multiprocessing.py:292-341 in the reference) with the
engine replaced by a recorder: what every chain is fed must be exactly what a reference process seeded with that pid would draw
(`np.random.seed(pid)`, the field from `Configuration.initialize`, then N*L uniforms per sweep - multiprocessing.py:45-52,
lqmc.py:317), whatever the submission order; the npz cache of finished points (multiprocessing.py:312-341) is written, resumed
from and removed.  No GPU, no CUDA library: the kernels are covered by the `-m gpu` tests."""
import os

import numpy as np
import pytest

import latticeqmc_b200 as lq
from latticeqmc_b200 import multiprocessing as lmp
from latticeqmc_b200 import lqmc as lmod


class FakeEngine:
    instances = []

    def __init__(self, exp_k, lamb, n_slices, n_chains=1, **kw):
        self.n_sites, self.n_slices, self.n_chains = exp_k.shape[0], n_slices, n_chains
        self.exp_k, self.lamb, self.kw = exp_k, lamb, kw
        self.field = None
        self.calls = []          # (n_sweeps, uniforms copy or None, measure)
        self.synced = 0
        self.closed = False
        FakeEngine.instances.append(self)

    def set_field(self, f):
        self.field = np.array(f, dtype=np.int8).reshape(self.n_chains, self.n_sites, self.n_slices)

    def get_field(self):
        return self.field.copy()

    def sweep_submit(self, n, uniforms=None, seed=0, measure=False):
        self.calls.append((n, None if uniforms is None else np.array(uniforms).copy(), bool(measure)))

    sweep = sweep_submit

    def sync(self):
        self.synced += 1

    def reset_measurements(self):
        self.calls.append(("reset",))

    def get_measurements(self):
        n, c = self.n_sites, self.n_chains
        meas = sum(x[0] for x in self.calls if len(x) == 3 and x[2])
        return dict(g_sum=np.full((c, 2, n, n), float(meas) * self.lamb), obs_sum=np.zeros((c, 3, n)),
                    n_meas=np.full(c, meas), n_accepted=np.zeros(c, dtype=np.int64))

    def close(self):
        self.closed = True

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_sweep_counter(self, c):
        self.counter = c


@pytest.fixture
def fake_engine(monkeypatch):
    FakeEngine.instances = []
    monkeypatch.setattr(lmod, "SweepEngine", FakeEngine)
    monkeypatch.setattr(lmp, "SweepEngine", FakeEngine)
    return FakeEngine


def _model():
    m = lq.HubbardModel(u=4, t=1)
    m.build_square(2)
    return m


def test_scan_feeds_every_chain_its_reference_stream(fake_engine, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(lmp, "_UNIFORM_CHUNK_BYTES", 4 * 3 * 20 * 4 * 8)       # forces several submission rounds
    betas = [0.5, 1.0, 2.0]
    n, lt, warm, sweeps = 4, 20, 5, 7
    mgr = lq.SerialProcessManager(_model(), lt, warm, sweeps, procs=2, caching=True)
    mgr.set_jobs(betas)
    mgr.run()
    assert len(fake_engine.instances) == 3 and all(e.closed for e in fake_engine.instances)
    pid = os.getpid()
    for j, eng in enumerate(fake_engine.instances):
        rs = np.random.RandomState(pid + j)
        field = (2 * rs.randint(0, 2, size=(n, lt)) - 1).astype(np.int8)
        assert np.array_equal(eng.field[0], field)
        sweeps_seen = [c for c in eng.calls if len(c) == 3]
        assert sum(c[0] for c in sweeps_seen if not c[2]) == warm and sum(c[0] for c in sweeps_seen if c[2]) == sweeps
        # warm-up calls come before the reset, measured ones after it
        kinds = ["reset" if len(c) == 1 else c[2] for c in eng.calls]
        assert kinds.index("reset") == sum(1 for c in sweeps_seen if not c[2])
        fed = np.concatenate([c[1].ravel() for c in sweeps_seen])
        assert np.array_equal(fed, rs.rand((warm + sweeps) * lt * n))
        # every point has its own dtau / lamb / exp_k
        dtau = betas[j] / lt
        assert np.isclose(eng.lamb, np.arccosh(np.exp(4 * dtau / 2.)))
    res = mgr.get_result()
    assert res.shape == (3, 2, n, n)
    assert np.allclose(res[:, 0, 0, 0], [e.lamb for e in fake_engine.instances])      # g_sum / sweeps of the recorder
    assert not os.path.exists("tmp_gf_series.npz")                                   # removed on success


def test_scan_resumes_from_the_cache(fake_engine, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    betas = [0.5, 1.0, 2.0, 4.0]
    n, lt = 4, 10
    done = np.array([True, True, False, False])
    data = np.zeros((4, 2, n, n))
    data[0] += 11.0
    data[1] += 22.0
    np.savez("tmp_gf_series.npz", beta=np.asarray(betas), data=data, done=done)
    mgr = lq.SerialProcessManager(_model(), lt, 1, 2, procs=2, caching=True, concurrent=1)
    mgr.set_jobs(betas)
    mgr.run()
    assert len(fake_engine.instances) == 2                                           # only the two unfinished points ran
    res = mgr.get_result()
    assert np.all(res[0] == 11.0) and np.all(res[1] == 22.0)
    pid = os.getpid()
    rs = np.random.RandomState(pid + 2)                                              # job index 2 keeps its own seed
    assert np.array_equal(fake_engine.instances[0].field[0], (2 * rs.randint(0, 2, size=(n, lt)) - 1).astype(np.int8))


def test_scan_rejects_ragged_jobs(fake_engine):
    pm = lmp.ProcessManager(procs=2)
    with pytest.raises(ValueError):
        pm.set_jobs(beta=[1.0, 2.0], sweeps=[10])


def test_parallel_manager_streams_and_job_split(fake_engine):
    """`ParallelProcessManager` (multiprocessing.py:252-289): chain c draws its field and then N*L uniforms per sweep from
    `RandomState(seeds[c])`; `sweeps / procs` measured sweeps each and the remainder on chain 0 (its own one-chain engine,
    continuing chain 0's stream and sweep counter); result = unweighted mean over chains."""
    n, lt, warm, sweeps, procs = 4, 20, 3, 8, 3
    seeds = [11, 22, 33]
    mgr = lq.ParallelProcessManager(_model(), 2.0, lt, warmup=warm, procs=procs, seeds=seeds)
    mgr.set_jobs(sweeps)
    mgr.run()
    main, tail = fake_engine.instances
    assert main.n_chains == 3 and tail.n_chains == 1 and tail.counter == warm + 2
    streams = [np.random.RandomState(s) for s in seeds]
    fields = np.stack([(2 * rs.randint(0, 2, size=(n, lt)) - 1).astype(np.int8) for rs in streams])
    assert np.array_equal(main.field, fields)
    fed = np.concatenate([c[1] for c in main.calls if len(c) == 3], axis=1)           # (chains, sweeps, L, N)
    assert fed.shape == (3, warm + 2, lt, n)
    for c in range(3):
        assert np.array_equal(fed[c].ravel(), streams[c].rand((warm + 2) * lt * n))
    extra = np.concatenate([c[1].ravel() for c in tail.calls if len(c) == 3])
    assert np.array_equal(extra, streams[0].rand(2 * lt * n))                          # 8 = 2 + 2 + 2, remainder 2 on chain 0
    gf = mgr.get_result()
    assert gf.shape == (2, n, n)
