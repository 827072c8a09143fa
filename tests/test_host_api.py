"""Host-side drop-in API: lattice / model / configuration behave like the reference's
(`/root/reference/lqmc/{lattice,hubbard,configuration,tools}.py`)."""
import numpy as np
import pytest

from oracle import sweep_oracle as so


def test_kinetic_matches_reference_builder(golden):
    from latticeqmc_b200 import HubbardModel
    g = golden("lattice_ham")
    for size in (2, 3, 4, 5, 8):
        m = HubbardModel(u=4, t=1)
        m.build_square(size)
        assert np.array_equal(m.ham_kinetic(), g[f"square{size}"]), size
    m = HubbardModel(u=8, t=1)
    m.build(64)
    assert np.array_equal(m.ham_kinetic(), g["ring64"])
    m = HubbardModel(u=0, t=1, mu=0)
    m.build(10, cycling=None)
    assert np.array_equal(m.ham_kinetic(), g["open10"])
    m = HubbardModel(u=6, t=1)
    m.build(3, 2, cycling=0)
    assert np.array_equal(m.ham_kinetic(), g["rect3x2_c0"])


@pytest.mark.parametrize("size", [6, 11, 16, 24])
def test_square_lattices_the_reference_cannot_build(size):
    """the reference fails at L=6,11,16; the drop-in must give the ideal K."""
    from latticeqmc_b200 import HubbardModel
    m = HubbardModel(u=4, t=1)
    m.build_square(size)
    ham = m.ham_kinetic()
    assert np.array_equal(ham, so.ideal_square_kinetic(size, 1.0, 2.0))
    assert all(len(set(m.lattice.nearest_neighbours(i))) == 4 for i in range(size * size))


def test_configuration_stream_and_ops():
    """configuration.py:123-136 + SURVEY.md B.8 reference values of the legacy stream."""
    from latticeqmc_b200 import Configuration
    np.random.seed(12345)
    assert np.allclose(np.random.rand(3), [0.92961609, 0.31637555, 0.18391881])
    np.random.seed(12345)
    c = Configuration(2, 3)
    assert c.config.dtype == np.int8 and c.config.shape == (2, 3)
    assert np.array_equal(c.config, [[-1, 1, 1], [1, -1, 1]])
    assert np.random.rand() == 0.2045602785530397
    c.update(0, 1)
    assert c.get(0, 1) == -1 and c[0, 1] == -1
    d = c.copy()
    assert d == c
    d.update(1, 1)
    assert not (d == c)
    assert c.mean() == np.mean(c.config) and c.var() == np.var(c.config)
    assert str(c).count("\n") == 2


def test_tools_observables():
    from latticeqmc_b200 import filling, local_moment, local_gf, fermi_fct, compute_pole_gf_tau
    g = np.array([[0.3, 0.1], [0.1, 0.6]])
    assert np.allclose(filling(g), [0.7, 0.4])
    assert np.allclose(local_moment(g, g), 2 * np.array([0.7, 0.4]) - 2 * np.array([0.49, 0.16]))
    assert np.allclose(local_gf(g), [0.3, 0.6])
    assert np.isclose(fermi_fct(0.0, 3.0), 0.5)
    ham = so.ideal_ring_kinetic(6, 1.0, 0.0)
    tau, gf = compute_pole_gf_tau(ham, 2.0)
    from scipy.linalg import expm
    assert np.allclose(-gf[:, :, 0], np.linalg.inv(np.eye(6) + expm(-2.0 * ham)), atol=1e-12)


def test_chain_statistics_error_bars():
    """`chain_statistics`: means and chain-to-chain standard errors of the device accumulators."""
    import numpy as np
    from latticeqmc_b200.tools import chain_statistics
    rs = np.random.RandomState(0)
    chains, n, meas = 400, 6, 25
    # per-chain means scatter around (0.5, 0.5, 0.1) with known spreads
    base = np.stack([0.5 + 0.02 * rs.randn(chains), 0.5 + 0.02 * rs.randn(chains), 0.1 + 0.01 * rs.randn(chains)], axis=1)
    obs = base[:, :, None] * meas * np.ones((1, 1, n))
    m = dict(obs_sum=obs, n_meas=np.full(chains, meas), n_accepted=np.zeros(chains, dtype=np.int64))
    st = chain_statistics(m)
    assert st["n_chains"] == chains and st["n_meas"] == chains * meas
    assert abs(st["n_up"][0] - 0.5) < 4 * st["n_up"][1] and abs(st["n_up"][1] - 0.02 / np.sqrt(chains)) < 3e-4
    assert abs(st["docc"][0] - 0.1) < 4 * st["docc"][1]
    mom = base[:, 0] + base[:, 1] - 2 * base[:, 2]
    assert abs(st["moment"][0] - mom.mean()) < 1e-12
    assert abs(st["moment"][1] - mom.std(ddof=1) / np.sqrt(chains)) < 1e-12        # jackknife of a mean = standard error
    # chains without measurements are ignored
    m["n_meas"][:10] = 0
    assert chain_statistics(m)["n_chains"] == chains - 10


def test_scan_jobs_get_distinct_philox_streams():
    """Every job slot of a process manager keys the device Philox stream with its own chain offset (all points of a
    beta scan used to consume identical uniforms).  Host-side check: no engine is created."""
    from latticeqmc_b200 import HubbardModel
    from latticeqmc_b200.multiprocessing import LqmcProcess
    model = HubbardModel(u=4, t=1)
    model.build_square(2)
    procs = [LqmcProcess(i, None, None, model, 2.0, 20, warmup=0, sweeps=0, rng="philox") for i in range(4)]
    assert [p.chain_offset for p in procs] == [0, 1, 2, 3]
    assert LqmcProcess(2, None, None, model, 2.0, 20, warmup=0, sweeps=0, chain_offset=17).chain_offset == 17
    assert procs[0].last_trace is None
    big = LqmcProcess(0, None, None, model, 2.0, 20, warmup=0, sweeps=10**6)
    assert procs[0].trace is True and big.trace is False          # per-proposal record only while it stays small
