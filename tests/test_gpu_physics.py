"""Physics mode (textbook DQMC) on the GPU: against the physics oracle on the same
uniforms, and against exact diagonalisation within statistical error (BASELINE.json north_star, second
correctness gate)."""
import numpy as np
import pytest
from scipy.linalg import expm

from oracle import ed
from oracle import sweep_oracle as so

pytestmark = pytest.mark.gpu


def _setup(u, beta, lt, mu=0.0):
    ham = so.ideal_square_kinetic(2, 1.0, mu)          # diag = -mu; mu = 0 <=> true half filling (H6)
    dtau, lamb, exp_k = so.set_beta_constants(ham, u, beta, lt)
    return ham, lamb, exp_k, expm(dtau * ham)


def test_physics_sweep_matches_oracle():
    from latticeqmc_b200 import SweepEngine
    ham, lamb, exp_k, exp_k_inv = _setup(4.0, 2.0, 20)
    n, lt = 4, 20
    fields = np.stack([so.initial_field(n, lt, 40 + c) for c in range(3)])
    uni = np.random.RandomState(5).rand(3, 2, lt, n)
    with SweepEngine(exp_k, lamb, lt, n_chains=3, exp_k_inv=exp_k_inv, mode="physics", trace=True) as eng:
        eng.set_field(fields)
        eng.sweep(2, uni)
        acc, ratio = eng.get_trace()
        gg = eng.get_g()
    for c in range(3):
        h = fields[c].copy()
        for s in range(2):
            gu, gd, r, a = so.physics_sweep(h, exp_k, exp_k_inv, lamb, uni[c, s])
            assert np.array_equal(a, acc[c, s])
            # the oracle recomputes G(L-1) through QR/UDV, the engine through the plain product: ~1e-8 apart
            assert np.allclose(r, ratio[c, s], rtol=1e-6, atol=1e-8)
            assert np.all(r > 0)                                  # a true determinant ratio at half filling
        assert np.allclose(gg[c, 0], gu, atol=1e-7) and np.allclose(gg[c, 1], gd, atol=1e-7)


@pytest.mark.parametrize("size", [20, 24])
def test_physics_sweep_large_lattices_match_oracle(size):
    """Physics-mode sweep on the multi-column tensor-memory slice paths: N = 400 (padded 448 = 7 x 64, two columns per thread, delay
    depth 24) and N = 576 (unpadded = 9 x 64, three columns, depth 16; BASELINE configs[4] size) at beta = 0.4, L = 4 - same accept /
    reject decisions as the oracle's `physics_sweep` on the same uniforms, G to 1e-9."""
    from latticeqmc_b200 import SweepEngine
    ham = so.ideal_square_kinetic(size, 1.0, 0.0)
    n, lt = size * size, 4
    dtau, lamb, exp_k = so.set_beta_constants(ham, 4.0, 0.4, lt)
    exp_k_inv = expm(dtau * ham)
    field = so.initial_field(n, lt, 77)
    uni = np.random.RandomState(9).rand(1, 1, lt, n)
    with SweepEngine(exp_k, lamb, lt, n_chains=1, exp_k_inv=exp_k_inv, mode="physics", trace=True) as eng:
        assert eng.info()["n_pad"] == (448 if size == 20 else 576)
        eng.set_field(field[None])
        eng.sweep(1, uni)
        acc, ratio = eng.get_trace()
        gg = eng.get_g()[0]
        out = eng.get_field()[0]
    h = field.copy()
    gu, gd, r, a = so.physics_sweep(h, exp_k, exp_k_inv, lamb, uni[0, 0])
    assert 0.2 < a.mean() < 0.95
    assert np.array_equal(a, acc[0, 0]) and np.array_equal(h, out)
    assert np.allclose(r, ratio[0, 0], rtol=1e-8, atol=1e-10)
    assert np.abs(gg[0] - gu).max() < 1e-9 and np.abs(gg[1] - gd).max() < 1e-9


def test_physics_mode_agrees_with_exact_diagonalisation():
    """2x2, U=4, t=1, beta=2, dtau=0.1, half filling: 512 chains x 150 measured sweeps.  ED: n = 0.5/0.5,
    <n_up n_dn> = 0.0873, local moment 0.8254.  Tolerance = Trotter error O(U t dtau^2) + statistics."""
    from latticeqmc_b200 import SweepEngine
    ham, lamb, exp_k, exp_k_inv = _setup(4.0, 2.0, 20)
    hop = ham.copy()
    np.fill_diagonal(hop, 0.0)
    exact = ed.thermal_observables(hop, 4.0, 2.0, 2.0)
    n, lt, chains = 4, 20, 512
    fields = np.stack([so.initial_field(n, lt, 900 + c) for c in range(chains)])
    with SweepEngine(exp_k, lamb, lt, n_chains=chains, exp_k_inv=exp_k_inv, mode="physics", arith="fma") as eng:
        eng.set_field(fields)
        eng.sweep(50, None, seed=77)
        eng.sweep(150, None, seed=77, measure=True)
        m = eng.get_measurements()
    per_chain = m["obs_sum"].mean(axis=2) / m["n_meas"][:, None]           # (chains, 3)
    mean = per_chain.mean(0)
    err = per_chain.std(0, ddof=1) / np.sqrt(chains)
    n_up, n_dn, docc = mean
    assert err.max() < 5e-3
    assert abs(n_up - exact["n_up"]) < 0.01 + 4 * err[0]
    assert abs(n_dn - exact["n_dn"]) < 0.01 + 4 * err[1]
    assert abs(docc - exact["docc"]) < 0.01 + 4 * err[2]
    moment = n_up + n_dn - 2 * docc
    assert abs(moment - exact["moment"]) < 0.02 + 8 * err.max()


# ---- QR/UDV-stabilised recompute (north_star item 3) ----------------------------------------------------------------

def _kin(kind, size, mu):
    return so.ideal_square_kinetic(size, 1.0, mu) if kind == "square" else so.ideal_ring_kinetic(size, 1.0, mu)


@pytest.mark.parametrize("case", [
    # kind, size, U, beta, L, mu, l0, chunk
    ("square", 2, 4.0, 2.0, 20, 0.0, 0, 8),          # N = 4: one ragged QR panel
    ("square", 4, 4.0, 4.0, 40, 0.0, 7, 10),         # N = 16
    ("square", 6, 6.0, 6.0, 60, 0.0, 59, 8),         # N = 36 (padded to 64), ill-conditioned product
    ("ring", 64, 8.0, 8.0, 80, 0.0, 0, 10),          # BASELINE configs[2]: cond(prod B) ~ 1e24
    ("ring", 64, 8.0, 8.0, 80, 4.0, 41, 10),         # same with the reference's diag(K) = -U/2 (H6)
    ("square", 10, 4.0, 4.0, 40, 0.0, 3, 8),         # N = 100 (padded to 128)
    ("square", 12, 4.0, 2.0, 20, 0.0, 19, 5),        # N = 144 (padded to 256)
    ("square", 16, 4.0, 8.0, 80, 0.0, 0, 8),         # BASELINE configs[3] at true half filling
    ("square", 9, 4.0, 3.0, 30, 0.0, 11, 6),         # N = 81: odd size above 64 (column tiles of the DMMA reflector straddle N)
    ("square", 24, 6.0, 10.0, 100, 0.0, 0, 8),       # BASELINE configs[4] size: N = 576 = 9 x 64 (half GEMM tile at the right edge), beta = 10
    ("square", 23, 4.0, 4.0, 40, 0.0, 5, 8),         # N = 529, padded to 576
])
def test_stabilised_recompute_matches_oracle(case):
    """`lqmc_recompute_stable` against the oracle's NumPy QR/UDV (`physics_g_stable`, same pre-pivoted scheme) and,
    where the product is well conditioned, against the plain inverse."""
    from latticeqmc_b200 import SweepEngine
    kind, size, u, beta, lt, mu, l0, chunk = case
    ham = _kin(kind, size, mu)
    n = ham.shape[0]
    dtau, lamb, exp_k = so.set_beta_constants(ham, u, beta, lt)
    chains = 2
    fields = np.stack([so.initial_field(n, lt, 700 + c) for c in range(chains)])
    with SweepEngine(exp_k, lamb, lt, n_chains=chains, exp_k_inv=expm(dtau * ham), mode="physics") as eng:
        eng.set_field(fields)
        eng.recompute_stable(l0, chunk)
        gg = eng.get_g()
    for c in range(chains):
        for si, sigma in enumerate((+1, -1)):
            ref = so.physics_g_stable(fields[c], exp_k, lamb, l0, sigma, chunk)
            err = np.abs(gg[c, si] - ref).max()
            # both are roundoff-limited approximations of the same matrix (chunk conditioning ~1e-9 at U=8, 10 factors)
            assert err < 2e-9 * max(1.0, np.abs(ref).max()), (case, c, sigma, err)
    if beta <= 2.0:
        naive = so.physics_g_naive(fields[0], exp_k, lamb, l0, +1)
        assert np.abs(gg[0, 0] - naive).max() < 1e-9


@pytest.mark.parametrize("case", [("square", 2, 4.0, 2.0, 20, 5, 3), ("square", 4, 4.0, 4.0, 40, 8, 2), ("ring", 64, 8.0, 8.0, 80, 10, 1),
                                   ("square", 4, 4.0, 2.0, 20, 8, 2),       # ragged: segments of 8, 8, 4 slices
                                   ("square", 10, 4.0, 1.0, 10, 4, 2),      # N = 100: large-lattice kernel, segments 4, 4, 2
                                   ("square", 3, 4.0, 2.0, 20, 20, 2),      # one segment = stabilised sweep start only
                                   # the benchmarked sizes (round 2): two-sided combine + wrap_first / skip_last_wrap segmentation at
                                   # N = 256 (BASELINE configs[3] lattice; 1 chain runs on a cluster of 8 CTAs, 2 chains on 8 as well)
                                   # and N = 576 (configs[4] lattice, multi-column tensor-memory slice path)
                                   ("square", 16, 4.0, 2.4, 24, 8, 2),
                                   ("square", 24, 6.0, 1.6, 16, 8, 1)])
def test_stabilised_physics_sweep_matches_oracle(case):
    """Physics-mode sweeps with `stab_every`: G rebuilt by QR/UDV at the top of every segment, wraps inside.
    Same decisions as the oracle's `physics_sweep(stab_every=k)` on the same uniforms, G(0) within 1e-8."""
    from latticeqmc_b200 import SweepEngine
    kind, size, u, beta, lt, k, chains = case
    ham = _kin(kind, size, 0.0)
    n = ham.shape[0]
    dtau, lamb, exp_k = so.set_beta_constants(ham, u, beta, lt)
    exp_k_inv = expm(dtau * ham)
    fields = np.stack([so.initial_field(n, lt, 800 + c) for c in range(chains)])
    sweeps = 2
    uni = np.random.RandomState(9).rand(chains, sweeps, lt, n)
    with SweepEngine(exp_k, lamb, lt, n_chains=chains, exp_k_inv=exp_k_inv, mode="physics", trace=True, stab_every=k) as eng:
        eng.set_field(fields)
        eng.sweep(sweeps, uni, measure=True)
        acc, ratio = eng.get_trace()
        gg, ff, m = eng.get_g(), eng.get_field(), eng.get_measurements()
    for c in range(chains):
        h = fields[c].copy()
        for s in range(sweeps):
            gu, gd, r, a = so.physics_sweep(h, exp_k, exp_k_inv, lamb, uni[c, s], stab_every=k)
            # decisions can only differ where u sits within roundoff of the ratio
            close_call = np.abs(uni[c, s] - r) < 1e-7
            assert np.array_equal(a | close_call, acc[c, s] | close_call), (case, c, s)
            assert np.allclose(r, ratio[c, s], rtol=1e-6 if k <= 10 else 1e-4, atol=1e-8)
            assert np.all(r > -1e-9)
        assert np.array_equal(h, ff[c])
        # the engine joins a left stack with the running right product (two-sided), the oracle rebuilds from scratch;
        # in between both propagate by wraps, whose error grows with U and the segment length
        tol = 1e-8 if (u <= 4.0 and k <= 10) else 1e-5
        assert np.abs(gg[c, 0] - gu).max() < tol and np.abs(gg[c, 1] - gd).max() < tol
        assert m["n_meas"][c] == sweeps


def test_stabilisation_is_refused_in_parity_mode():
    from latticeqmc_b200 import SweepEngine
    ham = so.ideal_square_kinetic(2, 1.0, 2.0)
    dtau, lamb, exp_k = so.set_beta_constants(ham, 4.0, 2.0, 20)
    with SweepEngine(exp_k, lamb, 20) as eng:
        with pytest.raises(ValueError):
            eng.set_stabilization(5)


def test_stabilised_physics_agrees_with_ed_at_low_temperature():
    """2x2, U=6, t=1, beta=8 (L=80), half filling: the unstabilised product is useless here (cond ~ 1e20); with
    QR/UDV every 8 slices density, double occupancy and local moment agree with exact diagonalisation."""
    from latticeqmc_b200 import SweepEngine
    u, beta, lt = 6.0, 8.0, 80
    ham, lamb, exp_k, exp_k_inv = _setup(u, beta, lt)
    hop = ham.copy()
    np.fill_diagonal(hop, 0.0)
    exact = ed.thermal_observables(hop, u, u / 2, beta)
    n, chains = 4, 256
    fields = np.stack([so.initial_field(n, lt, 1900 + c) for c in range(chains)])
    with SweepEngine(exp_k, lamb, lt, n_chains=chains, exp_k_inv=exp_k_inv, mode="physics", arith="fma", stab_every=8) as eng:
        eng.set_field(fields)
        eng.sweep(40, None, seed=78)
        eng.sweep(120, None, seed=78, measure=True)
        m = eng.get_measurements()
    per_chain = m["obs_sum"].mean(axis=2) / m["n_meas"][:, None]
    mean = per_chain.mean(0)
    err = per_chain.std(0, ddof=1) / np.sqrt(chains)
    n_up, n_dn, docc = mean
    assert np.all(np.isfinite(mean)) and err.max() < 1e-2
    assert abs(n_up + n_dn - 1.0) < 0.01 + 4 * (err[0] + err[1])
    assert abs(docc - exact["docc"]) < 0.015 + 4 * err[2]
    moment = n_up + n_dn - 2 * docc
    assert abs(moment - exact["moment"]) < 0.03 + 8 * err.max()


def test_checkpoint_resume_repeats_the_uninterrupted_run(tmp_path):
    """field + sweep counter + accumulators is the whole Markov state with the device Philox stream; a run
    resumed in a fresh engine is bit-identical to the uninterrupted one (field, G, accumulators)."""
    from latticeqmc_b200 import SweepEngine
    ham, lamb, exp_k, exp_k_inv = _setup(4.0, 2.0, 20)
    n, lt, chains, seed = 4, 20, 5, 4242
    fields = np.stack([so.initial_field(n, lt, 60 + c) for c in range(chains)])
    kw = dict(n_chains=chains, exp_k_inv=exp_k_inv, mode="physics", stab_every=5, chain_offset=3)
    with SweepEngine(exp_k, lamb, lt, **kw) as eng:
        eng.set_field(fields)
        eng.sweep(3, None, seed=seed)
        eng.sweep(4, None, seed=seed, measure=True)
        ref_field, ref_g, ref_m = eng.get_field(), eng.get_g(), eng.get_measurements()
    path = str(tmp_path / "state.npz")
    with SweepEngine(exp_k, lamb, lt, **kw) as eng:
        eng.set_field(fields)
        eng.sweep(3, None, seed=seed)
        eng.sweep(2, None, seed=seed, measure=True)
        eng.save_checkpoint(path, seed=seed)
    with SweepEngine(exp_k, lamb, lt, n_chains=chains, exp_k_inv=exp_k_inv, mode="physics", stab_every=5) as eng:
        assert eng.load_checkpoint(path) == seed
        assert eng.info()["sweep_counter"] == 5
        eng.sweep(2, None, seed=seed, measure=True)
        f, g, m = eng.get_field(), eng.get_g(), eng.get_measurements()
    assert np.array_equal(f, ref_field) and np.array_equal(g, ref_g)
    for key in ("g_sum", "obs_sum", "n_meas", "n_accepted"):
        assert np.array_equal(m[key], ref_m[key]), key


def test_checkpoint_identity_is_checked_and_suffix_is_optional(tmp_path):
    """A checkpoint of another simulation (different beta -> lamb / exp_k, different arithmetic) must not load silently, and
    `save('ckpt')` / `load('ckpt')` agree on the file name (np.savez appends `.npz`).  With `numpy_rng=True` the host MT19937
    stream - the one the drop-in `rng="numpy"` mode consumes - is stored and restored."""
    from latticeqmc_b200 import SweepEngine
    ham, lamb, exp_k, exp_k_inv = _setup(4.0, 2.0, 20)
    n, lt = 4, 20
    fields = np.stack([so.initial_field(n, lt, 70 + c) for c in range(2)])
    stem = str(tmp_path / "ckpt")                                   # no suffix
    with SweepEngine(exp_k, lamb, lt, n_chains=2, exp_k_inv=exp_k_inv, mode="physics") as eng:
        eng.set_field(fields)
        eng.sweep(2, None, seed=1, measure=True)
        np.random.seed(99)
        np.random.rand(7)
        eng.save_checkpoint(stem, seed=1, numpy_rng=True)
        expect_next = np.random.rand(3)
        np.random.seed(5)                                           # disturb the stream
        assert eng.load_checkpoint(stem) == 1
        assert np.array_equal(np.random.rand(3), expect_next)
    ham2, lamb2, exp_k2, exp_k_inv2 = _setup(4.0, 3.0, 20)            # other beta
    with SweepEngine(exp_k2, lamb2, lt, n_chains=2, exp_k_inv=exp_k_inv2, mode="physics") as eng:
        with pytest.raises(ValueError):
            eng.load_checkpoint(stem)
    with SweepEngine(exp_k, lamb, lt, n_chains=2, exp_k_inv=exp_k_inv, mode="physics", arith="fma") as eng:
        with pytest.raises(ValueError):
            eng.load_checkpoint(stem + ".npz")


def test_chain_statistics_on_device_accumulators():
    """Error bars from `chain_statistics` on a real run: ED values inside mean +- 4 sigma + Trotter error."""
    from latticeqmc_b200 import SweepEngine
    from latticeqmc_b200.tools import chain_statistics
    ham, lamb, exp_k, exp_k_inv = _setup(4.0, 2.0, 20)
    hop = ham.copy()
    np.fill_diagonal(hop, 0.0)
    exact = ed.thermal_observables(hop, 4.0, 2.0, 2.0)
    chains = 256
    fields = np.stack([so.initial_field(4, 20, 300 + c) for c in range(chains)])
    with SweepEngine(exp_k, lamb, 20, n_chains=chains, exp_k_inv=exp_k_inv, mode="physics", arith="fma") as eng:
        eng.set_field(fields)
        eng.sweep(40, None, seed=5)
        eng.sweep(100, None, seed=5, measure=True)
        st = chain_statistics(eng.get_measurements())
    assert st["n_chains"] == chains and st["n_meas"] == chains * 100
    assert 0 < st["docc"][1] < 5e-3 and 0 < st["moment"][1] < 1e-2
    assert abs(st["density"][0] - 1.0) < 0.01 + 4 * st["density"][1]
    assert abs(st["docc"][0] - exact["docc"]) < 0.01 + 4 * st["docc"][1]
    assert abs(st["moment"][0] - exact["moment"]) < 0.02 + 4 * st["moment"][1]
