"""det mode on the device (`lqmc_sweep_det`, csrc/sweep_det.cuh) against the reference's own det-mode runs
(`tests/golden/det_*.npz`, recorded from `run_lqmc(det_mode=True)`, lqmc.py:236-299) and against the oracle.

Tolerance: accept/reject sequences and fields must be identical and the measured G within 1e-10 of max|G|
(BASELINE.json north_star tolerance).  The determinant ratios themselves carry the conditioning of the B product
(cond ~ 5e5 at 2x2 beta=2) times the roundoff of two different summation orders (OpenBLAS dgemm /
getrf and NumPy's sign * exp(sum log|u_ii|) against sequential FMA and a pivot product): 1e-9 relative, written below.
"""
import numpy as np
import pytest

from oracle import sweep_oracle as so
from oracle import ed

pytestmark = pytest.mark.gpu

RTOL = 1e-10       # G, relative to max|G|
RTOL_RATIO = 1e-9  # determinant ratios (measured: 1.7e-10 at 2x2 beta=2, 2e-13 at the better-conditioned cases)


def _rel(a, b):
    err = float(np.max(np.abs(a - b) / np.abs(b)))
    print("max relative ratio deviation", err)
    return err


def _engine(exp_k, lamb, lt, **kw):
    from latticeqmc_b200 import SweepEngine
    return SweepEngine(exp_k, lamb, lt, **kw)


@pytest.mark.parametrize("name", ["det_2x2", "det_3x2"])
def test_det_mode_reference_run(golden, name):
    g = golden(name)
    n, lt = g["field0"].shape
    warm, meas = int(g["warm"]), int(g["meas"])
    with _engine(g["exp_k"], float(g["lamb"]), lt, trace=True) as eng:
        eng.set_field(g["field0"][None])
        eng.sweep_det(warm, g["uniforms"][None, :warm])                       # warmup_loop_det
        acc, ratio = eng.get_trace()
        assert np.array_equal(acc[0], g["accs"][:warm])
        assert _rel(ratio[0], g["ratios"][:warm]) <= RTOL_RATIO
        assert np.array_equal(eng.get_field()[0], g["fields"][warm - 1])
        eng.reset_measurements()
        eng.sweep_det(meas, g["uniforms"][None, warm:], measure=True)          # measure_loop_det
        acc, ratio = eng.get_trace()
        assert np.array_equal(acc[0], g["accs"][warm:])
        assert _rel(ratio[0], g["ratios"][warm:]) <= RTOL_RATIO
        assert np.array_equal(eng.get_field()[0], g["fields"][-1])
        m = eng.get_measurements()
        assert m["n_meas"][0] == meas
        gf = m["g_sum"][0] / meas
        assert np.max(np.abs(gf - g["gf"])) <= RTOL * np.max(np.abs(g["gf"]))
        assert m["n_accepted"][0] == int(g["accs"][warm:].sum())
        # old_det handed back = det of the final field's M(0) up to the roundoff of a different cyclic order
        want = so.det_product(g["fields"][-1].copy(), g["exp_k"], float(g["lamb"]))
        assert np.isclose(eng.get_det()[0], want, rtol=1e-9, atol=0)


def test_det_mode_batch_against_oracle():
    """Several chains, N = 16 (4x4, U=4, beta=1, L=8), device Philox uniforms replayed on the host for the oracle."""
    from latticeqmc_b200 import philox_uniforms
    ham = so.ideal_square_kinetic(4, t=1.0, mu=2.0)
    n, lt, chains, seed = 16, 8, 3, 77
    dtau, lamb, exp_k = so.set_beta_constants(ham, 4.0, 1.0, lt)
    fields = np.stack([so.initial_field(n, lt, 100 + c) for c in range(chains)])
    with _engine(exp_k, lamb, lt, n_chains=chains, trace=True, chain_offset=5) as eng:
        eng.set_field(fields)
        eng.sweep_det(2, None, seed=seed)
        acc, ratio = eng.get_trace()
        out = eng.get_field()
    for c in range(chains):
        h = fields[c].copy()
        old = so.det_product(h, exp_k, lamb)
        for s in range(2):
            u = philox_uniforms(seed, 5 + c, s, n * lt).reshape(lt, n)
            old, r, a = so.det_update_step(h, exp_k, lamb, old, u)
            assert np.array_equal(a, acc[c, s]), f"chain {c} sweep {s}"
            assert _rel(ratio[c, s], r) <= RTOL_RATIO
        assert np.array_equal(h, out[c])


def test_det_mode_drop_in_replays_reference_run(golden):
    """`LatticeQMC(det_mode=True).run_lqmc()` through the kept API, seeded like the recorded reference run: the field
    drawn by `Configuration`, the uniforms of the global legacy stream, warm-up + measurement loops - same G."""
    from latticeqmc_b200 import HubbardModel, LatticeQMC
    g = golden("det_2x2")
    model = HubbardModel(u=4, t=1)
    model.build_square(2)
    np.random.seed(61)
    solver = LatticeQMC(model, 2.0, 20, warmup=int(g["warm"]), sweeps=int(g["meas"]), det_mode=True, log_lvl=None, trace=True)
    assert np.array_equal(solver.config.config, g["field0"])
    gf = solver.run_lqmc()
    assert np.array_equal(solver.config.config, g["fields"][-1])
    assert np.max(np.abs(gf - g["gf"])) <= RTOL * np.max(np.abs(g["gf"]))
    acc, ratio = solver.last_trace
    assert np.array_equal(acc, g["accs"][int(g["warm"]):])


def test_det_mode_matches_exact_diagonalisation():
    """The det-mode sampler is a correct one - of the model with mu_true = mu + U/2.  2x2, U=4, beta=2,
    L=20, 64 chains x 400 measured sweeps against ED at mu = 4: n = 0.6802, <n_up n_dn> = 0.3874."""
    ham = so.ideal_square_kinetic(2, t=1.0, mu=2.0)
    hop = ham.copy()
    np.fill_diagonal(hop, 0.0)
    exact = ed.thermal_observables(hop, 4.0, 4.0, 2.0)
    n, lt, chains = 4, 20, 64
    dtau, lamb, exp_k = so.set_beta_constants(ham, 4.0, 2.0, lt)
    fields = np.stack([so.initial_field(n, lt, 300 + c) for c in range(chains)])
    with _engine(exp_k, lamb, lt, n_chains=chains) as eng:
        eng.set_field(fields)
        eng.sweep_det(100, None, seed=5)
        eng.reset_measurements()
        eng.sweep_det(400, None, seed=5, measure=True)
        m = eng.get_measurements()
    per_chain = m["obs_sum"].mean(axis=2) / m["n_meas"][:, None]
    mean = per_chain.mean(0)
    err = per_chain.std(0, ddof=1) / np.sqrt(chains)
    tol = 5 * err + 0.01                                   # statistics + Trotter error O(U t dtau^2)
    assert abs(mean[0] - exact["n_up"]) < tol[0] and abs(mean[1] - exact["n_dn"]) < tol[1]
    assert abs(mean[2] - exact["docc"]) < tol[2]


def test_old_det_is_carried_between_calls(golden):
    """`_update_step_det(old_det)` takes the determinant the ratios are measured against as an argument and the reference's loops
    carry it from sweep to sweep (lqmc.py:236-259, 264-270).  A loop split into several engine calls must therefore continue from
    the carried value, not from a re-derived one: 3 sweeps in one call == 1 + 2 sweeps with `old_det="carry"` == the same with the
    value read back and passed in explicitly - decisions, ratios, field and final old_det bit for bit."""
    g = golden("det_2x2")
    n, lt = g["field0"].shape
    uni = g["uniforms"][None, :3]
    runs = []
    for how in ("one", "carry", "explicit"):
        with _engine(g["exp_k"], float(g["lamb"]), lt, trace=True) as eng:
            eng.set_field(g["field0"][None])
            if how == "one":
                eng.sweep_det(3, uni)
                acc, ratio = eng.get_trace()
            else:
                eng.sweep_det(1, uni[:, :1])
                a1, r1 = eng.get_trace()
                eng.sweep_det(2, uni[:, 1:], old_det=("carry" if how == "carry" else eng.get_det()))
                a2, r2 = eng.get_trace()
                acc, ratio = np.concatenate([a1, a2], axis=1), np.concatenate([r1, r2], axis=1)
            runs.append((acc.copy(), ratio.copy(), eng.get_field().copy(), eng.get_det().copy()))
    for other in runs[1:]:
        for x, y in zip(runs[0], other):
            assert np.array_equal(x, y)
    assert np.array_equal(runs[0][0][0], g["accs"][:3])
    with _engine(g["exp_k"], float(g["lamb"]), lt) as eng:
        with pytest.raises(Exception):
            eng.sweep_det(1, uni[:, :1], old_det="carry")            # nothing to carry yet


@pytest.mark.parametrize("n,lt,chains", [(100, 4, 2), (260, 2, 1)])
def test_det_mode_large_lattice_against_oracle(n, lt, chains):
    """The reference's det mode has no size limit (lqmc.py:236-259).  Above 64 sites the device kernel keeps its matrices
    in a global-memory workspace: ring of 100 sites (one thread per row element) and of 260 sites (more rows than threads
    per spin group), two chains, one sweep + one measured sweep against the oracle on the same uniforms."""
    # determinants ~ 2^N per spin: mu = 0 and a small beta keep det(M_up) det(M_dn) inside the double range at N = 260
    ham = so.ideal_ring_kinetic(n, t=1.0, mu=0.0 if n > 200 else 2.0)
    dtau, lamb, exp_k = so.set_beta_constants(ham, 4.0, 0.2 if n > 200 else 0.4, lt)
    fields = np.stack([so.initial_field(n, lt, 500 + c) for c in range(chains)])
    uni = np.random.RandomState(9).rand(chains, 2, lt, n)
    with _engine(exp_k, lamb, lt, n_chains=chains, trace=True) as eng:
        eng.set_field(fields)
        eng.sweep_det(1, uni[:, :1])
        a1, r1 = eng.get_trace()
        eng.sweep_det(1, uni[:, 1:], measure=True, old_det="carry")
        a2, r2 = eng.get_trace()
        out = eng.get_field()
        m = eng.get_measurements()
    for c in range(chains):
        h = fields[c].copy()
        old = so.det_product(h, exp_k, lamb)
        for s, (a, r) in enumerate(((a1, r1), (a2, r2))):
            old, rr, aa = so.det_update_step(h, exp_k, lamb, old, uni[c, s])
            assert np.array_equal(aa, a[c, 0]), f"chain {c} sweep {s}"
            assert _rel(r[c, 0], rr) <= RTOL_RATIO
        assert np.array_equal(h, out[c])
        gu = np.linalg.inv(so.get_m(h, exp_k, lamb, 0, +1))
        gd = np.linalg.inv(so.get_m(h, exp_k, lamb, 0, -1))
        assert np.max(np.abs(m["g_sum"][c, 0] - gu)) <= RTOL * np.max(np.abs(gu))
        assert np.max(np.abs(m["g_sum"][c, 1] - gd)) <= RTOL * np.max(np.abs(gd))
