"""Generate the golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py [--only NAME]

The reference (`/root/reference/lqmc`) imports matplotlib at module top
(lattice.py:11-13, configuration.py:9), which is not installed; five empty stub modules in
``sys.modules`` make ``import lqmc`` work without touching reference code.

What is recorded per case (all from the reference's own ``LatticeQMC._update_step``,
lqmc.py:301-347):
  * the int8 field before and after each sweep, the uniforms the sweep consumed (replayed from a
    copy of the MT19937 state: one ``rand()`` per proposal, lqmc.py:317),
  * per-proposal ``ratio``/``acc`` captured by replacing the *instance* attribute ``_debug``
    (called once per proposal after they are set, lqmc.py:316-317,335),
  * G after the proposals of selected slices, read from the caller frame's locals inside that
    ``_debug`` hook (the reference exposes no other seam), and the returned end-of-sweep G.

Cases: cfg1 free-running (10 sweeps); cfg2 / cfg3 one full sweep with snapshots of a few slices;
cfg4 (16x16, ideal periodic K because the reference's lattice builder cannot build it, H10) two
slices; the U=0 known answer of exact.py:27-54.
"""
import argparse
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.collections", "mpl_toolkits",
                 "mpl_toolkits.axes_grid1"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.collections"].LineCollection = object
    sys.modules["mpl_toolkits.axes_grid1"].make_axes_locatable = lambda *a, **k: None
    # make sure `lqmc` resolves to the reference, not to this repo's drop-in shim
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != os.path.abspath(os.path.join(HERE, "..", ".."))]
    sys.path.insert(0, REFERENCE)
    import lqmc  # noqa
    assert os.path.abspath(lqmc.__file__).startswith(REFERENCE), lqmc.__file__
    return lqmc


class IdealModel:
    """Duck-typed stand-in for HubbardModel where the reference cannot build the lattice (H10).
    LatticeQMC only touches ``n_sites, u, t, mu, ham_kinetic()`` (lqmc.py:47-76,105-112)."""

    def __init__(self, ham, u, t=1.0):
        self._ham = ham
        self.u, self.t, self.mu = u, t, u / 2
        self.n_sites = ham.shape[0]

    def ham_kinetic(self):
        return self._ham.copy()


def record_sweep(solver, snap_slices=()):
    """Run one reference ``_update_step`` and return everything it did."""
    n, lt = solver.n_sites, solver.time_steps
    field_in = solver.config.config.copy()
    state = np.random.get_state()
    uniforms = np.random.rand(lt * n).reshape(lt, n)       # what the sweep is about to consume
    np.random.set_state(state)
    ratios = np.empty((lt, n))
    accs = np.zeros((lt, n), dtype=bool)
    snaps = {}

    def hook(i, l):
        step = lt - 1 - l
        ratios[step, i] = solver.ratio
        accs[step, i] = solver.acc
        if i == n - 1 and l in snap_slices:
            loc = sys._getframe(1).f_locals
            snaps[l] = (loc["gf_up"].copy(), loc["gf_dn"].copy())

    solver._debug = hook
    t0 = time.time()
    gf_up, gf_dn = solver._update_step()
    dt = time.time() - t0
    # the stream must have advanced by exactly N*L draws
    probe = np.random.get_state()
    np.random.set_state(state)
    np.random.rand(lt * n)
    assert np.array_equal(np.random.get_state()[1], probe[1]) and np.random.get_state()[2] == probe[2]
    return dict(field_in=field_in, field_out=solver.config.config.copy(), uniforms=uniforms,
                ratios=ratios, accs=accs, gf_up=np.array(gf_up), gf_dn=np.array(gf_dn),
                snaps=snaps, seconds=dt)


def make_solver(lqmc, model, beta, lt, seed):
    np.random.seed(seed)
    return lqmc.LatticeQMC(model, beta, lt, warmup=0, sweeps=0, log_lvl=None)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def case_cfg1(lqmc):
    """2x2 square, U=4, t=1, beta=2, L=20, 10 free-running sweeps from seed 11."""
    model = lqmc.HubbardModel(u=4, t=1)
    model.build_square(2)
    solver = make_solver(lqmc, model, 2.0, 20, seed=11)
    recs = [record_sweep(solver) for _ in range(10)]
    save("cfg1_2x2_free", ham=model.ham_kinetic(), u=4.0, beta=2.0, lamb=solver.lamb, exp_k=solver.exp_k,
         field0=recs[0]["field_in"],
         uniforms=np.stack([r["uniforms"] for r in recs]), ratios=np.stack([r["ratios"] for r in recs]),
         accs=np.stack([r["accs"] for r in recs]), fields=np.stack([r["field_out"] for r in recs]),
         gf_up=np.stack([r["gf_up"] for r in recs]), gf_dn=np.stack([r["gf_dn"] for r in recs]))


def case_small(lqmc):
    """4x4 U=4 beta=2 L=20 (3 sweeps) and 2x3 U=6 beta=1 L=10 (3 sweeps): extra small shapes,
    including a non-power-of-two N, for padding / ragged-size coverage."""
    for tag, (w, h), u, beta, lt, seed in (("4x4", (4, 4), 4, 2.0, 20, 5), ("3x2", (3, 2), 6, 1.0, 10, 6)):
        model = lqmc.HubbardModel(u=u, t=1)
        model.build(w, h, cycling=(0, 1) if w == h else 0)
        solver = make_solver(lqmc, model, beta, lt, seed=seed)
        recs = [record_sweep(solver) for _ in range(3)]
        save(f"small_{tag}_free", ham=model.ham_kinetic(), u=float(u), beta=beta, lamb=solver.lamb,
             exp_k=solver.exp_k, field0=recs[0]["field_in"],
             uniforms=np.stack([r["uniforms"] for r in recs]), ratios=np.stack([r["ratios"] for r in recs]),
             accs=np.stack([r["accs"] for r in recs]), fields=np.stack([r["field_out"] for r in recs]),
             gf_up=np.stack([r["gf_up"] for r in recs]), gf_dn=np.stack([r["gf_dn"] for r in recs]))


def case_n64(lqmc, name, model, beta, lt, seed, snap_slices):
    solver = make_solver(lqmc, model, beta, lt, seed=seed)
    m_up = solver.get_m(0, +1)
    m_dn = solver.get_m(0, -1)
    g0_up, g0_dn = np.linalg.inv(m_up), np.linalg.inv(m_dn)       # what _update_step starts from
    rec = record_sweep(solver, snap_slices)
    print(f"{name}: reference sweep took {rec['seconds']:.1f} s, accept {rec['accs'].mean():.3f}")
    arrays = dict(ham=model.ham_kinetic(), u=float(model.u), beta=beta, lamb=solver.lamb, exp_k=solver.exp_k,
                  field0=rec["field_in"], field1=rec["field_out"], uniforms=rec["uniforms"],
                  ratios=rec["ratios"], accs=rec["accs"], gf_up=rec["gf_up"], gf_dn=rec["gf_dn"],
                  g0_up=g0_up, g0_dn=g0_dn, snap_slices=np.array(sorted(rec["snaps"])))
    for l, (gu, gd) in rec["snaps"].items():
        arrays[f"post{l}_up"] = gu
        arrays[f"post{l}_dn"] = gd
    save(name, **arrays)


def case_cfg2(lqmc):
    """8x8 square, U=4, beta=4, L=40: one full sweep, snapshots after slices 39,38,20,19,1,0."""
    model = lqmc.HubbardModel(u=4, t=1)
    model.build_square(8)
    case_n64(lqmc, "cfg2_8x8_sweep", model, 4.0, 40, seed=21, snap_slices=(39, 38, 20, 19, 1, 0))


def case_cfg3(lqmc):
    """Ring N=64, U=8, beta=8, L=80: one full sweep, snapshots after slices 79,78,40,39."""
    model = lqmc.HubbardModel(u=8, t=1)
    model.build(64)
    case_n64(lqmc, "cfg3_ring64_sweep", model, 8.0, 80, seed=31, snap_slices=(79, 78, 40, 39))


def case_cfg4(lqmc):
    """16x16 (ideal periodic K, H10), U=4, beta=8, L=80: the reference needs ~20 min per sweep, so
    only the first two slices are run: proposals(79) -> wrap -> proposals(78).  A subclass-free
    trick stops the sweep early: the `_debug` hook raises after slice 78's last proposal."""
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle.sweep_oracle import ideal_square_kinetic
    ham = ideal_square_kinetic(16, t=1.0, mu=2.0)
    model = IdealModel(ham, u=4)
    lt, n = 80, 256
    solver = make_solver(lqmc, model, 8.0, lt, seed=41)
    field_in = solver.config.config.copy()
    g0_up = np.linalg.inv(solver.get_m(0, +1))
    g0_dn = np.linalg.inv(solver.get_m(0, -1))
    state = np.random.get_state()
    uniforms = np.random.rand(2 * n).reshape(2, n)
    np.random.set_state(state)
    ratios = np.empty((2, n))
    accs = np.zeros((2, n), dtype=bool)
    snaps = {}

    class Stop(Exception):
        pass

    def hook(i, l):
        step = lt - 1 - l
        ratios[step, i] = solver.ratio
        accs[step, i] = solver.acc
        if i == n - 1:
            loc = sys._getframe(1).f_locals
            snaps[l] = (loc["gf_up"].copy(), loc["gf_dn"].copy())
            if l == lt - 2:
                raise Stop

    solver._debug = hook
    t0 = time.time()
    try:
        solver._update_step()
    except Stop:
        pass
    print(f"cfg4: two reference slices took {time.time() - t0:.1f} s, accept {accs.mean():.3f}")
    save("cfg4_16x16_slices", ham=ham, u=4.0, beta=8.0, lamb=solver.lamb, exp_k=solver.exp_k,
         field0=field_in, field1=solver.config.config.copy(), uniforms=uniforms, ratios=ratios, accs=accs,
         g0_up=g0_up, g0_dn=g0_dn, post79_up=snaps[79][0], post79_dn=snaps[79][1],
         post78_up=snaps[78][0], post78_dn=snaps[78][1])


def case_u0(lqmc):
    """U=0 known answer (exact.py:37-54): 10-site open chain, beta=4, L=40.  lamb=0, every ratio
    is 1, every proposal accepted, and G = (I+exp(-beta K))^-1 = -pole_gf(tau=0)."""
    model = lqmc.HubbardModel(u=0, t=1, mu=0)
    model.build(10, cycling=None)
    solver = make_solver(lqmc, model, 4.0, 40, seed=51)
    rec = record_sweep(solver)
    tau, gf_tau = lqmc.compute_pole_gf_tau(model.ham_kinetic(), 4.0)
    save("u0_chain10", ham=model.ham_kinetic(), u=0.0, beta=4.0, lamb=float(solver.lamb), exp_k=solver.exp_k,
         field0=rec["field_in"], field1=rec["field_out"], uniforms=rec["uniforms"], ratios=rec["ratios"],
         accs=rec["accs"], gf_up=rec["gf_up"], gf_dn=rec["gf_dn"], pole_gf_tau0=gf_tau[:, :, 0])


def case_lattice(lqmc):
    """Kinetic matrices from the reference's own lattice builder wherever it can build
    (SURVEY.md B.1), for the drop-in HubbardModel/Lattice to be compared against."""
    out = {}
    for size in (2, 3, 4, 5, 8):
        m = lqmc.HubbardModel(u=4, t=1)
        m.build_square(size)
        out[f"square{size}"] = m.ham_kinetic()
    m = lqmc.HubbardModel(u=8, t=1)
    m.build(64)
    out["ring64"] = m.ham_kinetic()
    m = lqmc.HubbardModel(u=0, t=1, mu=0)
    m.build(10, cycling=None)
    out["open10"] = m.ham_kinetic()
    m = lqmc.HubbardModel(u=6, t=1)
    m.build(3, 2, cycling=0)
    out["rect3x2_c0"] = m.ham_kinetic()
    save("lattice_ham", **out)


def case_det(lqmc):
    """det_mode (lqmc.py:236-299), the reference's slow validation sampler: 2x2 U=4 beta=2 L=20 (2 warm-up + 3
    measured sweeps) and 3x2 U=6 beta=1 L=10 (1 + 2), through the reference's own `run_lqmc`.  Per-proposal
    ratio / acc via the `_debug` hook; the uniforms are replayed from a copy of the MT19937 state."""
    for tag, (w, h), u, beta, lt, seed, warm, meas in (("2x2", (2, 2), 4, 2.0, 20, 61, 2, 3),
                                                       ("3x2", (3, 2), 6, 1.0, 10, 62, 1, 2)):
        model = lqmc.HubbardModel(u=u, t=1)
        model.build(w, h, cycling=(0, 1) if w == h else 0)
        np.random.seed(seed)
        solver = lqmc.LatticeQMC(model, beta, lt, warmup=warm, sweeps=meas, det_mode=True, log_lvl=None)
        n = solver.n_sites
        field0 = solver.config.config.copy()
        state = np.random.get_state()
        uniforms = np.random.rand((warm + meas) * lt * n).reshape(warm + meas, lt, n)
        np.random.set_state(state)
        ratios, accs, fields = [], [], []

        def hook(i, l, solver=solver, ratios=ratios, accs=accs, fields=fields, n=n):
            ratios.append(solver.ratio)
            accs.append(bool(solver.acc))
            if l == 0 and i == n - 1:
                fields.append(solver.config.config.copy())

        solver._debug = hook
        solver.iter_sweeps = lambda count: range(count)        # no console progress
        gf = solver.run_lqmc()
        probe = np.random.get_state()
        np.random.set_state(state)
        np.random.rand((warm + meas) * lt * n)
        assert np.array_equal(np.random.get_state()[1], probe[1]) and np.random.get_state()[2] == probe[2]
        save(f"det_{tag}", ham=model.ham_kinetic(), u=float(u), beta=beta, lamb=solver.lamb, exp_k=solver.exp_k,
             field0=field0, uniforms=uniforms, warm=warm, meas=meas,
             ratios=np.array(ratios).reshape(warm + meas, lt, n), accs=np.array(accs).reshape(warm + meas, lt, n),
             fields=np.stack(fields), gf=np.asarray(gf))


CASES = dict(cfg1=case_cfg1, small=case_small, cfg2=case_cfg2, cfg3=case_cfg3, cfg4=case_cfg4, u0=case_u0,
             lattice=case_lattice, det=case_det)

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None, choices=sorted(CASES))
    args = ap.parse_args()
    ref = import_reference()
    for key, fn in CASES.items():
        if args.only in (None, key):
            fn(ref)
