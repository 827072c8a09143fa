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


def subsample(g, rows, cols):
    """The parts of a Green's function a fixture keeps when the full matrix would be megabytes: `rows`, `cols` and the
    diagonal.  (The full-matrix comparison is GPU vs oracle in the test; the oracle is pinned on these parts, on every
    ratio - each depends on the whole update history through the diagonal - and on every decision.)"""
    return g[rows, :].copy(), g[:, cols].copy(), np.diag(g).copy()


class Stop(Exception):
    pass


def case_cfg4mid(lqmc):
    """16x16 (ideal periodic K, H10), U=4, beta=8, L=80: the reference's own free-running sweep from slice 79 down to
    slice 10 (about 25 min of the interpreted loop), recording the two mid-sweep slices 44 and 10 where |G| ~ 1e2 and a
    third of the ratios are negative.  Kept: G after the proposals of slices 45 and 11 (the inputs: the
    wrap to 44 / 10 is replayed with the same NumPy calls, lqmc.py:338-345), every ratio / decision / uniform of slices
    44 and 10, the field columns involved, and rows / columns / diagonal of G after the proposals of 44 and 10."""
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle import sweep_oracle as so
    ham = so.ideal_square_kinetic(16, t=1.0, mu=2.0)
    model = IdealModel(ham, u=4)
    lt, n = 80, 256
    solver = make_solver(lqmc, model, 8.0, lt, seed=43)
    field_in = solver.config.config.copy()
    state = np.random.get_state()
    uniforms = np.random.rand(lt * n).reshape(lt, n)
    np.random.set_state(state)
    want = (45, 44, 11, 10)
    ratios = np.empty((lt, n))
    accs = np.zeros((lt, n), dtype=bool)
    snaps = {}

    def hook(i, l):
        step = lt - 1 - l
        ratios[step, i] = solver.ratio
        accs[step, i] = solver.acc
        if i == n - 1:
            if l in want:
                loc = sys._getframe(1).f_locals
                snaps[l] = (loc["gf_up"].copy(), loc["gf_dn"].copy())
            print(f"cfg4mid: slice {l} done, accept so far {accs[:step + 1].mean():.3f}", flush=True)
            if l == min(want):
                raise Stop

    solver._debug = hook
    t0 = time.time()
    try:
        solver._update_step()
    except Stop:
        pass
    print(f"cfg4mid: reference slices 79..10 took {time.time() - t0:.1f} s")
    field_out = solver.config.config.copy()
    rows = np.arange(0, n, 16)
    cols = np.arange(5, n, 16)
    arrays = dict(ham=ham, u=4.0, beta=8.0, lamb=solver.lamb, exp_k=solver.exp_k, rows=rows, cols=cols,
                  field0=field_in, field1=field_out)
    for l in (44, 10):
        step = lt - 1 - l
        # self-check: replaying wrap + slice with the same NumPy calls reproduces the reference bit for bit
        h = field_out.copy()
        h[:, l] = field_in[:, l]
        gu, gd = so.wrap(snaps[l + 1][0], snaps[l + 1][1], h, l + 1, solver.exp_k, solver.lamb)
        pre_max = max(np.abs(gu).max(), np.abs(gd).max())
        r, a = so.slice_proposals(gu, gd, h, l, solver.lamb, uniforms[step])
        assert np.array_equal(a, accs[step]) and np.array_equal(r, ratios[step])
        assert np.array_equal(gu, snaps[l][0]) and np.array_equal(gd, snaps[l][1]) and np.array_equal(h[:, l], field_out[:, l])
        print(f"cfg4mid: slice {l}: max|G| before {pre_max:.3e}, negative ratios {np.mean(ratios[step] < 0):.2f}, "
              f"accept {accs[step].mean():.2f}, max|ratio| {np.abs(ratios[step]).max():.3e}")
        arrays.update({f"in{l}_up": snaps[l + 1][0], f"in{l}_dn": snaps[l + 1][1],
                       f"uniforms{l}": uniforms[step], f"ratios{l}": ratios[step], f"accs{l}": accs[step]})
        for tag, g in (("up", snaps[l][0]), ("dn", snaps[l][1])):
            gr, gc, gdg = subsample(g, rows, cols)
            arrays.update({f"post{l}_{tag}_rows": gr, f"post{l}_{tag}_cols": gc, f"post{l}_{tag}_diag": gdg})
    save("cfg4_16x16_mid", **arrays)


def case_cfg5(lqmc):
    """24x24 square (N = 576, built by the reference's own lattice builder), U=6, beta=10, L=100:
    proposals(99) -> wrap -> proposals(98) of the reference's `_update_step`, teacher-forced from a well-scaled G.
    The reference always starts a sweep from `np.linalg.inv(get_m(0, sigma))` (lqmc.py:303-307), which at beta = 10 is
    pure round-off (max|G| ~ 1e-24).  To run its proposals and wrap on an O(1) Green's function without
    touching reference code, `np.linalg.inv` is wrapped for the two sweep-start calls only and hands back a prepared G0:
    float32-rounded `inv(I + B_99 ... B_90)` (a short, well-conditioned product; float32 values keep the fixture small).
    Everything after that - ratios, decisions, rank-1 loops, the wrap's own `inv(b)` - is the unmodified reference."""
    model = lqmc.HubbardModel(u=6, t=1)
    t0 = time.time()
    model.build_square(24)
    print(f"cfg5: reference lattice builder took {time.time() - t0:.1f} s")
    lt, n = 100, 576
    solver = make_solver(lqmc, model, 10.0, lt, seed=53)
    field_in = solver.config.config.copy()
    g0 = []
    for sigma in (+1, -1):
        prod = 1
        for l in range(lt - 1, lt - 11, -1):
            prod = np.dot(prod, np.dot(solver.exp_k, solver.get_exp_v(l, sigma)))
        g0.append(np.linalg.inv(np.eye(n) + prod).astype(np.float32))
    state = np.random.get_state()
    uniforms = np.random.rand(2 * n).reshape(2, n)
    np.random.set_state(state)
    ratios = np.empty((2, n))
    accs = np.zeros((2, n), dtype=bool)
    snaps = {}

    def hook(i, l):
        step = lt - 1 - l
        ratios[step, i] = solver.ratio
        accs[step, i] = solver.acc
        if i == n - 1:
            loc = sys._getframe(1).f_locals
            snaps[l] = (loc["gf_up"].copy(), loc["gf_dn"].copy())
            print(f"cfg5: slice {l} done after {time.time() - t0:.0f} s", flush=True)
            if l == lt - 2:
                raise Stop

    solver._debug = hook
    real_inv = np.linalg.inv
    pending = [g0[0].astype(np.float64), g0[1].astype(np.float64)]

    def inv_once(a):
        if pending and a.shape == (n, n):
            return pending.pop(0).copy()
        return real_inv(a)

    real_get_m = solver.get_m
    solver.get_m = lambda l0, sigma: np.eye(n)            # the product is not needed: its inverse is replaced
    np.linalg.inv = inv_once
    t0 = time.time()
    try:
        solver._update_step()
    except Stop:
        pass
    finally:
        np.linalg.inv = real_inv
        solver.get_m = real_get_m
    assert not pending
    print(f"cfg5: two reference slices took {time.time() - t0:.1f} s, accept {accs.mean():.3f}, "
          f"max|G| after 99: {np.abs(snaps[99][0]).max():.3e}")
    rows = np.arange(3, n, 36)
    cols = np.arange(7, n, 36)
    arrays = dict(ham=model.ham_kinetic(), u=6.0, beta=10.0, lamb=solver.lamb, exp_k=solver.exp_k, rows=rows, cols=cols,
                  g0_up=g0[0], g0_dn=g0[1], field0=field_in, field1=solver.config.config.copy(), uniforms=uniforms, ratios=ratios, accs=accs)
    for l in (99, 98):
        for tag, g in (("up", snaps[l][0]), ("dn", snaps[l][1])):
            gr, gc, gdg = subsample(g, rows, cols)
            arrays.update({f"post{l}_{tag}_rows": gr, f"post{l}_{tag}_cols": gc, f"post{l}_{tag}_diag": gdg})
    save("cfg5_24x24_slices", **arrays)


CASES = dict(cfg4mid=case_cfg4mid, cfg5=case_cfg5, cfg1=case_cfg1, small=case_small, cfg2=case_cfg2, cfg3=case_cfg3, cfg4=case_cfg4, u0=case_u0,
             lattice=case_lattice, det=case_det)

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None, choices=sorted(CASES))
    args = ap.parse_args()
    ref = import_reference()
    for key, fn in CASES.items():
        if args.only in (None, key):
            fn(ref)
