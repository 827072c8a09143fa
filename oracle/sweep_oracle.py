"""CPU oracle for the HS-field Metropolis sweep - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy restatement of the reference algorithm ``LatticeQMC._update_step`` and its helpers
(`/root/reference/lqmc/lqmc.py:93-117,132-185,236-299,301-375`).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module, and only as the checker / the timed CPU baseline; the product path
(``latticeqmc_b200``) never does and fails loudly when its CUDA library is missing.

Pinning: the reference ships no tests or golden vectors, so this oracle is pinned
against outputs of the reference itself, generated in the build container by
``tests/golden/make_golden.py`` (which imports ``/root/reference`` with matplotlib stubbed) and
committed under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them and demands
bit-equality of the accept/reject trace, the ratios, the field and G.  Third-party arithmetic at
the boundary is NumPy/OpenBLAS ``dot``/``inv`` and SciPy ``expm`` (unpinned by the reference:
no requirements file; numpy 2.3.5 / scipy 1.18.1 here) - the oracle calls the very same
functions in the same order, which is what makes bit-equality possible.

Two modes live here:

* **parity** (`update_step` and friends) follows the reference literally, including its known
  quirks (G(l0=0) used at every slice, opposite signs of ``arg`` in ratio and
  update, the ``V_{l-1} K`` similarity wrap).
* **physics** (`physics_*`) is the textbook DQMC recurrence in the reference's own ``get_m``
  convention; it is what exact diagonalisation is compared with.
"""
import numpy as np
from scipy.linalg import expm


# ----------------------------------------------------------------------------------------------
# constants  (reference: lqmc.py:93-117 ``set_beta``)
# ----------------------------------------------------------------------------------------------

def set_beta_constants(ham_kin, u, beta, time_steps):
    """``dtau, lamb, exp_k`` exactly as ``LatticeQMC.set_beta`` computes them (lqmc.py:102-106)."""
    dtau = beta / time_steps
    lamb = np.arccosh(np.exp(u * dtau / 2.)) if u else 0
    exp_k = expm(-1 * dtau * ham_kin)
    return dtau, lamb, exp_k


def ideal_square_kinetic(size, t=1.0, mu=2.0):
    """Periodic nearest-neighbour K on a ``size x size`` lattice, site = x*size + y, diagonal
    ``-mu`` - what ``HubbardModel.build_square(size).ham_kinetic()`` yields wherever the reference
    can build it (hubbard.py:91-111), and the stand-in where it cannot (16x16)."""
    n = size * size
    ham = -mu * np.eye(n, dtype=np.float64)
    for x in range(size):
        for y in range(size):
            i = x * size + y
            for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
                j = ((x + dx) % size) * size + (y + dy) % size
                if i != j:
                    ham[i, j] = -t
    return ham


def ideal_ring_kinetic(n, t=1.0, mu=2.0):
    """Periodic chain of ``n`` sites (``HubbardModel.build(n)``, hubbard.py:60-76)."""
    ham = -mu * np.eye(n, dtype=np.float64)
    for i in range(n):
        for j in ((i + 1) % n, (i - 1) % n):
            if i != j:
                ham[i, j] = -t
    return ham


# ----------------------------------------------------------------------------------------------
# parity mode: literal restatement
# ----------------------------------------------------------------------------------------------

def get_exp_v(h, l, sigma, lamb):
    """Dense ``diagflat(exp(-sigma*lamb*h[:, l]))`` (lqmc.py:132-154)."""
    diag = -1 * sigma * lamb * h[:, l]
    return np.diagflat(np.exp(diag))


def get_m(h, exp_k, lamb, l0, sigma):
    """``I + B_{idx[0]} B_{idx[1]} ...`` with the cyclic slice order of lqmc.py:175-185;
    ``l0 = 0`` gives ``L-1, ..., 0``.  Left-to-right accumulation starting from the scalar 1."""
    n, time_steps = h.shape
    l0 = l0 % time_steps
    indices = list(reversed(range(time_steps)))
    time_indices = indices[-l0:] + indices[:-l0]
    b_prod = 1
    for l in time_indices:
        exp_v = get_exp_v(h, l, sigma, lamb)
        b = np.dot(exp_k, exp_v)
        b_prod = np.dot(b_prod, b)
    return np.eye(n) + b_prod


def sweep_start_g(h, exp_k, lamb):
    """Unstabilised sweep-start Green's functions (lqmc.py:303-307)."""
    gf_up = np.linalg.inv(get_m(h, exp_k, lamb, 0, +1))
    gf_dn = np.linalg.inv(get_m(h, exp_k, lamb, 0, -1))
    return gf_up, gf_dn


def rank1_literal(gf, e, c):
    """The reference's interpreted element loop (lqmc.py:328-331): multiply, then subtract."""
    n = gf.shape[0]
    for j in range(n):
        for k in range(n):
            gf[j, k] = gf[j, k] - e[j] * c[k]


def slice_proposals(gf_up, gf_dn, h, l, lamb, uniforms=None, literal=False):
    """All N single-spin-flip proposals of time slice ``l`` (lqmc.py:311-335), in place.

    ``uniforms``: N numbers consumed one per proposal, accepted or not; ``None`` draws
    ``np.random.rand()`` from the global legacy stream exactly like the reference (lqmc.py:317).
    Returns ``(ratios, accs)``.  ``literal=True`` runs the interpreted rank-1 loop; the default
    ``G - outer(e, c)`` is bit-equal to it (each element is one multiply and one subtract).
    """
    n = h.shape[0]
    ratios = np.empty(n, dtype=np.float64)
    accs = np.zeros(n, dtype=bool)
    for i in range(n):
        arg = 2 * lamb * h[i, l]
        d_up = 1 + (1 - gf_up[i, i]) * (np.exp(+arg) - 1)
        d_dn = 1 + (1 - gf_dn[i, i]) * (np.exp(-arg) - 1)
        ratio = d_up * d_dn
        u = np.random.rand() if uniforms is None else uniforms[i]
        acc = u <= ratio
        ratios[i] = ratio
        accs[i] = acc
        if acc:
            c_up = -(np.exp(-arg) - 1) * gf_up[i, :]
            c_up[i] += (np.exp(-arg) - 1)
            c_dn = -(np.exp(+arg) - 1) * gf_dn[i, :]
            c_dn[i] += (np.exp(+arg) - 1)
            e_up = gf_up[:, i] / (1 + c_up[i])
            e_dn = gf_dn[:, i] / (1 + c_dn[i])
            if literal:
                rank1_literal(gf_up, e_up, c_up)
                rank1_literal(gf_dn, e_dn, c_dn)
            else:
                gf_up -= np.outer(e_up, c_up)
                gf_dn -= np.outer(e_dn, c_dn)
            h[i, l] *= -1
    return ratios, accs


def wrap(gf_up, gf_dn, h, l, exp_k, lamb):
    """Propagate from slice ``l`` to ``l-1`` the way the reference does (lqmc.py:338-345):
    ``b = diagflat(v(l-1)) @ exp_k``; ``G <- (b @ G) @ inv(b)``.  Returns new arrays."""
    b_up = np.dot(get_exp_v(h, l - 1, +1, lamb), exp_k)
    b_dn = np.dot(get_exp_v(h, l - 1, -1, lamb), exp_k)
    gf_up = np.dot(np.dot(b_up, gf_up), np.linalg.inv(b_up))
    gf_dn = np.dot(np.dot(b_dn, gf_dn), np.linalg.inv(b_dn))
    return gf_up, gf_dn


def update_step(h, exp_k, lamb, uniforms=None, literal=False, snapshots=None, g_start=None):
    """One full sweep = ``LatticeQMC._update_step`` (lqmc.py:301-347).  ``h`` (int8 ``(N, L)``) is
    mutated in place.

    ``uniforms``: ``(L, N)`` array in visiting order (row 0 belongs to slice ``L-1``) or ``None``
    for the global stream.  ``snapshots``: optional dict filled with ``('post', l) -> (G_up, G_dn)``
    copies taken after the proposals of slice ``l`` (before the wrap).  ``g_start``: teacher-forced
    sweep-start ``(G_up, G_dn)`` replacing the recompute.
    Returns ``gf_up, gf_dn, ratios (L, N), accs (L, N)`` (trace rows in visiting order).
    """
    n, time_steps = h.shape
    if g_start is None:
        gf_up, gf_dn = sweep_start_g(h, exp_k, lamb)
    else:
        gf_up, gf_dn = np.array(g_start[0], dtype=np.float64), np.array(g_start[1], dtype=np.float64)
    ratios = np.empty((time_steps, n), dtype=np.float64)
    accs = np.zeros((time_steps, n), dtype=bool)
    for step, l in enumerate(reversed(range(time_steps))):
        u = None if uniforms is None else uniforms[step]
        ratios[step], accs[step] = slice_proposals(gf_up, gf_dn, h, l, lamb, u, literal)
        if snapshots is not None:
            snapshots[("post", l)] = (gf_up.copy(), gf_dn.copy())
        if l > 0:
            gf_up, gf_dn = wrap(gf_up, gf_dn, h, l, exp_k, lamb)
    return gf_up, gf_dn, ratios, accs


def measure_loop(h, exp_k, lamb, sweeps, uniforms=None):
    """``LatticeQMC.measure_loop`` (lqmc.py:356-375): mean of the end-of-sweep G over sweeps."""
    n = h.shape[0]
    tot_up = np.zeros((n, n), dtype=np.float64)
    tot_dn = np.zeros((n, n), dtype=np.float64)
    for s in range(sweeps):
        u = None if uniforms is None else uniforms[s]
        gf_up, gf_dn, _, _ = update_step(h, exp_k, lamb, u)
        tot_up += gf_up
        tot_dn += gf_dn
    return np.asarray([tot_up, tot_dn]) / sweeps


# ----------------------------------------------------------------------------------------------
# det mode: the reference's slow validation sampler (lqmc.py:236-299), literal restatement
# ----------------------------------------------------------------------------------------------

def det_product(h, exp_k, lamb):
    """``det(M_up(0)) * det(M_dn(0))``: how ``warmup_loop_det`` / ``measure_loop_det`` initialise
    ``old_det`` (lqmc.py:264-268, 283-287)."""
    return np.linalg.det(get_m(h, exp_k, lamb, 0, +1)) * np.linalg.det(get_m(h, exp_k, lamb, 0, -1))


def det_update_step(h, exp_k, lamb, old_det, uniforms=None):
    """One det-mode sweep = ``LatticeQMC._update_step_det`` (lqmc.py:236-259): flip, rebuild
    ``get_m(l, +-1)`` from the field, ``ratio = det(M_up) det(M_dn) / old_det``, Metropolis test
    ``u <= ratio``, un-flip on reject.  ``h`` is mutated in place.  ``uniforms``: ``(L, N)`` in visiting
    order or ``None`` for the global stream.  Returns ``old_det, ratios (L, N), accs (L, N)``."""
    n, time_steps = h.shape
    ratios = np.empty((time_steps, n), dtype=np.float64)
    accs = np.zeros((time_steps, n), dtype=bool)
    for step, l in enumerate(reversed(range(time_steps))):
        for i in range(n):
            h[i, l] *= -1
            m_up = get_m(h, exp_k, lamb, l, +1)
            m_dn = get_m(h, exp_k, lamb, l, -1)
            new_det = np.linalg.det(m_up) * np.linalg.det(m_dn)
            ratio = new_det / old_det
            u = np.random.rand() if uniforms is None else uniforms[step, i]
            acc = u <= ratio
            if acc:
                old_det = new_det
            else:
                h[i, l] *= -1
            ratios[step, i] = ratio
            accs[step, i] = acc
    return old_det, ratios, accs


def det_measure_loop(h, exp_k, lamb, sweeps, uniforms=None, trace=None):
    """``LatticeQMC.measure_loop_det`` (lqmc.py:274-299): ``old_det`` from ``get_m(0, +-1)`` once, then
    per sweep one ``_update_step_det`` and ``inv(get_m(0, +-1))`` added to the totals.  ``trace``:
    optional list receiving ``(ratios, accs)`` per sweep."""
    n = h.shape[0]
    old_det = det_product(h, exp_k, lamb)
    tot_up = np.zeros((n, n), dtype=np.float64)
    tot_dn = np.zeros((n, n), dtype=np.float64)
    for s in range(sweeps):
        u = None if uniforms is None else uniforms[s]
        old_det, ratios, accs = det_update_step(h, exp_k, lamb, old_det, u)
        if trace is not None:
            trace.append((ratios, accs))
        tot_up += np.linalg.inv(get_m(h, exp_k, lamb, 0, +1))
        tot_dn += np.linalg.inv(get_m(h, exp_k, lamb, 0, -1))
    return np.asarray([tot_up, tot_dn]) / sweeps


def initial_field(n_sites, time_steps, seed):
    """Benchmark / test field: ``2*RandomState(seed).randint(0,2,(N,L))-1`` as int8 - the same
    generator the reference uses on the global stream (configuration.py:123-124)."""
    rs = np.random.RandomState(seed)
    return (2 * rs.randint(0, 2, size=(n_sites, time_steps)) - 1).astype(np.int8)


# ----------------------------------------------------------------------------------------------
# physics mode: correct DQMC in the get_m convention
# ----------------------------------------------------------------------------------------------

def physics_b(h, l, sigma, exp_k, lamb):
    """``B_l = exp_k @ diag(exp(-sigma*lamb*h[:, l]))`` (the factor get_m multiplies, lqmc.py:182)."""
    return exp_k * np.exp(-sigma * lamb * h[:, l])[None, :]


def physics_g_naive(h, exp_k, lamb, l0, sigma):
    """``G(l0) = inv(get_m(l0))`` - unstabilised; fine for the small beta used in tests."""
    return np.linalg.inv(get_m(h, exp_k, lamb, l0, sigma))


def physics_g_stable(h, exp_k, lamb, l0, sigma, stab_every=8):
    """``G(l0)`` through a QR/UDV-accumulated product with column-norm pre-pivoting:
    ``A = U D V``, ``G = (D_b^-1 U^T + D_s V)^-1 D_b^-1 U^T`` with ``D = D_b D_s`` split at 1."""
    n, time_steps = h.shape
    l0 = l0 % time_steps
    indices = list(reversed(range(time_steps)))
    order = indices[-l0:] + indices[:-l0]          # leftmost factor first
    u = np.eye(n)
    d = np.ones(n)
    v = np.eye(n)
    chunk = np.eye(n)
    count = 0
    # A = B_{order[0]} ... B_{order[-1]}: absorb factors from the right end towards the left
    for l in reversed(order):
        chunk = physics_b(h, l, sigma, exp_k, lamb) @ chunk
        count += 1
        if count == stab_every or l == order[0]:
            m = (chunk @ u) * d[None, :]
            norms = np.linalg.norm(m, axis=0)
            perm = np.argsort(-norms)
            q, r = np.linalg.qr(m[:, perm])
            d_new = np.abs(np.diag(r))
            d_new[d_new == 0] = 1e-300
            t = (r / d_new[:, None])
            inv_perm = np.empty(n, dtype=int)
            inv_perm[perm] = np.arange(n)
            v = t[:, inv_perm] @ v
            u, d = q, d_new
            chunk = np.eye(n)
            count = 0
    d_b = np.maximum(d, 1.0)
    d_s = np.minimum(d, 1.0)
    lhs = (u.T / d_b[:, None]) + d_s[:, None] * v
    return np.linalg.solve(lhs, u.T / d_b[:, None])


def physics_sweep(h, exp_k, exp_k_inv, lamb, uniforms, stab_every=0, observables=None):
    """One correct DQMC sweep, slices ``L-1 .. 0``, sites ``0 .. N-1`` (same visiting order and
    one uniform per proposal as the reference).  ``stab_every = k > 0`` recomputes G with
    `physics_g_stable` before every k-th slice; 0 only at sweep start.  Returns
    ``(G_up(0), G_dn(0), ratios, accs)``; `observables` (a dict) accumulates equal-time sums.
    """
    n, time_steps = h.shape
    ratios = np.empty((time_steps, n))
    accs = np.zeros((time_steps, n), dtype=bool)
    g = {}
    for step, l in enumerate(reversed(range(time_steps))):
        fresh = step == 0 or (stab_every and step % stab_every == 0)
        for sigma in (+1, -1):
            if fresh:
                g[sigma] = physics_g_stable(h, exp_k, lamb, l, sigma, stab_every or 8)
            else:
                # G(l) = B_l^-1 G(l+1) B_l
                v = np.exp(-sigma * lamb * h[:, l])
                g[sigma] = ((exp_k_inv @ g[sigma] @ exp_k) * v[None, :]) / v[:, None]
        for i in range(n):
            delta = {s: np.exp(2 * s * lamb * h[i, l]) - 1 for s in (+1, -1)}
            r = {s: 1 + (1 - g[s][i, i]) * delta[s] for s in (+1, -1)}
            ratio = r[+1] * r[-1]
            acc = uniforms[step, i] <= ratio
            ratios[step, i] = ratio
            accs[step, i] = acc
            if acc:
                for s in (+1, -1):
                    col = -g[s][:, i].copy()
                    col[i] += 1.0
                    g[s] -= np.outer(col, g[s][i, :]) * (delta[s] / r[s])
                h[i, l] *= -1
    if observables is not None:
        n_up = 1 - np.diag(g[+1])
        n_dn = 1 - np.diag(g[-1])
        observables["n_up"] = observables.get("n_up", 0.0) + n_up.mean()
        observables["n_dn"] = observables.get("n_dn", 0.0) + n_dn.mean()
        observables["docc"] = observables.get("docc", 0.0) + (n_up * n_dn).mean()
        observables["count"] = observables.get("count", 0) + 1
    return g[+1], g[-1], ratios, accs
