"""Synthetic workloads of BASELINE.json / SURVEY.md 8(d), built through the drop-in model API.

Used by bench.py and the tests' product side; deliberately independent of `oracle/`.
"""
import numpy as np
from scipy.linalg import expm

from .hubbard import HubbardModel


def build_model(kind, size, u, t=1.0, mu=None):
    """`kind`: 'square' (size x size, periodic in both directions) or 'ring' (size sites).  `mu=None` keeps the
    reference default mu = U/2 on the diagonal of K (hubbard.py:30,99-100); physics-mode workloads pass mu = 0,
    which is true half filling in the reference's HS convention."""
    model = HubbardModel(u=u, t=t, mu=mu)
    if kind == "square":
        model.build_square(size)
    elif kind == "ring":
        model.build(size)
    else:
        raise ValueError(f"unknown lattice kind {kind!r}")
    return model


def kinetic_and_constants(kind, size, u, beta, time_steps, mu=None):
    """`(K, dtau, lamb, exp_k, exp_k_inv)` the way `LatticeQMC.set_beta` derives them (lqmc.py:102-106)."""
    ham = build_model(kind, size, u, mu=mu).ham_kinetic()
    dtau = beta / time_steps
    lamb = np.arccosh(np.exp(u * dtau / 2.)) if u else 0
    return ham, dtau, lamb, expm(-1 * dtau * ham), expm(dtau * ham)


def synthetic_fields(n_sites, time_steps, n_chains, seed0=0):
    """Chain c starts from `2*RandomState(seed0+c).randint(0,2,(N,L))-1` - the generator of
    configuration.py:123-124 on a private stream."""
    out = np.empty((n_chains, n_sites, time_steps), dtype=np.int8)
    for c in range(n_chains):
        rs = np.random.RandomState(seed0 + c)
        out[c] = 2 * rs.randint(0, 2, size=(n_sites, time_steps)) - 1
    return out
