"""Post-processing of measured Green's functions and the U=0 pole Green's function.

API-compatible with the reference ``lqmc/tools.py:12-200``.  Everything here works on the
``(2, N, N)`` result after the run; nothing is on the sweep path.  ``pole_gf_tau`` /
``compute_pole_gf_tau`` are the reference's only analytic known answer (the non-interacting
G(tau), ``exact.py:27-54``) and back the U=0 parity test.

Note: ``local_moment`` factorises <n_up n_dn> from the *averaged* G, as the
reference does (``tools.py:107-109``).  The correct per-configuration estimator is accumulated
on the device; see ``SweepEngine.get_observables``.
"""
import os

import numpy as np


def get_datapath(filename, model, post="", mkdir=True, **kwargs):
    w, h = model.lattice.shape
    folder = os.path.join("data", f"u={model.u}_t={model.t}_mu={model.mu}_w={w}_h={h}")
    if mkdir:
        os.makedirs(folder, exist_ok=True)
    tags = "_".join(f"{key}={val}" for key, val in kwargs.items())
    return os.path.abspath(os.path.join(folder, f"{filename} {tags}{post}.npz"))


def check_params(u, t, dtau):
    """Trotter-error check ``U t dtau^2 < 0.1`` (printed, not enforced)."""
    value = u * t * dtau ** 2
    if value < 0.1:
        print(f"Check-value {value:.2} is smaller than 0.1!")
    else:
        print(f"Check-value {value:.2} should be smaller than 0.1!")


def filling(g_sigma, site=None, axis1=-2, axis2=-1):
    """Local filling ``n_i = 1 - G_ii`` of one spin channel."""
    n = 1 - np.diagonal(g_sigma, axis1=axis1, axis2=axis2)
    return n if site is None else n[..., site]


def local_moment(gf_up, gf_dn, site=None, axis1=-2, axis2=-1):
    """``n_up + n_dn - 2 n_up n_dn`` from the averaged Green's functions."""
    n_up = filling(gf_up, site, axis1, axis2)
    n_dn = filling(gf_dn, site, axis1, axis2)
    return n_up + n_dn - 2 * n_up * n_dn


def local_gf(gf):
    return np.diagonal(gf, axis1=-2, axis2=-1)


def matsubara_frequencies(points, beta):
    n = np.asanyarray(points).astype(dtype=int, casting="safe")
    return 1j * np.pi / beta * (2 * n + 1)


def fermi_fct(eps, beta):
    """``1/(exp(beta*eps)+1)`` in the overflow-safe tanh form."""
    return 0.5 * (1. + np.tanh(-0.5 * beta * eps))


def decompose(a):
    xi, rv = np.linalg.eigh(a)
    return rv, xi, np.linalg.inv(rv)


def reconstruct(rv, xi, rv_inv, diag=False):
    if diag:
        return ((np.transpose(rv_inv) * rv) @ xi[..., np.newaxis])[..., 0]
    return (rv * xi[..., np.newaxis, :]) @ rv_inv


def pole_gf_tau(tau, poles, weights, beta):
    """Imaginary-time Green's function of a sum of poles, ``tau`` in ``[0, beta]``."""
    assert np.all((tau >= 0.) & (tau <= beta))
    poles, weights = np.atleast_1d(*np.broadcast_arrays(poles, weights))
    tau = np.asanyarray(tau)
    tau = tau.reshape(tau.shape + (1,) * poles.ndim)
    # exp(-tau*eps) f(-eps) == exp((beta-tau)*eps) f(eps): pick the branch that cannot overflow
    exponent = np.where(poles.real >= 0, -tau, beta - tau) * poles
    per_pole = np.exp(exponent) * fermi_fct(-np.sign(poles.real) * poles, beta)
    return -np.sum(weights * per_pole, axis=-1)


def compute_pole_gf_tau(ham, beta):
    """Non-interacting ``G_ij(tau)`` on 2049 tau points; returns ``(tau, gf[site, site, tau])``."""
    rv, xi, rv_inv = decompose(ham)
    tau = np.linspace(0, beta, num=2049)
    diag_gf = pole_gf_tau(tau, xi[..., np.newaxis], weights=1, beta=beta)
    gf = reconstruct(rv, diag_gf, rv_inv)
    return tau, np.moveaxis(gf, 0, -1)
