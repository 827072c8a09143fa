"""Post-processing of measured Green's functions and the U=0 pole Green's function.

API-compatible with the reference ``lqmc/tools.py:12-200``.  Everything here works on the
``(2, N, N)`` result after the run; nothing is on the sweep path.  ``pole_gf_tau`` /
``compute_pole_gf_tau`` are the reference's only analytic known answer (the non-interacting
G(tau), ``exact.py:27-54``) and back the U=0 parity test.

Note: ``local_moment`` factorises <n_up n_dn> from the *averaged* G, as the
reference does (``tools.py:107-109``).  The correct per-configuration estimator is accumulated
on the device: ``SweepEngine.get_measurements()['obs_sum']`` and ``multiprocessing.chain_statistics``.
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


def chain_statistics(measurements):
    """Equal-time observables with error bars from independent chains.

    `measurements`: the dict of `SweepEngine.get_measurements()` (per-chain sums over measured sweeps of the site-resolved
    `n_up`, `n_dn` and the **per-configuration** product `n_up * n_dn`, which the device accumulates; the reference only
    averages G and factorises afterwards, `tools.py:107-109`, which is wrong for correlated quantities).  Chains are
    independent Markov chains, so the standard error of the mean over chains is an honest error bar, and a delete-one
    jackknife over chains gives the error of the non-linear local moment.  Returns
    `{name: (mean, stderr)}` for `n_up, n_dn, density, docc, moment`, plus `n_chains`, `n_meas`.
    """
    obs = np.asarray(measurements["obs_sum"], dtype=np.float64)            # (C, 3, N)
    n_meas = np.asarray(measurements["n_meas"], dtype=np.float64)
    keep = n_meas > 0
    if not keep.any():
        raise ValueError("no measured sweeps")
    per_chain = obs[keep].mean(axis=2) / n_meas[keep, None]                  # (C', 3): site-averaged chain means
    c = per_chain.shape[0]

    def mean_err(x):
        return float(x.mean()), float(x.std(ddof=1) / np.sqrt(c)) if c > 1 else float("nan")

    n_up, n_dn, docc = per_chain[:, 0], per_chain[:, 1], per_chain[:, 2]
    out = dict(n_up=mean_err(n_up), n_dn=mean_err(n_dn), density=mean_err(n_up + n_dn), docc=mean_err(docc))
    moment = n_up + n_dn - 2 * docc
    if c > 1:
        total = moment.sum()
        jack = (total - moment) / (c - 1)                                    # delete-one means
        out["moment"] = (float(moment.mean()), float(np.sqrt((c - 1) / c * np.sum((jack - jack.mean()) ** 2))))
    else:
        out["moment"] = (float(moment.mean()), float("nan"))
    out["n_chains"] = int(c)
    out["n_meas"] = int(n_meas[keep].sum())
    return out
