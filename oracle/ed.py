"""Dense exact diagonalisation of small Hubbard clusters - TEST INFRASTRUCTURE (oracle), not product code.

The reference has no interacting known answer (`exact.py` is the U=0 pole Green's function), so the
physics-mode gate of BASELINE.json ("density, double occupancy and local moment agree with ED") needs its own
ED: H = sum_{ij,s} K_ij c+_is c_js + U sum_i n_iu n_id - mu sum_i n_i on N <= 6 sites (4^N states), grand
canonical averages at inverse temperature beta.
"""
import numpy as np


def _ops(n_sites):
    """Jordan-Wigner annihilation operators for 2*n_sites spin-orbitals (orbital index = 2*site + spin)."""
    m = 2 * n_sites
    dim = 1 << m
    ops = []
    states = np.arange(dim)
    for orb in range(m):
        mat = np.zeros((dim, dim))
        occ = (states >> orb) & 1
        sign = (-1) ** np.array([bin(s & ((1 << orb) - 1)).count("1") for s in states])
        src = states[occ == 1]
        mat[src ^ (1 << orb), src] = sign[occ == 1]
        ops.append(mat)
    return ops


def thermal_observables(hop, u, mu, beta):
    """`hop`: (N, N) hopping matrix with zero diagonal (entries -t on bonds).  Returns dict with per-site
    averages n_up, n_dn, docc = <n_up n_dn>, moment = <(n_up - n_dn)^2>."""
    n = hop.shape[0]
    c = _ops(n)
    dim = 1 << (2 * n)
    num = [op.T @ op for op in c]
    ham = np.zeros((dim, dim))
    for i in range(n):
        for j in range(n):
            if i != j and hop[i, j] != 0.0:
                for s in range(2):
                    ham += hop[i, j] * (c[2 * i + s].T @ c[2 * j + s])
        ham += u * (num[2 * i] @ num[2 * i + 1]) - mu * (num[2 * i] + num[2 * i + 1])
    w, v = np.linalg.eigh(ham)
    p = np.exp(-beta * (w - w.min()))
    p /= p.sum()

    def avg(op):
        return float(np.einsum("k,ik,ij,jk->", p, v, op, v))

    n_up = np.mean([avg(num[2 * i]) for i in range(n)])
    n_dn = np.mean([avg(num[2 * i + 1]) for i in range(n)])
    docc = np.mean([avg(num[2 * i] @ num[2 * i + 1]) for i in range(n)])
    return dict(n_up=n_up, n_dn=n_dn, docc=docc, moment=n_up + n_dn - 2 * docc)
