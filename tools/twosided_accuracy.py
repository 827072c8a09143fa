"""Accuracy of the two-sided UDV combination (stab.cuh / DESIGN.md 5.4) against an 80-100 digit mpmath product.
NumPy prototype of what the kernels do: left stack, running right product, Loh combination ("loh"), an alternative
that re-factorises D_L (V_L V_R^T) D_R ("qr2"), and the oracle one-sided scheme; 4x4 lattices where mpmath is cheap.
TEST INFRASTRUCTURE (imports oracle/); run:  python tools/twosided_accuracy.py"""
import sys; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, mpmath as mp
from oracle import sweep_oracle as so
mp.mp.dps = 100
def absorb(M):
    n = M.shape[0]
    norms = np.linalg.norm(M, axis=0); perm = np.argsort(-norms, kind='stable')
    q, r = np.linalg.qr(M[:, perm]); d = np.abs(np.diag(r)); t = r / d[:, None]
    inv = np.empty(n, int); inv[perm] = np.arange(n)
    return q, d, t[:, inv]
def Bm(h, l, sigma, E, lamb): return E * np.exp(-sigma*lamb*h[:, l])[None, :]
def run(n_side, U_, beta, L, k, seed=3):
    n = n_side*n_side
    ham = so.ideal_square_kinetic(n_side, 1.0, 0.0); dtau, lamb, E = so.set_beta_constants(ham, U_, beta, L)
    h = so.initial_field(n, L, seed); sigma = +1
    def g_exact(h, l0):
        order = list(reversed(range(L))); order = order[-l0:] + order[:-l0] if l0 else order
        A = mp.eye(n)
        for l in order: A = A * mp.matrix(Bm(h, l, sigma, E, lamb).tolist())
        return np.array(((mp.eye(n) + A)**-1).tolist(), dtype=float)
    nseg = (L + k - 1)//k
    seg = [(max(0, L-(j+1)*k), L-1-j*k) for j in range(nseg)]
    stack = [None]*nseg
    U, D, V = np.eye(n), np.ones(n), np.eye(n)
    for j in reversed(range(nseg)):
        lo, hi = seg[j]; M = U.copy()
        for l in range(lo, hi+1): M = Bm(h, l, sigma, E, lamb) @ M
        U, D, T = absorb(M * D[None, :]); V = T @ V
        stack[j] = (U, D, V)
    Ur, Dr, Vr = np.eye(n), np.ones(n), np.eye(n)
    rng = np.random.RandomState(0)
    for j in range(nseg):
        if j > 0:
            lo, hi = seg[j-1]; M = Ur.copy()
            for l in range(hi, lo-1, -1): M = Bm(h, l, sigma, E, lamb).T @ M
            Ur, Dr, T = absorb(M * Dr[None, :]); Vr = T @ Vr
        UL, DL, VL = stack[j]
        DLb, DLs, DRb, DRs = np.maximum(DL,1), np.minimum(DL,1), np.maximum(Dr,1), np.minimum(Dr,1)
        innerT = (Ur.T @ UL) / DRb[:, None] / DLb[None, :] + ((VL @ Vr.T) * DLs[:, None] * DRs[None, :]).T
        G1 = ((np.linalg.inv(innerT).T @ (UL.T / DLb[:, None])).T @ (Ur.T / DRb[:, None])).T
        # method 2: one more QR on D_L (V_L Vr^T) D_R
        M = (VL @ Vr.T) * DL[:, None] * Dr[None, :]
        Q, Dm, Tm = absorb(M)
        Uc = UL @ Q; Vc = Tm @ Ur.T
        Db, Ds = np.maximum(Dm, 1), np.minimum(Dm, 1)
        lhs = Uc.T / Db[:, None] + Ds[:, None] * Vc
        G2 = np.linalg.solve(lhs, Uc.T / Db[:, None])
        l0 = (seg[j][1] + 1) % L
        ref = so.physics_g_stable(h, E, lamb, l0, sigma, k)
        ge = g_exact(h, l0)
        print(j, l0, "loh %.1e  qr2 %.1e  oracle %.1e   cond(innerT) %.1e cond(VL) %.1e cond(Vr) %.1e" % (np.abs(G1-ge).max(), np.abs(G2-ge).max(), np.abs(ref-ge).max(),
              np.linalg.cond(innerT), np.linalg.cond(VL), np.linalg.cond(Vr)))
        lo, hi = seg[j]
        flips = rng.rand(n, hi-lo+1) < 0.5
        h[:, lo:hi+1] = np.where(flips, -h[:, lo:hi+1], h[:, lo:hi+1])
run(4, 6.0, 8.0, 40, 8)
run(4, 8.0, 8.0, 80, 10)
