"""Debug aid: stabilised recompute on the GPU vs the oracle, error per case (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.linalg import expm
from oracle import sweep_oracle as so
from latticeqmc_b200 import SweepEngine

cases = [("square", 2, 4.0, 2.0, 20, 0.0, 0, 8), ("square", 4, 4.0, 4.0, 40, 0.0, 7, 10), ("square", 6, 6.0, 6.0, 60, 0.0, 59, 8),
         ("ring", 64, 8.0, 8.0, 80, 0.0, 0, 10), ("square", 10, 4.0, 4.0, 40, 0.0, 3, 8), ("square", 12, 4.0, 2.0, 20, 0.0, 19, 5),
         ("square", 16, 4.0, 8.0, 80, 0.0, 0, 8), ("ring", 64, 8.0, 8.0, 80, 4.0, 41, 10), ("ring", 64, 8.0, 8.0, 80, 4.0, 41, 5)]
if len(sys.argv) > 1:
    cases = [cases[int(a)] for a in sys.argv[1:]]
for case in cases:
    kind, size, u, beta, lt, mu, l0, chunk = case
    ham = so.ideal_square_kinetic(size, 1.0, mu) if kind == "square" else so.ideal_ring_kinetic(size, 1.0, mu)
    n = ham.shape[0]
    dtau, lamb, exp_k = so.set_beta_constants(ham, u, beta, lt)
    fields = np.stack([so.initial_field(n, lt, 700 + c) for c in range(2)])
    try:
        with SweepEngine(exp_k, lamb, lt, n_chains=2, exp_k_inv=expm(dtau * ham), mode="physics") as eng:
            eng.set_field(fields)
            t0 = time.time()
            eng.recompute_stable(l0, chunk)
            dt = time.time() - t0
            gg = eng.get_g()
        errs = []
        for c in range(2):
            for si, sigma in enumerate((+1, -1)):
                ref = so.physics_g_stable(fields[c], exp_k, lamb, l0, sigma, chunk)
                errs.append(float(np.abs(gg[c, si] - ref).max()))
        print("  max|G|", float(np.abs(gg).max()), "max|ref|", float(np.abs(ref).max()))
        print(case, "N", n, "errs", ["%.2e" % e for e in errs], "nan", int(np.isnan(gg).sum()), "t %.3fs" % dt, flush=True)
    except Exception as ex:
        print(case, "FAILED", repr(ex), flush=True)
