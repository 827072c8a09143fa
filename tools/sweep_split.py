#!/usr/bin/env python
"""Where a CTA's time goes inside the ONE-launch sweep (recompute / slice phases / wraps), from clock64 stamps of thread 0
(library built with -DLQMC_PHASE_CLOCKS, LQMC_B200_LIB pointing at it).
usage: python tools/sweep_split.py [workload] [chains] [arith]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import build_workload
from latticeqmc_b200 import SweepEngine
from latticeqmc_b200.workloads import synthetic_fields

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
w = build_workload(name)
chains = int(sys.argv[2]) if len(sys.argv) > 2 else w["chains"]
arith = sys.argv[3] if len(sys.argv) > 3 else "exact"
n, lt = w["n"], w["lt"]
eng = SweepEngine(w["exp_k"], w["lamb"], lt, n_chains=chains, exp_k_inv=w["exp_k_inv"], arith=arith)
eng.set_field(synthetic_fields(n, lt, chains))
eng.sweep(1, None, seed=1)
import time
for rep in range(2):
    eng.reset_measurements()
    t0 = time.perf_counter(); eng.sweep(1, None, seed=2 + rep); dt = (time.perf_counter() - t0) * 1e3
    full = eng.get_measurements()["obs_sum"][:, 0, :19]
    tot = full[:, 16:19].sum(1) / 1.9e6
    smid = full[:, 10].astype(int)
    order = np.argsort(tot)
    print("  per-CTA total ms: min %.1f  p10 %.1f  median %.1f  p90 %.1f  max %.1f" % (tot.min(), np.percentile(tot, 10), np.median(tot), np.percentile(tot, 90), tot.max()))
    print("  slowest CTAs (chain, sm, ms, rec, slices, wraps):", [(int(c), int(smid[c]), round(float(tot[c]), 1), *(np.round(full[c, 16:19] / 1.9e6, 1))) for c in order[-6:]])
    print("  fastest CTAs:", [(int(c), int(smid[c]), round(float(tot[c]), 1), *(np.round(full[c, 16:19] / 1.9e6, 1))) for c in order[:6]])
    by_sm = {}
    for c in range(len(tot)): by_sm.setdefault(int(smid[c]), []).append(float(tot[c]))
    pair_tot = np.array([max(v) for v in by_sm.values()])
    print("  per-SM max: min %.1f median %.1f max %.1f; SMs %d" % (pair_tot.min(), np.median(pair_tot), pair_tot.max(), len(by_sm)))
    ob = full[:, 16:19]
    rec, sl, wr = ob.mean(0) / 1.9e6           # ms at ~1.9 GHz
    print(f"{name} chains={chains} arith={arith}: sweep {dt:.1f} ms; per CTA (ms at 1.9 GHz): recompute {rec:.1f}  slices {sl:.1f}  wraps {wr:.1f}  sum {rec+sl+wr:.1f}")
