#!/usr/bin/env python
"""Where a CTA's time goes inside the one-launch sweep of the register kernel (N <= 64): clock64 stamps of thread 0
(library built with -DLQMC_PHASE_CLOCKS, LQMC_B200_LIB pointing at it).
usage: python tools/reg_split.py [workload] [chains] [arith]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import build_workload
from latticeqmc_b200 import SweepEngine
from latticeqmc_b200.workloads import synthetic_fields

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = build_workload(name)
chains = int(sys.argv[2]) if len(sys.argv) > 2 else w["chains"]
arith = sys.argv[3] if len(sys.argv) > 3 else "exact"
n, lt = w["n"], w["lt"]
eng = SweepEngine(w["exp_k"], w["lamb"], lt, n_chains=chains, exp_k_inv=w["exp_k_inv"], arith=arith)
eng.set_field(synthetic_fields(n, lt, chains))
eng.sweep(1, None, seed=1)
for rep in range(2):
    eng.reset_measurements()
    t0 = time.perf_counter(); eng.sweep(1, None, seed=2 + rep); dt = (time.perf_counter() - t0) * 1e3
    ob = eng.get_measurements()["obs_sum"][:, 0, :16]
    rec, sl, wr, scan, build, flush, flips, scans = ob.mean(0)[:8]
    fine = ob.mean(0)[8:] / max(flips, 1)
    print(f"{name} chains={chains} arith={arith}: sweep {dt:.3f} ms; per CTA (K clocks): recompute {rec/1e3:.0f}  slices {sl/1e3:.0f}  wraps {wr/1e3:.0f}"
          f"  | per flip (clocks): slice {sl/max(flips,1):.0f} = scan {scan/max(flips,1):.0f} + build {build/max(flips,1):.0f} + flush {flush/max(flips,1):.0f}"
          f"; flips/slice {flips/lt:.1f} scans/flip {scans/max(flips,1):.2f}; per wrap {wr/(lt-1):.0f}")
    print("   build, row thread: load %.0f chain %.0f [accept -> before barrier %.0f] barrier %.0f | column thread: load %.0f chain %.0f [to barrier %.0f] barrier %.0f" % tuple(fine))
