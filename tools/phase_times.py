#!/usr/bin/env python
"""Wall-clock of the three phases of a sweep, each through its own C-ABI call (every call synchronises).
usage: python tools/phase_times.py [workload] [chains]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import build_workload, flops_per_sweep
from latticeqmc_b200 import SweepEngine
from latticeqmc_b200.workloads import synthetic_fields

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
w = build_workload(name)
chains = int(sys.argv[2]) if len(sys.argv) > 2 else w["chains"]
arith = sys.argv[3] if len(sys.argv) > 3 else "exact"
n, lt = w["n"], w["lt"]
eng = SweepEngine(w["exp_k"], w["lamb"], lt, n_chains=chains, exp_k_inv=w["exp_k_inv"], arith=arith)
eng.set_field(synthetic_fields(n, lt, chains))
def t(fn, reps=1):
    fn(); eng.sync()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    eng.sync()
    return (time.perf_counter() - t0) / reps * 1e3
rec = t(lambda: eng.recompute(0))
eng.sweep(1, None, seed=1)          # realistic G / field
ts, tw = [], []
for l in range(lt - 1, lt - 6, -1):
    ts.append(t(lambda: eng.slice(l, None, seed=2)))
    tw.append(t(lambda: eng.wrap(l)))
m0 = eng.get_measurements()["n_accepted"].sum()
eng.slice(lt - 7, None, seed=3)
acc = (eng.get_measurements()["n_accepted"].sum() - m0) / (chains * n)
sweep = t(lambda: eng.sweep(1, None, seed=4))
print(f"{name} chains={chains} N={n} L={lt} arith={arith}")
print(f"recompute  {rec:9.3f} ms   ({2 * (lt * 2 * n**3 + 2 * n**3) * chains / rec / 1e9:7.2f} TFLOP/s)")
print(f"slice      {np.mean(ts):9.3f} ms   (accept ~{acc:.2f}; rank-1 {acc * n * 4 * n * n * chains / np.mean(ts) / 1e9:7.2f} TFLOP/s)  {ts}")
print(f"wrap       {np.mean(tw):9.3f} ms   ({8 * n**3 * chains / np.mean(tw) / 1e9:7.2f} TFLOP/s)  {tw}")
print(f"sweep      {sweep:9.3f} ms   vs sum of phases {rec + lt * np.mean(ts) + (lt - 1) * np.mean(tw):9.3f} ms")
