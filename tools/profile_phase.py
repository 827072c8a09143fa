#!/usr/bin/env python
"""Launch single phases through the C ABI so that ncu can capture them in isolation.
usage: python tools/profile_phase.py workload chains phase   (phase in slice|wrap|recompute)
Launch order: recompute, sweep, then 3 x phase  ->  profile with  -k regex:sweep -s 2 -c 1"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_workload
from latticeqmc_b200 import SweepEngine
from latticeqmc_b200.workloads import synthetic_fields
name, chains, phase = sys.argv[1], int(sys.argv[2]), sys.argv[3]
arith = sys.argv[4] if len(sys.argv) > 4 else "exact"
w = build_workload(name)
n, lt = w["n"], w["lt"]
eng = SweepEngine(w["exp_k"], w["lamb"], lt, n_chains=chains, exp_k_inv=w["exp_k_inv"], arith=arith)
eng.set_field(synthetic_fields(n, lt, chains))
eng.recompute(0)
eng.sweep(1, None, seed=1)
for r in range(3):
    l = lt - 1 - r
    if phase == "slice": eng.slice(l, None, seed=2)
    elif phase == "wrap": eng.wrap(l)
    else: eng.recompute(0)
