#!/usr/bin/env python
"""Throughput of cfg4 at small chain counts: one CTA per chain vs the automatic cluster choice.  usage: python tools/strong_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import build_workload
from latticeqmc_b200 import SweepEngine
from latticeqmc_b200.workloads import synthetic_fields
w = build_workload("cfg4"); n, lt = w["n"], w["lt"]
for chains in (9, 18, 24, 37, 49, 74):
    row = []
    for cs in ("1", ""):
        if cs: os.environ["LQMC_L2_CLUSTER"] = cs
        else: os.environ.pop("LQMC_L2_CLUSTER", None)
        with SweepEngine(w["exp_k"], w["lamb"], lt, n_chains=chains, exp_k_inv=w["exp_k_inv"]) as eng:
            eng.set_field(synthetic_fields(n, lt, chains))
            eng.sweep(1, None, seed=5)
            t0 = time.perf_counter(); eng.sweep(2, None, seed=5); dt = (time.perf_counter() - t0) / 2 * 1e3
            row.append((dt, chains * n * lt / dt * 1e3))
    print(f"chains {chains:4d}: one CTA per chain {row[0][0]:8.2f} ms ({row[0][1]:.3e} prop/s)   auto cluster {row[1][0]:8.2f} ms ({row[1][1]:.3e} prop/s)   x{row[0][0] / row[1][0]:.2f}", flush=True)
