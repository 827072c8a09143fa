import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from bench import build_workload
from latticeqmc_b200 import SweepEngine
from latticeqmc_b200.workloads import synthetic_fields
w = build_workload("cfg4"); n, lt = w["n"], w["lt"]; chains = 296
eng = SweepEngine(w["exp_k"], w["lamb"], lt, n_chains=chains, exp_k_inv=w["exp_k_inv"])
eng.set_field(synthetic_fields(n, lt, chains))
eng.sweep(2, None, seed=1)
for k in (1, 1, 5, 10):
    t0 = time.perf_counter(); eng.sweep(k, None, seed=3); dt = (time.perf_counter() - t0) * 1e3
    print(f"sweep({k}) in one launch: {dt / k:.1f} ms per sweep")
