#!/usr/bin/env python
"""One chain per thread-block cluster vs one chain per CTA: bit-identical results (field, G, accumulators, decisions) and timing.
usage: python tools/cluster_check.py [chains] [arith]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import build_workload
from latticeqmc_b200 import SweepEngine
from latticeqmc_b200.workloads import synthetic_fields

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 4
arith = sys.argv[2] if len(sys.argv) > 2 else "exact"
w = build_workload("cfg4"); n, lt = w["n"], w["lt"]
fields = synthetic_fields(n, lt, chains)
res = {}
for cs in ("1", "2", "3", "5", "6", "7", "8", ""):
    if cs: os.environ["LQMC_L2_CLUSTER"] = cs
    else: os.environ.pop("LQMC_L2_CLUSTER", None)
    with SweepEngine(w["exp_k"], w["lamb"], lt, n_chains=chains, exp_k_inv=w["exp_k_inv"], arith=arith, trace=True) as eng:
        eng.set_field(fields)
        eng.sweep(1, None, seed=5, measure=True)          # warm-up + first result
        t0 = time.perf_counter(); eng.sweep(2, None, seed=5, measure=True); dt = (time.perf_counter() - t0) / 2 * 1e3
        acc, ratio = eng.get_trace()
        res[cs] = (eng.get_field().copy(), eng.get_g().copy(), eng.get_measurements(), acc.copy(), ratio.copy())
        print(f"cluster={cs or 'auto':>4}: {dt:8.2f} ms per sweep, {chains * n * lt / dt * 1e3:.3e} proposals/s", flush=True)
ref = res["1"]
for cs, r in res.items():
    same = (np.array_equal(r[0], ref[0]) and np.array_equal(r[1], ref[1]) and np.array_equal(r[3], ref[3]) and np.array_equal(r[4], ref[4])
            and all(np.array_equal(r[2][k], ref[2][k]) for k in ("g_sum", "obs_sum", "n_meas", "n_accepted")))
    print(f"cluster={cs or 'auto':>4}: bit-identical to one CTA per chain: {same}")
    assert same
