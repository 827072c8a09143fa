#!/usr/bin/env python
"""clock64 split of the large-lattice slice phase (scan / build / flush) - needs a library built with
-DLQMC_PHASE_CLOCKS:  nvcc ... -DLQMC_PHASE_CLOCKS -o latticeqmc_b200/liblqmc_b200_clk.so
usage: python tools/phase_clocks.py [workload] [chains] [arith]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from latticeqmc_b200 import engine
engine._lib = engine.load_library(os.path.join(os.path.dirname(engine.LIB_PATH), "liblqmc_b200_clk.so"))
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
for l in (lt - 1, lt - 2, lt - 3):
    eng.reset_measurements()
    eng.slice(l, None, seed=2)
    ob = eng.get_measurements()["obs_sum"][:, 0, :7]
    scan, build, flush, nacc, b1, b2, b3 = ob.mean(0)
    full = eng.get_measurements()["obs_sum"][:, 0, :7]
    if full[:, 5].max() > 1e6:          # tensor-memory path: columns 4..6 = SM id, start, end (globaltimer ns)
        sm_id, t0, t1 = full[:, 4], full[:, 5], full[:, 6]
        pairs = ov = 0
        for a in range(chains):
            for b2 in range(a + 1, chains):
                if sm_id[a] == sm_id[b2]:
                    pairs += 1
                    ov += (t0[a] < t1[b2]) and (t0[b2] < t1[a])
        print(f"  CTAs sharing an SM: {pairs} pairs, {ov} overlapping in time; kernel span {(t1.max() - t0.min()) / 1e3:.0f} us, mean CTA {np.mean(t1 - t0) / 1e3:.0f} us")
    print(f"slice {l}: per CTA clocks scan {scan:9.0f} build {build:9.0f} flush {flush:9.0f} total {scan+build+flush:9.0f}  accepted {nacc:6.1f}"
          f"  per flip: scan {scan/nacc:7.0f} build {build/nacc:7.0f} flush {flush/nacc:7.0f}   build split: loads {b1/nacc:6.0f} apply(2 spins) {b2/nacc:6.0f} vectors {b3/nacc:6.0f}")
