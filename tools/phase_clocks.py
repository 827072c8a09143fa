#!/usr/bin/env python
"""clock64 split of the large-lattice slice phase (scan / build / flush) - needs a library built with
-DLQMC_PHASE_CLOCKS (set LQMC_B200_LIB to it, or put it at latticeqmc_b200/liblqmc_b200_clk.so).
usage: python tools/phase_clocks.py [workload] [chains] [arith]
Columns written by thread 0 of every CTA into obs_sum (tensor-memory path, NP <= 256): scan, build, flush clocks, accepted flips,
clocks from the end of the scan to the start of the history application (b1), to its end (b2), flips that took the
owner-publish path (one more barrier), flips whose G0 row / column was not in a prefetch slot."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from latticeqmc_b200 import engine
if not os.environ.get("LQMC_B200_LIB"):
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
print(f"{name} chains={chains} arith={arith}")
for l in (lt - 1, lt - 2, lt - 3):
    eng.reset_measurements()
    eng.slice(l, None, seed=2)
    full = eng.get_measurements()["obs_sum"][:, 0, :15]
    ob = full[:, :8]
    scan, build, flush, nacc, b1, b2, nslow, nmiss = ob.mean(0)
    t0, t1, smid, clk = full[:, 8], full[:, 9], full[:, 10], full[:, 11]
    dur = t1 - t0
    print(f"  globaltimer: kernel span {(t1.max() - t0.min()) / 1e3:.0f} us, CTA mean {dur.mean() / 1e3:.0f} us, max {dur.max() / 1e3:.0f} us, start spread "
          f"{(t0.max() - t0.min()) / 1e3:.0f} us; SM clock seen {np.mean(clk / np.maximum(dur, 1)):.3f} GHz; SMs used {len(set(smid))}; accepted min/max {full[:, 3].min():.0f}/{full[:, 3].max():.0f}")
    print(f"slice {l}: per CTA clocks scan {scan:9.0f} build {build:9.0f} flush {flush:9.0f} total {scan+build+flush:9.0f}  accepted {nacc:6.1f}"
          f"  per flip: scan {scan/nacc:7.0f} build {build/nacc:7.0f} flush {flush/nacc:7.0f}   build split: to-apply {b1/nacc:6.0f} "
          f"loads-landed {full[:, 12].mean()/nacc:6.0f} apply-done {b2/nacc:6.0f} vectors {full[:, 13].mean()/nacc:6.0f} stored {full[:, 14].mean()/nacc:6.0f} barrier {build/nacc:6.0f}  owner-publish flips {nslow/nacc:.3f}  prefetch misses {nmiss/nacc:.3f}")
