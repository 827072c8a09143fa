"""Beta scan (comp.py:20-46 shape: 4x4, L=50, 20 temperatures) through SerialProcessManager: all points in flight at
once against one point at a time.  Usage: python tools/beta_scan_timing.py [warmup] [sweeps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lqmc  # noqa: E402

warm = int(sys.argv[1]) if len(sys.argv) > 1 else 50
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 500
model = lqmc.HubbardModel(u=4, t=1)
model.build_square(4)
betas = list(1.0 / np.linspace(0.1, 5.0, 20))
out = {}
for label, conc in (("all points concurrent", None), ("one point at a time", 1)):
    for rep in range(2):           # first repetition warms up the context
        mgr = lqmc.SerialProcessManager(model, 50, warm, sweeps, caching=False, concurrent=conc, rng="philox")
        mgr.set_jobs(betas)
        t0 = time.time()
        mgr.run()
        dt = time.time() - t0
    out[label] = (dt, mgr.get_result())
    print(f"\n{label}: {dt:.3f} s for {len(betas)} betas x ({warm} + {sweeps}) sweeps of 16 x 50 proposals")
a, b = out["all points concurrent"], out["one point at a time"]
print("identical results:", np.array_equal(a[1], b[1]), " speed-up:", b[0] / a[0])
