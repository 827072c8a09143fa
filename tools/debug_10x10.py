import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from oracle import sweep_oracle as so
from latticeqmc_b200 import SweepEngine
ham = so.ideal_square_kinetic(10, 1.0, 2.0)
n, lt = 100, 10
dtau, lamb, exp_k = so.set_beta_constants(ham, 4.0, 1.0, lt)
fields = np.stack([so.initial_field(n, lt, seed=300 + c) for c in range(3)])
uni = np.random.RandomState(17).rand(3, 2, lt, n)
with SweepEngine(exp_k, lamb, lt, n_chains=3, trace=True) as eng:
    eng.set_field(fields)
    eng.sweep(2, uni, measure=True)
    acc, ratio = eng.get_trace()
    gg, ff, m = eng.get_g(), eng.get_field(), eng.get_measurements()
for c in range(3):
    h = fields[c].copy(); tot = np.zeros((2, n, n)); ends = []
    for s in range(2):
        gu, gd, r, a = so.update_step(h, exp_k, lamb, uni[c, s]); ends.append(np.stack([gu, gd]))
        print("chain", c, "sweep", s, "acc equal", np.array_equal(a, acc[c, s]), "max|G|", np.abs(ends[-1]).max())
        tot += ends[-1]
    print("  final G rel err", np.abs(gg[c] - ends[1]).max() / np.abs(ends[1]).max(), " g_sum rel err", np.abs(m["g_sum"][c] - tot).max() / np.abs(tot).max(),
          " (g_sum - G_end2) vs G_end1 rel err", np.abs(m["g_sum"][c] - gg[c] - ends[0]).max() / np.abs(ends[0]).max())
