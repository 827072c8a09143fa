#!/usr/bin/env python
"""Timed comparison only (BASELINE.json north_star: cuBLAS / cuSOLVER are allowed as a comparison, not on the path): the batched FP64
GEMMs of one wrap (4 products of N x N per chain, lqmc.py:338-345) through torch.bmm (cuBLAS) against the engine's fused wrap
(`lqmc_wrap`: two DMMA GEMMs per spin with the exp(V) scaling in the epilogue), and one batched QR (cuSOLVER geqrf through
torch.linalg.qr) against the engine's pre-pivoted Householder QR launch time from the ncu launch list.
usage: python tools/cublas_compare.py [workload] [chains]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_workload
from latticeqmc_b200 import SweepEngine
from latticeqmc_b200.workloads import synthetic_fields

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
w = build_workload(name)
chains = int(sys.argv[2]) if len(sys.argv) > 2 else w["chains"]
n, lt = w["n"], w["lt"]
dev = torch.device("cuda:0")


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps): fn()
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


g = torch.randn(2 * chains, n, n, dtype=torch.float64, device=dev)
e = torch.randn(n, n, dtype=torch.float64, device=dev).expand(2 * chains, n, n).contiguous()
out = torch.empty_like(g)
ms_bmm = timed(lambda: (torch.bmm(e, g, out=out), torch.bmm(out, e, out=g)))
flops = 2 * chains * 2 * 2.0 * n ** 3
print(f"{name}: {chains} chains, N = {n}")
print(f"cuBLAS (torch.bmm, FP64) two products per spin : {ms_bmm:8.3f} ms  {flops / ms_bmm / 1e9:7.2f} TFLOP/s   (no diagonal scaling, no transposed store)")

eng = SweepEngine(w["exp_k"], w["lamb"], lt, n_chains=chains, exp_k_inv=w["exp_k_inv"])
eng.set_field(synthetic_fields(n, lt, chains))
eng.recompute(0)


def wrap():
    eng.wrap(lt - 1)


wrap(); eng.sync()
t0 = time.perf_counter()
for _ in range(5): wrap()
eng.sync()
ms_wrap = (time.perf_counter() - t0) / 5 * 1e3
print(f"engine lqmc_wrap (DMMA, scaling fused, call + sync)   : {ms_wrap:8.3f} ms  {flops / ms_wrap / 1e9:7.2f} TFLOP/s")

a = torch.randn(2 * chains, n, n, dtype=torch.float64, device=dev)
ms_qr = timed(lambda: torch.linalg.qr(a), reps=2)
print(f"cuSOLVER (torch.linalg.qr, FP64, Q formed) batched : {ms_qr:8.3f} ms for {2 * chains} matrices   (engine st_qr_kernel incl. column pivoting + explicit Q: see profiles/r01e_launches_phys_cfg4_by_kernel.txt)")
