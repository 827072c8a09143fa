#!/usr/bin/env python
"""Benchmark of the HS-field Metropolis sweep (BASELINE.json metric: HS spin-flip updates/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|cfg1] [--impl reference]

A *step* is one full sweep (`LatticeQMC._update_step`, lqmc.py:301-347) of every chain of the
workload: N*L proposals per chain.  `value` = proposals/s with field, G and RNG state already
resident in HBM (device Philox uniforms); `e2e` = the same sweep through the reference-facing call
with HOST buffers: field and uniforms uploaded from pinned memory, field and (gf_up, gf_dn) read
back, every step.  One process per GPU (torchrun for N > 1), chains sharded in contiguous blocks,
no communication inside the sweeps (weak scaling: chains per GPU fixed); one NCCL all-reduce of
the observables after the timed region.

`--impl reference` times the CPU restatement of the reference (oracle/, the interpreted rank-1
loop of lqmc.py:328-331 kept literal) on all host cores - the reference itself is Python under
/root/reference and does not exist on the GPU box.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (lattice kind, linear size, U, beta, L, chains per GPU, BASELINE.json config text)
    "cfg1": ("square", 2, 4.0, 2.0, 20, 1, "2x2 square Hubbard, U=4, t=1, beta=2, dtau=0.1"),
    "cfg2": ("square", 8, 4.0, 4.0, 40, 256, "8x8 square Hubbard, U=4, beta=4, dtau=0.1, 256 chains on 1 B200"),
    "cfg3": ("ring", 64, 8.0, 8.0, 80, 1024, "1D periodic chain N=64, U=8, beta=8, 1024 chains"),
    "cfg4": ("square", 16, 4.0, 8.0, 80, 296, "16x16 square Hubbard, U=4, beta=8, dtau=0.1, delayed rank-k updates"),
    "cfg5": ("square", 24, 6.0, 10.0, 100, 148, "24x24 square Hubbard (N=576), U=6, beta=10, dtau=0.1"),
}
METRIC = "HS spin-flip updates/sec"
UNIT = "proposals/s"


def build_workload(name, mu=None):
    from latticeqmc_b200.workloads import kinetic_and_constants
    kind, size, u, beta, lt, chains, text = WORKLOADS[name]
    ham, dtau, lamb, exp_k, exp_k_inv = kinetic_and_constants(kind, size, u, beta, lt, mu=mu)
    return dict(name=name, text=text, ham=ham, n=ham.shape[0], lt=lt, u=u, beta=beta, lamb=lamb, exp_k=exp_k,
                exp_k_inv=exp_k_inv, chains=chains)


def flops_per_sweep(n, lt, accept):
    """ALGORITHMIC FP64 flops of one parity-schedule sweep of one chain (BASELINE.md section 4,
    DESIGN.md): rank-1 4N^2 per accepted flip (2 spins), wrap 8N^3 per slice but the last,
    sweep-start product 2 spins x L GEMMs x 2N^3 plus the inverse 2 x 2N^3."""
    return accept * n * lt * 4.0 * n * n + (lt - 1) * 8.0 * n ** 3 + 2.0 * (lt * 2.0 * n ** 3 + 2.0 * n ** 3)


def flops_per_sweep_physics(n, lt, accept, stab):
    """ALGORITHMIC FP64 flops of one stabilised physics-mode sweep of one chain: rank-1 updates,
    one wrap per slice, the left UDV stack (L chain GEMMs + one QR / Q / V-GEMM per segment), the running right
    product (the same minus the last segment) and one combination (4 GEMMs + inverse) per segment; both spins."""
    nseg = -(-lt // stab)
    len_last = lt - (nseg - 1) * stab
    qr_block = (4.0 / 3 + 4.0 / 3 + 2.0) * n ** 3
    per_spin = (2 * lt - len_last) * 2.0 * n ** 3 + (2 * nseg - 1) * qr_block + nseg * 10.0 * n ** 3
    return accept * n * lt * 4.0 * n * n + lt * 8.0 * n ** 3 + 2.0 * per_spin


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, smax, power = [], set(), None, []
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1])); smax = float(parts[2]); power.append(float(parts[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            load = [c for c, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
            out.update(sm_mhz=float(np.median(load)), sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        return out


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------

def _cpu_worker(args):
    """One host core: a bounded sample of the reference schedule for this workload.  A full sweep of the
    interpreted reference takes 9 s (cfg2) to 22 min (cfg4) per core, so the sweep's three phases are each
    timed on a sample - sweep-start G once, as many proposals as the wall budget allows (at least 8, in
    visiting order from the true sweep-start state), one wrap - and combined with the algorithmic counts:
    T_sweep = t_start + N*L*t_proposal + (L-1)*t_wrap."""
    name, seed, budget, literal = args
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:
        limiter = None
    from oracle import sweep_oracle as so
    w = build_workload(name)
    n, lt = w["n"], w["lt"]
    h = so.initial_field(n, lt, seed)
    rs = np.random.RandomState(seed + 100000)
    t0 = time.perf_counter()
    gu, gd = so.sweep_start_g(h, w["exp_k"], w["lamb"])
    t_start = time.perf_counter() - t0
    proposals, l = 0, lt - 1
    t1 = time.perf_counter()
    while True:
        if literal:
            us = rs.rand(n)
            for i in range(n):
                proposals += _one_proposal(so, gu, gd, h, i, l, w["lamb"], us[i])
                if proposals >= 8 and time.perf_counter() - t1 > budget:
                    break
        else:
            so.slice_proposals(gu, gd, h, l, w["lamb"], rs.rand(n))
            proposals += n
        if time.perf_counter() - t1 > budget:
            break
        gu, gd = so.wrap(gu, gd, h, l, w["exp_k"], w["lamb"])     # keep the state realistic between slices
        l = l - 1 if l > 0 else lt - 1
    t_prop = (time.perf_counter() - t1) / proposals
    t2 = time.perf_counter()
    so.wrap(gu, gd, h, max(l, 1), w["exp_k"], w["lamb"])
    t_wrap = time.perf_counter() - t2
    t_sweep = t_start + n * lt * t_prop + (lt - 1) * t_wrap
    return n * lt / t_sweep, proposals, time.perf_counter() - t0


def _one_proposal(so, gu, gd, h, i, l, lamb, u):
    """One iteration of the reference's site loop with the literal element loop (lqmc.py:313-333)."""
    arg = 2 * lamb * h[i, l]
    d_up = 1 + (1 - gu[i, i]) * (np.exp(+arg) - 1)
    d_dn = 1 + (1 - gd[i, i]) * (np.exp(-arg) - 1)
    if u <= d_up * d_dn:
        c_up = -(np.exp(-arg) - 1) * gu[i, :]
        c_up[i] += (np.exp(-arg) - 1)
        c_dn = -(np.exp(+arg) - 1) * gd[i, :]
        c_dn[i] += (np.exp(+arg) - 1)
        e_up = gu[:, i] / (1 + c_up[i])
        e_dn = gd[:, i] / (1 + c_dn[i])
        so.rank1_literal(gu, e_up, c_up)
        so.rank1_literal(gd, e_dn, c_dn)
        h[i, l] *= -1
    return 1


def cpu_baseline(name, budget_s, literal=True, cores=None):
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(name, 1000 + c, budget_s, literal) for c in range(cores)])
    wall = time.perf_counter() - t0
    rate = sum(r for r, _, _ in res)
    how = ("the reference's interpreted rank-1 element loop kept literal (lqmc.py:328-331)" if literal
           else "the rank-1 update vectorised as G - outer(e, c)")
    return dict(value=rate, unit=UNIT, cores=cores, kind="port",
                sample=(f"{cores} forked single-threaded workers on the same workload ({name}), oracle/sweep_oracle.py with {how}; "
                        f"per worker: sweep-start G once, proposals in visiting order for {budget_s:.0f} s "
                        f"({int(sum(p for _, p, _ in res))} in total), one wrap; proposals/s = N*L / (t_start + N*L*t_proposal + (L-1)*t_wrap), summed over workers"),
                proposals=int(sum(p for _, p, _ in res)), wall_s=wall)


# ------------------------------------------------------------------------------------------------
# arms
# ------------------------------------------------------------------------------------------------

def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    return rank, world, local


def run_reference(args):
    rank, world, local = dist_setup(args.gpus)
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    budget = max(2.0, min(30.0, 60.0 / max(1, args.steps + args.warmup)))
    samples = []
    for _ in range(args.warmup):
        cpu_baseline(args.workload, min(budget, 2.0))
    for _ in range(args.steps):
        samples.append(cpu_baseline(args.workload, budget))
    rate = float(np.mean([s["value"] for s in samples]))
    base = samples[-1]
    n, lt = (w[1] ** 2 if w[0] == "square" else w[1]), w[4]
    line = dict(metric=METRIC, value=rate, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * budget, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic", impl="reference",
                config=dict(workload=args.workload, description=w[6], n_sites=n, n_slices=lt,
                            step="bounded sample: each step runs every host core for a fixed wall budget"),
                cpu_baseline=dict(value=rate, unit=UNIT, cores=base["cores"], kind="port", sample=base["sample"]),
                e2e=dict(value=rate, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from latticeqmc_b200 import SweepEngine

    rank, world, local = dist_setup(args.gpus)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")

    physics = args.mode == "physics"
    w = build_workload(args.workload, mu=0.0 if physics else None)
    n, lt = w["n"], w["lt"]
    chains = args.chains or w["chains"]
    from latticeqmc_b200.workloads import synthetic_fields
    fields = synthetic_fields(n, lt, chains, seed0=rank * chains)
    eng = SweepEngine(w["exp_k"], w["lamb"], lt, n_chains=chains, exp_k_inv=w["exp_k_inv"], device=local,
                      mode=args.mode, arith=args.arith, chain_offset=rank * chains,
                      stab_every=args.stab if physics else 0)
    eng.set_field(fields)
    stream = torch.cuda.Stream(dev)          # a real (non-default) stream: its handle goes to the C ABI, events see it
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def device_step():
        eng.sweep_async(1, 0, seed=args.seed, measure=True, stream=stream.cuda_stream)

    # ---- device-resident throughput ----
    for _ in range(args.warmup):
        device_step()
    barrier()
    eng.reset_measurements()
    launches0 = eng.info()["launches"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (not timed)
        starts[k].record(stream)
        device_step()
        stops[k].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall
    step_ms = [a.elapsed_time(b) for a, b in zip(starts, stops)]
    launches = eng.info()["launches"] - launches0
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    m = eng.get_measurements()
    accept = float(m["n_accepted"].sum()) / (chains * args.steps * n * lt)

    # ---- end to end through the host-buffer API ----
    pin_field = torch.from_numpy(fields.copy()).pin_memory()
    pin_uni = torch.empty((chains, 1, lt, n), dtype=torch.float64).pin_memory()
    pin_uni.copy_(torch.from_numpy(np.random.RandomState(7 + rank).rand(chains, 1, lt, n)))
    f_np, u_np = pin_field.numpy(), pin_uni.numpy()

    pin_g = torch.empty((chains, 2, n, n), dtype=torch.float64).pin_memory()
    g_np = pin_g.numpy()

    def host_step():
        eng.set_field(f_np)                          # H2D: field (pinned)
        eng.sweep(1, u_np, measure=False)            # H2D: uniforms (pinned); kernel
        eng.get_field(out=f_np)                      # D2H: field, becomes the next step's input
        return eng.get_g(out=g_np)                   # D2H: (gf_up, gf_dn) of every chain, into pinned memory

    for _ in range(max(1, args.warmup)):
        host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- the only collective of the path: observables, after the sweeps ----
    reduce_ms = None
    if world > 1:
        gsum = torch.from_numpy(m["g_sum"].sum(0)).to(dev)
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dist.all_reduce(gsum); b.record(); torch.cuda.synchronize(dev)
        reduce_ms = a.elapsed_time(b)

    if rank == 0:
        proposals_per_step = world * chains * n * lt
        value = proposals_per_step * args.steps / (total_ms * 1e-3)
        e2e_value = proposals_per_step * args.steps / e2e_s
        flops = (flops_per_sweep_physics(n, lt, accept, args.stab) if physics and args.stab
                 else flops_per_sweep(n, lt, accept)) * chains            # per step (one rank)
        kernel_ms = total_ms / args.steps
        peaks = load_peaks()
        achieved = flops / (kernel_ms * 1e-3) * 1e-12
        info = eng.info()
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=kernel_ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
            config=dict(workload=args.workload, description=w["text"], n_sites=n, n_slices=lt, chains_per_gpu=chains,
                        chains_total=world * chains, mode=args.mode, stab_every=args.stab if physics else 0,
                        arith=args.arith, rng="device philox4x32-10",
                        kernel_family=info["family"], l2="flushed between timed steps (256 MiB memset, untimed)",
                        parallelism=f"chains sharded over {world} GPU(s), no collective in the sweep"),
            accept_rate=accept, accepted_flips_per_s=value * accept,
            roofline=dict(bound="tensor", pipe="FP64 (DFMA and DMMA issue to the same pipe on sm_100a; no tcgen05 f64 kind)",
                          achieved=achieved, peak=peaks["fp64_tflops"], unit="TFLOP/s", frac=achieved / peaks["fp64_tflops"],
                          traffic=None if physics else load_traffic(args.workload), peak_source=peaks["source"],
                          kernel="sweep_reg_kernel" if info["family"] == "reg" else "sweep_l2_kernel",
                          flops_per_launch=flops, hbm=None if physics else hbm_leg(args.workload, kernel_ms, peaks)),
            e2e=dict(value=e2e_value, unit=UNIT,
                     h2d_bytes_per_step=int(world * (fields.nbytes + pin_uni.numel() * 8)),
                     d2h_bytes_per_step=int(world * (fields.nbytes + chains * 2 * n * n * 8)),
                     ms_per_step=1e3 * e2e_s / args.steps),
            gpu_launches=int(launches), clocks=clocks, wall_s_timed_region=t_wall, observable_allreduce_ms=reduce_ms,
        )
        if not args.no_cpu and world == 1:
            base = cpu_baseline(args.workload, args.cpu_budget, literal=True)
            vec = cpu_baseline(args.workload, min(args.cpu_budget, 5.0), literal=False)
            base["vectorised_port_value"] = vec["value"]
            line["cpu_baseline"] = base
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def load_peaks():
    """FP64 pipe peak measured on this pool's B200 by tools/fp64_peak.cu (MEASURED_PEAKS.json carries
    no FP64 figure); HBM from MEASURED_PEAKS.json, else the profiling guide's fallback."""
    out = dict(fp64_tflops=36.8, source="profiles/fp64_peaks_r01.json (tools/fp64_peak.cu on this pool's B200, sustained DFMA)",
               hbm_gbs=6650.0)
    try:
        p = json.load(open(os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")))
        out["fp64_tflops"] = float(p["dfma_tflops_sustained"])
    except Exception:
        out["source"] = "fallback 36.8 TFLOP/s (profiles/fp64_peaks_r01.json unreadable)"
    try:
        out["hbm_gbs"] = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    return out


def hbm_leg(workload, kernel_ms, peaks):
    """Secondary roofline: measured DRAM bytes of the launch (ncu, profiles/traffic.json) over the live event time."""
    t = load_traffic(workload)
    if not t:
        return None
    gbs = t / (kernel_ms * 1e-3) * 1e-9
    return dict(achieved=gbs, peak=peaks["hbm_gbs"], unit="GB/s", frac=gbs / peaks["hbm_gbs"])


def load_traffic(workload):
    """dram bytes per launch from the committed ncu capture of this workload, if any."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("LQMC_BENCH_WORKLOAD", "cfg4"), choices=sorted(WORKLOADS),
                    help="default cfg4 = the 16x16, beta=8 configuration north_star quotes the metric on")
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the workload's)")
    ap.add_argument("--arith", default="exact", choices=["exact", "fma"])
    ap.add_argument("--mode", default="parity", choices=["parity", "physics"],
                    help="parity = the reference recurrence (headline); physics = textbook DQMC at true half filling")
    ap.add_argument("--stab", type=int, default=0, help="physics mode: QR/UDV stabilisation every this many slices")
    ap.add_argument("--seed", type=int, default=20260101)
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
