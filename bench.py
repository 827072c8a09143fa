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


class _DeviceArray:
    """A raw engine buffer (lqmc_device_ptr) as a torch tensor, through the CUDA array interface - no copy."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=3)


def device_observables(eng, torch, dev):
    """The path's final reduction input, built ON THE DEVICE from the engine's accumulators (no host staging): per-chain means
    of G (2 N^2), of n_up / n_dn / n_up n_dn per site (3 N) and their squares (for the error bars over chains), and the chain
    count - one f64 vector, summed over ranks by ONE all_reduce.  Reference analogue: the Pipe + `np.sum(gf_data, 0) / procs`
    of multiprocessing.py:265-267 (which has no error bars)."""
    c, n = eng.n_chains, eng.n_sites
    g_sum = torch.as_tensor(_DeviceArray(eng.device_ptr(2)[0], (c, 2 * n * n), "<f8"), device=dev)
    obs = torch.as_tensor(_DeviceArray(eng.device_ptr(3)[0], (c, 3 * n), "<f8"), device=dev)
    n_meas = torch.as_tensor(_DeviceArray(eng.device_ptr(4)[0], (c,), "<i8"), device=dev)
    w = n_meas.clamp(min=1).to(torch.float64)[:, None]
    g_mean = g_sum / w
    o_mean = obs / w
    return torch.cat([g_mean.sum(0), o_mean.sum(0), (o_mean * o_mean).sum(0),
                      torch.tensor([float(c)], dtype=torch.float64, device=dev)])


def measure_workload(torch, dist, dev, rank, world, local, name, mode, stab, arith, chains, steps, warmup, seed, clocks=True):
    """Device-resident throughput of one workload on this rank's GPU: `warmup` untimed sweeps, then exactly `steps` sweeps, each
    its own launch, timed with CUDA events on the launching stream; max over ranks.  Returns the engine (still alive) and a dict."""
    from latticeqmc_b200 import SweepEngine
    from latticeqmc_b200.workloads import synthetic_fields
    physics = mode == "physics"
    w = build_workload(name, mu=0.0 if physics else None)
    n, lt = w["n"], w["lt"]
    fields = synthetic_fields(n, lt, chains, seed0=rank * chains)
    eng = SweepEngine(w["exp_k"], w["lamb"], lt, n_chains=chains, exp_k_inv=w["exp_k_inv"], device=local,
                      mode=mode, arith=arith, chain_offset=rank * chains, stab_every=stab if physics else 0)
    eng.set_field(fields)
    stream = torch.cuda.current_stream(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def device_step():
        eng.sweep_async(1, 0, seed=seed, measure=True, stream=stream.cuda_stream)

    for _ in range(warmup):
        device_step()
    barrier()
    eng.reset_measurements()
    launches0 = eng.info()["launches"]
    sampler = ClockSampler(local) if (clocks and rank == 0) else None
    if sampler:
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    barrier()
    t_wall = time.perf_counter()
    for k in range(steps):
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
    counts = torch.tensor([float(eng.get_measurements()["n_accepted"].sum()), float(chains)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(counts)
    accepted_total, chains_total = float(counts[0].item()), int(round(float(counts[1].item())))
    accept = accepted_total / max(1, chains_total * steps * n * lt)
    del flush
    res = dict(w=w, n=n, lt=lt, fields=fields, chains=chains, chains_total=chains_total, total_ms=total_ms, ms_per_step=total_ms / steps,
               accept=accept, launches=int(launches), wall_s=t_wall, clocks=sampler.stop() if sampler else None,
               family=eng.info()["family"], ctas_per_chain=eng.ctas_per_chain(), physics=physics, stab=stab if physics else 0, barrier=barrier)
    res["value"] = chains_total * n * lt * steps / (total_ms * 1e-3)
    res["flops_per_step_all_ranks"] = (flops_per_sweep_physics(n, lt, accept, stab) if physics and stab
                                       else flops_per_sweep(n, lt, accept)) * chains_total
    return eng, res


def roofline_block(res, world, peaks, name, mode):
    """FP64-pipe roofline of the sweep kernel: ALGORITHMIC flops of the step / live event time / GPUs, against the measured DFMA
    peak of this pool's B200 (tools/fp64_peak.cu; MEASURED_PEAKS.json has no FP64 figure) and the 40 TFLOP/s datasheet figure."""
    per_gpu_flops = res["flops_per_step_all_ranks"] / world
    achieved = per_gpu_flops / (res["ms_per_step"] * 1e-3) * 1e-12
    parity = mode == "parity"
    traffic = load_traffic(name) if parity else None
    return dict(bound="tensor", pipe="FP64 (DFMA and DMMA issue to the same pipe on sm_100a; no tcgen05 f64 kind)",
                achieved=achieved, peak=peaks["fp64_tflops"], unit="TFLOP/s", frac=achieved / peaks["fp64_tflops"],
                frac_vs_datasheet=achieved / 40.0, datasheet_peak=40.0,
                traffic=traffic,
                traffic_source=("profiles/traffic.json: dram__bytes_read + dram__bytes_write of one launch from the committed ncu --set full "
                                "capture of this workload (static, not re-measured in this run)") if traffic else None,
                peak_source=peaks["source"],
                kernel="sweep_reg_kernel" if res["family"] == "reg" else "sweep_l2_kernel",
                flops_per_launch=per_gpu_flops if not (res["physics"] and res["stab"]) else None,
                flops_per_step_per_gpu=per_gpu_flops,
                hbm=hbm_leg(name, res["ms_per_step"], peaks) if parity else None)


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank, world, local = dist_setup(args.gpus)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    stream = torch.cuda.Stream(dev)          # a real (non-default) stream: its handle goes to the C ABI, events see it
    torch.cuda.set_stream(stream)
    peaks = load_peaks()

    physics = args.mode == "physics"
    base_chains = args.chains or WORKLOADS[args.workload][5]
    if args.scaling == "strong":
        # a FIXED job (the workload's chain count) split over the ranks, like the reference's ParallelProcessManager splits a
        # fixed number of sweeps over its processes (multiprocessing.py:260-267); remainder chains go to the low ranks
        chains = base_chains // world + (1 if rank < base_chains % world else 0)
    else:
        chains = base_chains
    eng, res = measure_workload(torch, dist, dev, rank, world, local, args.workload, args.mode, args.stab, args.arith, chains,
                                args.steps, args.warmup, args.seed)
    w, n, lt, fields, barrier = res["w"], res["n"], res["lt"], res["fields"], res["barrier"]
    accept = res["accept"]

    # ---- the only collective of the path: observables + error-bar sums, from the device accumulators, after the sweeps ----
    torch.cuda.synchronize(dev)
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    device_observables(eng, torch, dev)              # untimed first call: torch compiles / loads its elementwise kernels lazily
    torch.cuda.synchronize(dev)
    a.record(stream)
    vec = device_observables(eng, torch, dev)
    b.record(stream)
    if world > 1:
        dist.all_reduce(vec)
    c.record(stream)
    torch.cuda.synchronize(dev)
    reduce_build_ms, reduce_ms = a.elapsed_time(b), (b.elapsed_time(c) if world > 1 else None)
    n_chains_all = float(vec[-1].item())
    o_mean = (vec[2 * n * n:2 * n * n + 3 * n] / n_chains_all).cpu().numpy().reshape(3, n)
    o_sq = (vec[2 * n * n + 3 * n:2 * n * n + 6 * n] / n_chains_all).cpu().numpy().reshape(3, n)
    o_err = np.sqrt(np.maximum(o_sq - o_mean ** 2, 0.0) / max(n_chains_all - 1, 1))
    observables = dict(n_up=float(o_mean[0].mean()), n_dn=float(o_mean[1].mean()), docc=float(o_mean[2].mean()),
                       docc_stderr_over_chains=float(np.sqrt((o_err[2] ** 2).mean() / n)),
                       reduced_doubles=int(vec.numel()), chains=int(n_chains_all))

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

    for _ in range(max(1, min(args.warmup, 2))):
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
    io = torch.tensor([float(fields.nbytes + pin_uni.numel() * 8), float(fields.nbytes + chains * 2 * n * n * 8)],
                      dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(io)
    h2d_bytes, d2h_bytes = int(io[0].item()), int(io[1].item())
    eng.close()
    del pin_field, pin_uni, pin_g
    torch.cuda.empty_cache()

    # ---- the other BASELINE.json configurations, measured in the same run (short: they are context for the headline) ----
    extra = {}
    if args.configs and args.workload == "cfg4" and args.mode == "parity" and args.scaling == "weak":
        plan = [("cfg2", "parity", 0, 3, 5), ("cfg3", "physics", 10, 3, 3), ("cfg5", "physics", 8, 1, 2)]
        for name, mode, stab, wu, st in plan:
            try:
                e2, r2 = measure_workload(torch, dist, dev, rank, world, local, name, mode, stab, "exact", WORKLOADS[name][5], st, wu,
                                          args.seed, clocks=True)
                e2.close()
                torch.cuda.empty_cache()
                rf = roofline_block(r2, world, peaks, name, mode)
                extra[name] = dict(description=WORKLOADS[name][6], mode=mode, stab_every=stab, arith="exact", chains_per_gpu=r2["chains"],
                                   steps=st, warmup=wu, value=r2["value"], unit=UNIT, ms_per_step=r2["ms_per_step"],
                                   accept_rate=r2["accept"], roofline_frac=rf["frac"], roofline_achieved_tflops=rf["achieved"],
                                   gpu_launches=r2["launches"], clocks=r2["clocks"])
            except Exception as exc:          # a context line must never take the headline down with it
                extra[name] = dict(error=f"{type(exc).__name__}: {exc}")

    if rank == 0:
        chains_total = res["chains_total"]
        value = res["value"]
        e2e_value = chains_total * n * lt * args.steps / e2e_s
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=res["ms_per_step"], higher_is_better=True, scaling=args.scaling, vs_baseline=None, dtype="f64", data="synthetic",
            config=dict(workload=args.workload, description=w["text"], n_sites=n, n_slices=lt, chains_per_gpu=chains,
                        chains_total=chains_total, mode=args.mode, stab_every=args.stab if physics else 0,
                        arith=args.arith, rng="device philox4x32-10",
                        kernel_family=res["family"], ctas_per_chain=res["ctas_per_chain"],
                        l2="flushed between timed steps (256 MiB memset, untimed)",
                        parallelism=(f"{chains_total} chains sharded over {world} GPU(s) "
                                     f"({'fixed total, split' if args.scaling == 'strong' else 'fixed per GPU'}), no collective in the sweep")),
            accept_rate=accept, accepted_flips_per_s=value * accept,
            roofline=roofline_block(res, world, peaks, args.workload, args.mode),
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=d2h_bytes,
                     ms_per_step=1e3 * e2e_s / args.steps),
            gpu_launches=res["launches"], clocks=res["clocks"], wall_s_timed_region=res["wall_s"],
            observable_allreduce_ms=reduce_ms,
            observable_reduction=dict(device_build_ms=reduce_build_ms, allreduce_ms=reduce_ms, doubles=observables["reduced_doubles"],
                                      source="engine accumulators through lqmc_device_ptr (no host staging); G sums, per-site n_up / n_dn / "
                                             "n_up n_dn sums and their squares over chains, chain count",
                                      n_up=observables["n_up"], n_dn=observables["n_dn"], docc=observables["docc"],
                                      docc_stderr_over_chains=observables["docc_stderr_over_chains"]),
        )
        if extra:
            line["configs"] = extra
        if not args.no_cpu and world == 1:
            base = cpu_baseline(args.workload, args.cpu_budget, literal=True)
            vec = cpu_baseline(args.workload, min(args.cpu_budget, 5.0), literal=False)
            base["vectorised_port_value"] = vec["value"]
            base["calibration"] = ("tests/test_cpu_arm_calibration.py: seconds per proposal of this port within 20 % of the unmodified "
                                   "reference's _update_step on BASELINE configs[1] (build container)")
            line["cpu_baseline"] = base
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's chain count on EVERY GPU; strong: that chain count in total, split over the GPUs")
    ap.add_argument("--no-configs", dest="configs", action="store_false",
                    help="skip the short in-run measurements of the other BASELINE.json configurations (cfg2, cfg3, cfg5)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
