#!/usr/bin/env python
"""Benchmark of the svGPFA lower-bound hot path: ELBO+gradient evaluations per second, float64.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config config5]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One *step* = one unit of work of SURVEY.md §8d on one batch of synthetic input:
``model.buildKernelsMatrices(); v = model.eval(); (-v).backward()`` with ``requires_grad=True`` on
ALL parameter groups.  The workload is BASELINE.json's config #5 (R=20000 trials, N=500, K=20, M=32,
Q=200; it fits one B200).  With N GPUs the same R trials are sharded over the ranks (strong scaling)
and the packed [ELBO | dC | dd | dtheta] buffer is all-reduced once per evaluation.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      the stage that takes the most time (round 2: quad_latent_bwd), FP64-pipe bound.  achieved = (SURVEY.md
                8d's algorithmic flops of the stage + 26 flop-equivalents per kernel evaluation / exp) / t, t = the
                stage's mean device time inside the timed region (CUDA events recorded by the library on its stream);
                peak = DFMA throughput measured in this run (MEASURED_PEAKS.json has no FP64 entry); roofline.stages
                carries the same figures for every stage, roofline.spike_term the algorithm of the spike-time term
  stages_ms     mean device time of every kernel stage inside the timed region
  cpu_baseline  the oracle port timed on this box's host cores on a bounded sample of the same workload, with the
                parity of EVERY gradient group of the timed configuration on that sample
  e2e           same metric through the host-buffer C-ABI entry, copies inside the timed region
  elbo, shared_checksum   the bound and a checksum of the all-reduced [elbo | dC | dd | dtheta] buffer of the last
                timed step: the synthetic trials are seeded per trial block, so an N-rank run evaluates exactly the
                data of the 1-rank run and these must agree across N (to rounding)
  closures      secondary figures of SURVEY.md §8d: E-step closure (m, cholVecs gradients only) and embedding-M-step
                closure (cached statistics) evaluations per second
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "elbo_grad_evals_per_sec"
UNIT = "evals/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="config5")
    ap.add_argument("--trials", type=int, default=None, help="override R (debugging only; marks the line)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    return ap.parse_args()


def workload(args):
    from svgpfa_b200 import synthetic
    cfg = dict(synthetic.CONFIGS[args.config])
    if args.trials is not None:
        cfg["R"] = args.trials
    return cfg


def config_dict(args, cfg, world):
    return {"workload": f"{args.config}: synthetic R={cfg['R']} N={cfg['N']} K={cfg['K']} M={cfg['M']} Q={cfg['Q']} "
                        f"{'mixed ExpQuad/Periodic' if cfg['mixed'] else 'ExponentialQuadratic'} kernels"
                        f"{', heavy ragged spikes' if cfg['ragged'] else ''}, reg=1e-3",
            "unit_of_work": "buildKernelsMatrices(); eval(); backward() over all parameter groups",
            "trials_total": cfg["R"], "trials_per_gpu": -(-cfg["R"] // world), "parallelism": f"trial-shard x{world}",
            "sharding": "contiguous trial blocks balanced by estimated cost (spikes x pairs + per-trial quadrature work)",
            "l2": "inputs larger than L2 (spike times + per-trial parameters re-read every step)",
            "reduced": args.trials is not None}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except Exception:
                continue
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_seconds_per_trial(case_np, threads):
    from oracle import svgpfa_oracle as orc
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    orc.elbo_and_grads(case_np)
    return time.perf_counter() - t0


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference itself is Python and cannot
    travel to the GPU box) on the host cores, bounded sample per step, extrapolated linearly in trials (the
    reference's cost is linear in R: independent trials, python loops -- SURVEY.md §6.2)."""
    from svgpfa_b200 import synthetic
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload(args)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    budget = 150.0 / max(1, args.steps + args.warmup)         # seconds of CPU work per step
    probe = synthetic.make_case(dict(cfg, R=2), seed=0)
    cpu_seconds_per_trial(probe, threads)                     # warm caches / thread pools
    per_trial = cpu_seconds_per_trial(probe, threads) / 2.0
    r_sub = int(max(2, min(cfg["R"], 32, budget / max(per_trial, 1e-6))))
    case = synthetic.make_case(dict(cfg, R=r_sub), seed=0)
    from oracle import svgpfa_oracle as orc
    for _ in range(args.warmup):
        orc.elbo_and_grads(case)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.elbo_and_grads(case)
    dt = (time.perf_counter() - t0) / args.steps
    full = dt * cfg["R"] / r_sub
    value = 1.0 / full
    sample = f"{r_sub} of {cfg['R']} trials per step, time extrapolated linearly to {cfg['R']} trials"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": full * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": config_dict(args, cfg, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "cpu": cpu_model_name(), "measured_ms_per_sample_step": dt * 1e3},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def measure_peaks(device):
    """DFMA / libdevice exp / sincospi / library exp throughput of this GPU (per second); probes library."""
    from svgpfa_b200 import _cabi
    lib = _cabi.probes()
    nsm = torch.cuda.get_device_properties(device).multi_processor_count
    blocks = nsm * 8
    out = torch.empty(blocks * 256, dtype=torch.float64, device=device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    res = {}
    for kind, name, iters in ((0, "dfma", 20000), (1, "exp", 2000), (2, "sincospi", 1000), (3, "exp_lib", 2000)):
        best = None
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _cabi.check_probe(lib.svgpfa_peak_probe(kind, blocks, iters, out.data_ptr(), stream))
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        res[name] = blocks * 256 * iters * 8 / (best * 1e-3)
    return res


def hbm_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6553.9                                         # the pool's measured copy bandwidth (B200_PROFILING.md)


def algorithmic_counts(cfg, R, S):
    """SURVEY.md §8d per-unit work formulas for R trials holding S spikes."""
    N, K, M, Q = cfg["N"], cfg["K"], cfg["M"], cfg["Q"]
    n_per = sum(1 for k in range(K) if cfg["mixed"] and k % 2 == 1)
    return {
        "F_setup": R * K * (2 * M ** 3 + 6 * M ** 2),
        "F_quad": R * K * Q * (6 * M ** 2 + 18 * M),
        "F_embed": R * Q * (12 * N * K + 3 * N),
        "F_spike": S * K * 10 * M + S * 4 * K,
        "N_exp_spike": S * K * M,
        "N_sin_spike": S * n_per * M,
        "N_exp_other": R * K * M * (M + 1) // 2 + R * K * Q * M + R * Q * N,
        "bytes_min": 8 * (2 * R * K * (2 * M + M * (M + 1) // 2) + 2 * R * Q) + 12 * S,
    }


def bind_to_gpu_numa_node(index):
    """Pins this process (and hence its pinned host buffers, first-touch) to the CPUs of the NUMA node its GPU hangs off.
    With every rank on node 0 the eight ranks' host<->device copies of the e2e leg share one socket's memory
    controllers and root complexes (round 1: e2e efficiency 0.78 at 8 GPUs).  Best effort: silently skipped when the
    topology cannot be read."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def run_b200(args):
    import torch.distributed as dist
    from svgpfa_b200 import _cabi, synthetic
    from svgpfa_b200.testing import model_from_case, set_requires_grad

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
        pg = dist.group.WORLD
    cfg = workload(args)
    R = cfg["R"]
    # contiguous trial blocks balanced by estimated cost (spikes x pairs + per-trial quadrature work); every rank
    # derives the same cuts from the spike counts of ALL trials (cheap: counts only), then generates its own shard
    from svgpfa_b200 import sharding
    per_trial_spikes = synthetic.spike_counts_torch(cfg, device, seed=0).sum(1).cpu().numpy()
    costs = sharding.trial_costs(per_trial_spikes, cfg["N"], cfg["K"], cfg["M"], cfg["Q"])
    blocks = sharding.trial_blocks(costs, world)
    r0, r1 = blocks[rank]
    case = synthetic.make_case_torch(cfg, device, seed=0, r0=r0, r1=r1)
    S_local = int(case["spike_times"].numel())
    model = model_from_case(case, device=device, process_group=pg)
    set_requires_grad(model)
    leaves = model._leaves()
    lib = _cabi.lib()

    def step():
        for p in leaves:
            p.grad = None
        model.buildKernelsMatrices()
        v = model.eval()
        (-v).backward()
        return v

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(max(3, args.warmup)):
        v = step()
    model.checkErrors()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    # ---- timed region: K steps, device time, stage events recorded by the library
    n_ev = len(_cabi.STAGES) + 1
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(n_ev)] for _ in range(args.steps)]
    for row in evs:                       # events are created lazily: record once so that the handles exist
        for e in row:
            e.record()
    ev_arr = [(ctypes.c_void_p * n_ev)(*[e.cuda_event for e in row]) for row in evs]
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    start.record()
    for i in range(args.steps):          # no host synchronisation inside: the error state of every evaluation comes
        lib.svgpfa_set_stage_events(ev_arr[i])          # back through the model's asynchronous header copy
        v = step()
    lib.svgpfa_set_stage_events(None)
    stop.record()
    barrier()
    wall1 = time.time()
    model.checkErrors()
    elbo = float(v.item())
    last_shared = model._last_shared.detach()
    checksum = {"sum": float(last_shared[8:].sum().item()), "l2": float(last_shared[8:].norm().item()),
                "dC_l2": float(last_shared[8:8 + cfg["N"] * cfg["K"]].norm().item()),
                "dtheta_l2": float(last_shared[8 + cfg["N"] * cfg["K"] + cfg["N"]:].norm().item())}
    ms_total = start.elapsed_time(stop)
    stage_ms = np.zeros(len(_cabi.STAGES))
    for row in evs:
        for j in range(len(_cabi.STAGES)):
            stage_ms[j] += row[j].elapsed_time(row[j + 1])
    stage_ms /= args.steps
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    st = torch.tensor(stage_ms, dtype=torch.float64, device=device)
    tot_S = torch.tensor([float(S_local)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(st, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_S, op=dist.ReduceOp.SUM)
    ms_step = float(t.item()) / args.steps
    value = 1e3 / ms_step
    clocks = sampler.summary(wall0, wall1) if rank == 0 else None
    if rank == 0:
        sampler.stop()

    # ---- secondary figures: the closures svEM's E-step and embedding M-step evaluate
    def time_closure(fn, n):
        for _ in range(2):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        barrier()
        tt = torch.tensor([a.elapsed_time(b) / n], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def closure_of(params, objective):
        def fn():
            for p in params:
                p.grad = None
            cur = -objective()
            cur.backward()
        return fn

    n_sec = max(2, min(args.steps, 5))
    set_requires_grad(model, posterior=True, embedding=False, kernels=False, indlocs=False)
    mode0 = model.shard_mode
    ms_estep = time_closure(closure_of(model.getSVPosteriorOnIndPointsParams(), model.eval), n_sec)     # rank-local
    set_requires_grad(model, posterior=False, embedding=True, kernels=False, indlocs=False)
    stats = model.computeSVPosteriorOnLatentsStats()
    ms_emb = time_closure(closure_of(model.getSVEmbeddingParams(),
                                     lambda: model.evalELLSumAcrossTrialsAndNeurons(svPosteriorOnLatentsStats=stats)), n_sec)
    del stats
    model._cached_keepalive = None
    model.shard_mode = mode0
    set_requires_grad(model)
    model.checkErrors()
    closures = {"estep_closure": {"ms": ms_estep, "evals_per_sec": 1e3 / ms_estep,
                                  "what": "eval + backward w.r.t. m, cholVecs only (svEM.py:218-223); the spike term is "
                                          "served from its cached statistic (Z, theta, C unchanged); rank-local under sharding"},
                "mstep_embedding_closure": {"ms": ms_emb, "evals_per_sec": 1e3 / ms_emb,
                                            "what": "ELL from cached latent statistics + backward w.r.t. C, d "
                                                    "(svEM.py:225-232): quadrature embedding kernel + HBM-bound spike gather"}}

    # ---- e2e: host buffers through the C ABI, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        io = model.makeHostIO(pin=True)
        for _ in range(2):
            model.evalAndGradHost(io, copy_static=True)
        barrier()
        t0 = time.perf_counter()
        e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_start.record()
        n_e2e = max(2, min(args.steps, 5))
        for _ in range(n_e2e):
            elbo_h, h2d, d2h = model.evalAndGradHost(io, copy_static=True)      # all-reduces on the device when sharded
        e_stop.record()
        barrier()
        ms_e2e = e_start.elapsed_time(e_stop) / n_e2e
        wall_e2e = (time.perf_counter() - t0) / n_e2e * 1e3
        # parameters-only variant (spikes resident, as in an optimiser loop)
        e_start.record()
        for _ in range(n_e2e):
            _, h2d_p, d2h_p = model.evalAndGradHost(io, copy_static=False)
        e_stop.record()
        barrier()
        ms_e2e_p = e_start.elapsed_time(e_stop) / n_e2e
        # the same call without copy/compute overlap (one block, everything in order on one stream), for comparison
        e_start.record()
        for _ in range(2):
            model.evalAndGradHost(io, copy_static=True, n_blocks=1)
        e_stop.record()
        barrier()
        ms_e2e_serial = e_start.elapsed_time(e_stop) / 2
        te = torch.tensor([ms_e2e, ms_e2e_p, float(h2d), float(d2h), float(h2d_p), ms_e2e_serial], dtype=torch.float64,
                          device=device)
        if world > 1:
            mx = te.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = te.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            te = torch.stack([mx[0], mx[1], sm[2], sm[3], sm[4], mx[5]])
        e2e = {"value": 1e3 / float(te[0]), "unit": UNIT, "h2d_bytes_per_step": int(te[2]),
               "d2h_bytes_per_step": int(te[3]), "ms_per_step": float(te[0]), "wall_ms_per_step_rank0": wall_e2e,
               "params_only": {"value": 1e3 / float(te[1]), "h2d_bytes_per_step": int(te[4])},
               "unpipelined_ms_per_step": float(te[5]),
               "numa_node_rank0": numa_node,
               "api": "svgpfa_elbo_grad_host (pinned host buffers; spikes, quadrature and parameters copied every step; "
                      "copies and kernels pipelined over blocks of trials on three streams)"}
        del io

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline (FP64 pipe; see DESIGN.md "Roofline model"): every stage against the DFMA peak measured in this run
    # with SURVEY.md 8d's algorithmic flop counts (+ 26 flop-equivalents = 13 FP64 instruction slots per kernel
    # evaluation / exp); `roofline` itself describes the stage that takes the most time
    peaks = measure_peaks(device)
    Rl = r1 - r0
    cnt = algorithmic_counts(cfg, Rl, S_local)
    peak_tf = 2.0 * peaks["dfma"] / 1e12
    stage = {n: float(v) * 1e-3 for n, v in zip(_cabi.STAGES, st)}
    K_, M_, Q_, N_ = cfg["K"], cfg["M"], cfg["Q"], cfg["N"]
    exp_fl = 26.0
    panel = model._dims.spike_method == _cabi.SPIKE_PANEL
    NB = model._dims.pm_B * _cabi.PM_P if panel else 0
    c_sin = peaks["dfma"] / peaks["sincospi"]
    direct_flops = 2.0 * (13.0 * cnt["N_exp_spike"] + (c_sin + 3.0) * cnt["N_sin_spike"])
    flops = {
        "kzz_chol+indpoints_fwd+indpoints_bwd": cnt["F_setup"] + exp_fl * Rl * K_ * M_ * (M_ + 1),
        "quad_latent_fwd": cnt["F_quad"] / 3.0 + exp_fl * Rl * K_ * Q_ * M_,
        "quad_embed": cnt["F_embed"] + exp_fl * Rl * Q_ * N_,
        "quad_latent_bwd": cnt["F_quad"] * 2.0 / 3.0 + exp_fl * 2.0 * Rl * K_ * Q_ * M_,
        # panel path: kernel values at the NB nodes of every (trial, latent, inducing point), twice (latent means at the
        # nodes, then the adjoints), and the two skinny GEMMs over the panel moments (2 flops per multiply-add)
        "spike_fwd_bwd": (exp_fl * 2.0 * Rl * K_ * M_ * NB + 4.0 * Rl * N_ * NB * K_) if panel else direct_flops,
    }
    times = {"kzz_chol+indpoints_fwd+indpoints_bwd": stage["kzz_chol"] + stage["indpoints_fwd"] + stage["indpoints_bwd"],
             "quad_latent_fwd": stage["quad_latent_fwd"], "quad_embed": stage["quad_embed"],
             "quad_latent_bwd": stage["quad_latent_bwd"], "spike_fwd_bwd": stage["spike_fwd_bwd"]}
    total_t = float(st.sum()) * 1e-3
    stages = {n: {"ms": times[n] * 1e3, "algorithmic_flops": flops[n],
                  "frac": (flops[n] / times[n] / 1e12 / peak_tf) if times[n] > 0 else None,
                  "share_of_step": times[n] / total_t} for n in flops}
    dom = max(times, key=lambda n: times[n])
    kernel_names = {"quad_latent_bwd": "quad_latent_mma_kernel<MT,true>", "quad_latent_fwd": "quad_latent_mma_kernel<MT,false>",
                    "quad_embed": "quad_embed_mma_kernel", "spike_fwd_bwd": "panel_* kernels" if panel else "spike_tile_kernel",
                    "kzz_chol+indpoints_fwd+indpoints_bwd": "kzz_chol_warp / indpoints_fwd_warp / indpoints_bwd_mma"}
    roofline = {"bound": "fp64", "kernel": f"{dom}: {kernel_names[dom]}", "achieved": flops[dom] / times[dom] / 1e12,
                "peak": peak_tf, "unit": "TFLOP/s", "frac": stages[dom]["frac"], "traffic": None,
                "peak_source": "DFMA throughput measured in this run (probes library); MEASURED_PEAKS.json has no FP64 "
                               "entry; nominal 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2 TFLOP/s; mma.m8n8k4.f64 runs on the "
                               "same pipe at the same rate (tools/probe_dmma.py)",
                "model": "SURVEY.md 8d flop counts (quadrature: 1/3 forward, 2/3 adjoint) + 26 flop-equivalents per kernel "
                         "evaluation / exp, over the stage's CUDA-event time inside the timed region",
                "share_of_step": times[dom] / total_t, "launch_ms": times[dom] * 1e3, "stages": stages,
                "whole_step_frac": sum(flops.values()) / (ms_step * 1e-3) / 1e12 / peak_tf}
    # the spike-time term: which algorithm ran, and what the direct evaluation of SURVEY.md 8d would have cost
    roofline["spike_term"] = {
        "method": "panel" if panel else "direct", "panels": int(model._dims.pm_B) if panel else 0, "nodes_per_trial": NB,
        "direct_algorithmic_flops": direct_flops,
        "direct_equivalent_frac": direct_flops / times["spike_fwd_bwd"] / 1e12 / peak_tf if times["spike_fwd_bwd"] > 0 else None,
        "note": "direct_equivalent_frac > 1 means the stage finishes faster than the FP64 pipe could evaluate every (spike, "
                "latent, inducing point) kernel value: the panel path evaluates kernels at NB nodes per trial instead of "
                "S_r spikes (include/svgpfa_b200.h); the round-1 direct kernel ran this stage at 0.69",
        "hbm_bytes_panel_moments": 2.0 * 8.0 * Rl * N_ * NB if panel else 0.0,
        "hbm_frac": (2.0 * 8.0 * Rl * N_ * NB / times["spike_fwd_bwd"] / 1e9 / hbm_peak_gbs()) if panel and times["spike_fwd_bwd"] > 0 else None}
    try:        # DRAM traffic of the dominant kernel from the committed ncu --set full capture (bytes per trial x trials)
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        per_trial = tr.get(args.config, {}).get(dom)
        if per_trial:
            roofline["traffic"] = per_trial["dram_bytes_per_trial"] * Rl
            roofline["traffic_note"] = per_trial.get("note", "")
    except Exception:
        pass

    cpu = None
    if not args.no_cpu and world == 1:
        threads = os.cpu_count() or 1
        probe = synthetic.case_to_numpy(case, 0, 2)
        cpu_seconds_per_trial(probe, threads)
        per_trial = cpu_seconds_per_trial(probe, threads) / 2.0
        r_sub = int(max(2, min(r1 - r0, 32, args.cpu_seconds / max(per_trial, 1e-6))))
        sample = synthetic.case_to_numpy(case, 0, r_sub)
        from oracle import svgpfa_oracle as orc
        torch.set_num_threads(threads)
        dt = None
        for _ in range(2):                                 # first call pays first-touch / thread-pool start-up
            t0 = time.perf_counter()
            ref = orc.elbo_and_grads(sample)
            t_one = time.perf_counter() - t0
            dt = t_one if dt is None else min(dt, t_one)
        full = dt * R / r_sub
        cpu = {"value": 1.0 / full, "unit": UNIT, "cores": threads, "kind": "port", "cpu": cpu_model_name(),
               "sample": f"first {r_sub} of {R} trials, best of 2 evaluations, time extrapolated linearly to {R} trials",
               "measured_s_on_sample": dt}
        # parity of the timed configuration on that sample (same inputs, GPU path vs oracle): the bound and EVERY
        # gradient group, worst relative error (||delta||_2 / ||ref||_2) per group
        from svgpfa_b200.testing import grads_as_dict
        sub = model_from_case(synthetic.case_to_numpy(case, 0, r_sub), device=device)
        set_requires_grad(sub)
        vs = sub.eval()
        vs.backward()
        cpu["parity_elbo_rel_err"] = abs(vs.item() - ref["elbo"]) / abs(ref["elbo"])
        got = grads_as_dict(sub)
        worst = {}
        for key, g in got.items():
            grp = key.rsplit("_", 1)[0] if key[-1].isdigit() else key
            e = float(np.linalg.norm(g.reshape(-1) - ref[key].reshape(-1)) / np.linalg.norm(ref[key].reshape(-1)))
            worst[grp] = max(worst.get(grp, 0.0), e)
        cpu["parity_grad_rel_err"] = worst
        cpu["parity_ok"] = bool(cpu["parity_elbo_rel_err"] <= 1e-10 and max(worst.values()) <= 1e-8)
        del sub

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "b200",
            "config": config_dict(args, cfg, world), "elbo": elbo, "shared_checksum": checksum,
            "spikes_total": int(tot_S.item()), "trial_blocks": [list(b) for b in blocks],
            # library kernels per evaluation inside the timed region (profiles/r02_launches_config5_1gpu.csv): Cholesky
            # (+ inducing-point forward, one launch for M <= 32), quadrature forward, embedding, quadrature adjoint, the
            # spike-time term (4 panel kernels or the direct kernel), inducing-point adjoint, two finalize kernels
            "clocks": clocks, "gpu_launches": ((1 if cfg["M"] <= 32 else 2) + 3 + (4 if panel else 1) + 1 + 2) * args.steps,
            "closures": closures,
            "stages_ms": {n: float(x) for n, x in zip(_cabi.STAGES, st.tolist())},
            "roofline": roofline, "peaks_measured": {k: float(v) for k, v in peaks.items()}}
    if e2e is not None:
        line["e2e"] = e2e
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
