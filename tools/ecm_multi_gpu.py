"""One ECM iteration of a BASELINE.json configuration through the shard-aware driver (svgpfa_b200/ecm.py) on 1..N GPUs
(SURVEY.md 8e / 8f-3).  Launch with torchrun for N > 1:

    python tools/ecm_multi_gpu.py --config config3 --out gpurun_out/ecm_1gpu.json
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/ecm_multi_gpu.py --config config3 --out gpurun_out/ecm_2gpu.json [--compare gpurun_out/ecm_1gpu.json]

Two runs from the same initial state (the synthetic trials are seeded per trial block, so every world size sees the
same data):
  shared_steps   mstep_embedding + mstep_kernels only -- lock-step: every closure evaluation all-reduces, all ranks take
                 identical optimiser decisions; the step log must equal the 1-GPU one (niter, nfeval, bounds)
  full           estep, mstep_embedding, mstep_kernels, mstep_indpointslocs -- the per-trial steps run rank-locally with
                 one all-reduce of the final bound; must terminate and never decrease the bound
Rank 0 writes one JSON file with both step logs and wall times."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="config3")
    ap.add_argument("--trials", type=int, default=None)
    ap.add_argument("--max-iter", type=int, default=20)
    ap.add_argument("--out", default=None)
    ap.add_argument("--compare", default=None)
    ap.add_argument("--optimizer", default="torch", choices=("torch", "b200"),
                    help="torch.optim.LBFGS or the device-resident svgpfa_b200.lbfgs.LBFGS")
    ap.add_argument("--sharded-steps", default="blockwise", choices=("blockwise", "joint"),
                    help="per-trial steps: rank-local optimisations, or one joint optimisation with global reductions "
                         "(needs --optimizer b200; reproduces the 1-GPU trajectory)")
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from svgpfa_b200 import ecm, sharding, synthetic
    from svgpfa_b200.testing import model_from_case

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    cfg = dict(synthetic.CONFIGS[args.config])
    if args.trials:
        cfg["R"] = args.trials
    spikes = synthetic.spike_counts_torch(cfg, dev, seed=0).sum(1).cpu().numpy()
    blocks = sharding.trial_blocks(sharding.trial_costs(spikes, cfg["N"], cfg["K"], cfg["M"], cfg["Q"]), world)
    r0, r1 = blocks[rank]
    kw = dict(max_iter=args.max_iter, lr=1.0, tolerance_grad=1e-7, tolerance_change=1e-9, line_search_fn="strong_wolfe")

    def optim_params(estimate):
        p = {"em_max_iter": 1}
        for s in ecm.STEP_ORDER["ecm"]:
            p[f"{s}_estimate"] = s in estimate
            p[f"{s}_optim_params"] = dict(kw)
        return p

    out = {"config": args.config, "world": world, "trial_blocks": [list(b) for b in blocks], "cfg": cfg, "lbfgs": kw,
           "optimizer": args.optimizer, "sharded_steps": args.sharded_steps}
    for name, estimate in (("shared_steps", ("mstep_embedding", "mstep_kernels")), ("full", ecm.STEP_ORDER["ecm"])):
        case = synthetic.make_case_torch(cfg, dev, seed=0, r0=r0, r1=r1)
        model = model_from_case(case, device=dev, process_group=pg)
        del case
        torch.cuda.synchronize(dev)
        t0 = time.time()
        hist, elapsed, msg, log = ecm.maximize(model, optim_params(estimate), process_group=pg, out=None,
                                               optimizer=args.optimizer, sharded_steps=args.sharded_steps)
        torch.cuda.synchronize(dev)
        out[name] = {"lower_bound_hist": hist, "termination": msg, "wall_s": time.time() - t0,
                     "step_log": [list(row) for row in log],
                     "spike_method": int(model._dims.spike_method), "panels": int(model._dims.pm_B)}
        del model
        torch.cuda.empty_cache()
    if rank == 0:
        if args.compare and os.path.exists(args.compare):
            ref = json.load(open(args.compare))
            cmp = {}
            a, b = out["shared_steps"]["step_log"], ref["shared_steps"]["step_log"]
            cmp["shared_steps_same_counts"] = [r[:2] + r[3:] for r in a] == [r[:2] + r[3:] for r in b]
            cmp["shared_steps_bound_rel_diff"] = [abs(x[2] - y[2]) / abs(y[2]) for x, y in zip(a, b)]
            cmp["initial_bound_rel_diff"] = abs(out["full"]["lower_bound_hist"][0] - ref["full"]["lower_bound_hist"][0]) \
                / abs(ref["full"]["lower_bound_hist"][0])
            f = [r[2] for r in out["full"]["step_log"] if r[1] != "mstep_embedding"]
            cmp["full_monotone"] = bool(all(y >= x - 1e-9 * abs(x) for x, y in zip([out["full"]["lower_bound_hist"][0]] + f, f)))
            cmp["full_final_bound_vs_1gpu_rel_diff"] = (out["full"]["lower_bound_hist"][-1] - ref["full"]["lower_bound_hist"][-1]) \
                / abs(ref["full"]["lower_bound_hist"][-1])
            if args.sharded_steps == "joint":                 # the per-trial steps follow the 1-GPU trajectory too
                a, b = out["full"]["step_log"], ref["full"]["step_log"]
                cmp["full_same_counts"] = [r[:2] + r[3:] for r in a] == [r[:2] + r[3:] for r in b]
                cmp["full_counts"] = [[r[3:] for r in a], [r[3:] for r in b]]
                cmp["full_bound_rel_diff"] = [abs(x[2] - y[2]) / abs(y[2]) for x, y in zip(a, b)]
            out["compare_with_" + os.path.basename(args.compare)] = cmp
        text = json.dumps(out)
        print(text)
        if args.out:
            with open(args.out, "w") as fh:
                fh.write(text + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
