"""Aggregate pinned host <-> device copy bandwidth of the box against the number of ranks copying at once: what bounds
the e2e leg of bench.py (host buffers, every input copied every step) at world sizes > 1.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_h2d.py [--mb 512]
Every rank copies --mb MiB host->device and, on a second stream, device->host, --reps times; rank 0 prints one JSON line
with the per-rank and aggregate GB/s (max time over ranks), for H2D alone, D2H alone and both directions at once."""
import argparse
import json
import os

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=512)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = args.mb * (1 << 20) // 8
    h_in, h_out = torch.ones(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
    d_in, d_out = torch.empty(n, dtype=torch.float64, device=dev), torch.ones(n, dtype=torch.float64, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h):
        best = 1e30
        for _ in range(args.reps + 1):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            s1.wait_event(a)
            s2.wait_event(a)
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s1)
            torch.cuda.current_stream().wait_stream(s2)
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t))
        return best
    gb = n * 8 / 1e9
    out = {"world": world, "mib_per_rank_per_direction": args.mb}
    for name, (a, b) in (("h2d", (True, False)), ("d2h", (False, True)), ("both", (True, True))):
        ms = run(a, b)
        out[name] = {"ms": ms, "GBps_per_rank_per_direction": gb / ms * 1e3, "GBps_aggregate_per_direction": world * gb / ms * 1e3}
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
