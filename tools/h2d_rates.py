"""What the host can feed: every rank copies a pinned 1 GiB buffer to its GPU (and back) at the same time as the others.
The end-to-end leg of bench.py moves 1.24 GB up and 0.32 GB down per solve and rank; this is the ceiling of that leg at N
ranks, independent of the solver.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_rates.py [--bind 0|1]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bind", type=int, default=1)
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from iterative_solver_b200 import distributed as D
    rank, world, local = D.env_rank_world()
    info = D.bind_host_to_device(local) if args.bind else {"bound": False}
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    host = torch.empty(args.mib << 20, dtype=torch.uint8).pin_memory()
    host.fill_(1)  # first touch after binding
    dev = torch.empty_like(host, device="cuda")
    out = {}
    for name, (dst, src) in {"h2d": (dev, host), "d2h": (host, dev)}.items():
        dst.copy_(src)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(args.reps):
            dst.copy_(src, non_blocking=True)
        stop.record()
        torch.cuda.synchronize()
        out[name] = args.reps * host.numel() / (start.elapsed_time(stop) * 1e-3) / 1e9
    rec = {"rank": rank, "world": world, "binding": info, "h2d_gbs": round(out["h2d"], 2), "d2h_gbs": round(out["d2h"], 2)}
    if world > 1:
        recs = [None] * world
        dist.all_gather_object(recs, rec)
        dist.destroy_process_group()
    else:
        recs = [rec]
    if rank == 0:
        print(json.dumps({"world": world, "bind": args.bind, "h2d_gbs_sum": round(sum(r["h2d_gbs"] for r in recs), 1),
                          "d2h_gbs_sum": round(sum(r["d2h_gbs"] for r in recs), 1), "ranks": recs}), flush=True)


if __name__ == "__main__":
    main()
