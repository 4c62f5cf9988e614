#!/usr/bin/env python
"""Where does the host-buffer (e2e) path saturate when N ranks run at once?

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py [--mb 1024]

Per rank, with pinned host buffers and cudaMemcpyAsync on two streams:
  solo   : H2D, D2H and both directions at once while every other rank idles (ranks take turns)
  all    : the same three measurements with all ranks copying at the same time
plus what the container exposes of the host: CPU count / affinity mask, NUMA nodes under /sys, a multi-threaded host
memcpy bandwidth, and the effect of pinning each rank to a disjoint block of cores before its buffers are allocated.
Rank 0 prints one JSON object.  (profiles/r02_pcie_probe.json explains bench.py's e2e curve at N = 2, 4, 8.)
"""
import argparse
import glob
import json
import os
import time


def host_memcpy_gbs(torch, mb, threads):
    torch.set_num_threads(threads)
    a = torch.empty(mb * 1024 * 1024, dtype=torch.uint8)
    b = torch.empty_like(a)
    a.fill_(1)
    b.copy_(a)
    t0 = time.perf_counter()
    for _ in range(3):
        b.copy_(a)
    return 3 * a.numel() * 2 / (time.perf_counter() - t0) / 1e9      # read + write bytes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--bind", type=int, default=0, help="1: pin each rank to its own block of cores before allocating")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    ncpu = os.cpu_count() or 1
    if args.bind:
        per = max(1, ncpu // world)
        os.sched_setaffinity(0, set(range(local * per, min(ncpu, (local + 1) * per))))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.mb * 1024 * 1024
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(kind, reps=4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if kind in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if kind in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return reps * n * (2 if kind == "both" else 1) / dt / 1e9

    for k in ("h2d", "d2h", "both"):
        run(k, 1)
    solo = {}
    for r in range(world):                      # ranks take turns
        barrier()
        if r == rank:
            solo = {k: run(k) for k in ("h2d", "d2h", "both")}
        barrier()
    allr = {}
    for k in ("h2d", "d2h", "both"):
        barrier()
        allr[k] = run(k)
        barrier()
    vals = torch.tensor([solo["h2d"], solo["d2h"], solo["both"], allr["h2d"], allr["d2h"], allr["both"]], device=dev, dtype=torch.float64)
    gathered = [torch.zeros_like(vals) for _ in range(world)]
    if world > 1:
        dist.all_gather(gathered, vals)
    else:
        gathered = [vals]
    if rank == 0:
        nodes = sorted(os.path.basename(p) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
        table = [[round(float(x), 1) for x in g.tolist()] for g in gathered]
        out = {"n_ranks": world, "copy_mb": args.mb, "bound_to_core_blocks": bool(args.bind), "cpu_count": ncpu,
               "affinity_mask_size": len(os.sched_getaffinity(0)), "numa_nodes_visible": nodes,
               "host_memcpy_gbs": {"1_thread": round(host_memcpy_gbs(torch, 512, 1), 1),
                                   "all_threads": round(host_memcpy_gbs(torch, 512, ncpu), 1)},
               "columns": ["solo_h2d", "solo_d2h", "solo_both", "all_h2d", "all_d2h", "all_both"], "per_rank_gbs": table,
               "aggregate_all_both_gbs": round(sum(r[5] for r in table), 1), "sum_of_solo_both_gbs": round(sum(r[2] for r in table), 1)}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
