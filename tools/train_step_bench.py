#!/usr/bin/env python
"""Training-step throughput of the REFERENCE model (BASELINE.json configs[3]) with this repo's attention dropped in.

    python tools/train_step_bench.py [--msda ours|reference] [--sync-free 0|1] [--steps K] [--warmup W] ...
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_step_bench.py ...

The model is built by the reference's own builder (`build_groundingdino`, tools/refmodel.py) from its own config file:
SwinB backbone, 1024 x 1024 synthetic images, captions of 256 tokens, ~10 boxes per image, reference SetCriterion, bf16
autocast, activation checkpointing as configured (use_checkpoint / use_transformer_ckpt = True), AdamW, DistributedData-
Parallel with a static graph like main.py:259-260, batch 4 per GPU (weak scaling).  The attention class the reference
imports from `mmcv.ops.multi_scale_deform_attn` (transformer.py:29) is this repo's adapter (`--msda ours`, INTEGRATION.md
route 2) or the reference's own in-repo module (`--msda reference`, the grid_sample path the fork runs).
`--sync-free 1` also applies msda_b200.inject.patch_sync_sites (scope row f2).

Prints one JSON object; bench.py embeds the same measurement as its `train` record.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "open-groundingdino_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def apply_freeze(model, args):
    """main.py:266-281 verbatim in effect: parameters whose name contains a freeze keyword are frozen, except that every
    parameter with 'attn' in its name is (re-)enabled."""
    if getattr(args, "freeze_keywords", None) is None:
        return
    for name, p in model.named_parameters():
        if "attn" in name:
            p.requires_grad_(True)
        else:
            for kw in args.freeze_keywords:
                if kw in name:
                    p.requires_grad_(False)
                    break


def run(msda="ours", sync_free=True, steps=5, warmup=2, batch=4, size=1024, backbone="swin_B_384_22k", text_len=256,
        freeze_bert=True, profile_nccl=True, device=None, init_dist=True):
    import torch
    import torch.distributed as dist
    import refmodel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if device is None:
        torch.cuda.set_device(local)
        device = torch.device("cuda", local)
    if world > 1 and init_dist and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)

    torch.manual_seed(1234)                                  # identical initial weights on every rank, like a loaded checkpoint
    model, criterion, args = refmodel.build_model(backbone, device=str(device), msda=msda, use_checkpoint=True,
                                                   max_text_len=text_len)
    if not freeze_bert:
        args.freeze_keywords = None
    apply_freeze(model, args)
    transformer_mod, gdino_mod = refmodel.prepare(msda)
    if sync_free:
        from msda_b200 import inject
        inject.patch_sync_sites(gdino_mod, transformer_mod)
    n_params = sum(p.numel() for p in model.parameters())
    n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
    net = model
    if world > 1:
        # main.py:259-260: DistributedDataParallel(model, device_ids=[gpu], find_unused_parameters=...) + _set_static_graph().
        # The step is host-bound (the host needs ~345 ms to issue it), so DDP's own host work shows 1:1: with its defaults
        # (25 MB buckets, gradients copied into the buckets) a step took 373 ms on 2 GPUs against 346 ms on 1; with one
        # 512 MB bucket size and gradients as bucket views 354.5 ms (profiles/r02_train_step.md).  Results are unaffected.
        ddp_kw = {"bucket_cap_mb": int(os.environ.get("MSDA_DDP_BUCKET_MB", "512")),
                  "gradient_as_bucket_view": os.environ.get("MSDA_DDP_BUCKET_VIEW", "1") == "1"}
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[device.index], find_unused_parameters=False, **ddp_kw)
        net._set_static_graph()
    step = refmodel.TrainStep(net, criterion, args, amp_dtype=torch.bfloat16)
    images, targets, cats, caps = refmodel.make_batch(batch, size, size, device, max_text_len=text_len, seed=100 + rank)
    images = images.pin_memory()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step(images, targets, cats, caps)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_ms = 0.0
    t_wall = time.perf_counter()
    e0.record()
    for _ in range(steps):
        t0 = time.perf_counter()
        loss = step(images, targets, cats, caps)
        host_ms += (time.perf_counter() - t0) * 1e3           # time the host spends issuing a step (grows with every sync)
    e1.record()
    sync()
    wall_ms = (time.perf_counter() - t_wall) * 1e3 / steps
    dev_ms = e0.elapsed_time(e1) / steps
    t = torch.tensor([dev_ms, wall_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(t[0]), float(t[1])

    nccl_ms = msda_ms = None
    if profile_nccl:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            step(images, targets, cats, caps)
            torch.cuda.synchronize()
        nccl_us = msda_us = 0.0
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                name = ev.name.lower()
                if "nccl" in name:
                    nccl_us += ev.device_time
                elif "msda" in name or "_vec_kernel" in name or "_tile_kernel" in name or "f32_to_bf16" in name \
                        or "hm::" in name or "cost_kernel" in name or "lsap_kernel" in name:
                    msda_us += ev.device_time
        nccl_ms, msda_ms = nccl_us / 1e3, msda_us / 1e3
    rec = {"images_per_s": world * batch / (max(dev_ms, wall_ms) * 1e-3), "ms_per_step": max(dev_ms, wall_ms),
           "device_ms_per_step": dev_ms, "host_issue_ms_per_step": host_ms / steps, "n_gpus": world, "batch_per_gpu": batch,
           "steps": steps, "warmup": warmup, "loss_last": float(loss), "params_total": n_params, "params_trainable": n_train,
           "allreduce_bytes": n_train * 4 if world > 1 else 0, "nccl_ms": nccl_ms, "own_kernels_ms": msda_ms,
           "model": f"reference build_groundingdino, {backbone}, {size}x{size}, {text_len} text tokens, bf16 autocast, "
                    f"checkpointing on, AdamW, DDP static graph (bucket_cap_mb 512, gradient_as_bucket_view)", "msda": msda, "sync_free_patch": bool(sync_free),
           "peak_mem_gb": torch.cuda.max_memory_allocated(device) / 1e9}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--msda", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sync-free", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--backbone", default="swin_B_384_22k")
    ap.add_argument("--text-len", type=int, default=256)
    ap.add_argument("--train-bert", action="store_true", help="leave BERT trainable (config/cfg_coco.py) instead of frozen (cfg_odvg.py)")
    a = ap.parse_args()
    rec = run(a.msda, bool(a.sync_free), a.steps, a.warmup, a.batch, a.size, a.backbone, a.text_len, not a.train_bert)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(rec))
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
