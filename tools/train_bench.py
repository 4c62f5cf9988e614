"""Training-step harness around the op (the caller side of the path):
six deformable encoder layers in the layout of models/GroundingDINO/transformer.py:747-808 (self-attention =
MSDeformAttn, dropout 0, LayerNorm, FFN 256 -> 2048 -> 256) on synthetic SwinT-OGC 800x1333 features, batch 4 per
GPU, bf16 autocast, AdamW, DistributedDataParallel over NCCL when launched with torchrun.  Reports images/s.

    python tools/train_bench.py [--impl ours|ours-unfused|ref-op|grid-sample] [--steps 10] [--checkpoint]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py ...

Not a replica of the reference training loop (Swin, BERT, fusion, criterion are out of scope); it exercises the op
exactly as the encoder does: query = src + pos, pixel-centre reference points, value = src, 6 layers, with the
gradient all-reduce the only collective."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "open-groundingdino_b200"), os.path.join(ROOT, "tests")): sys.path.insert(0, p)
import torch
import torch.distributed as dist
from torch import nn
from torch.utils.checkpoint import checkpoint
import msda_b200.modules as mm
from msda_b200 import MSDeformAttn

SHAPES = [(100, 167), (50, 84), (25, 42), (13, 21)]


class EncoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=2048, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.norm2 = nn.LayerNorm(d_model)

    def forward(self, src, pos, ref, shapes, starts):
        src = self.norm1(src + self.self_attn(src + pos, ref, src, shapes, starts))
        return self.norm2(src + self.linear2(torch.relu(self.linear1(src))))


class Encoder(nn.Module):
    def __init__(self, layers=6, use_checkpoint=False, save_op_output=False):
        super().__init__()
        self.layers = nn.ModuleList(EncoderLayer() for _ in range(layers))
        self.use_checkpoint = use_checkpoint
        self.save_op_output = save_op_output

    def forward(self, src, pos, ref, shapes, starts):
        for layer in self.layers:
            if self.use_checkpoint and self.save_op_output:
                from torch.utils.checkpoint import create_selective_checkpoint_contexts
                from msda_b200.torch_ops import save_msda_policy
                src = checkpoint(layer, src, pos, ref, shapes, starts, use_reentrant=False,
                                 context_fn=lambda: create_selective_checkpoint_contexts(save_msda_policy))
            elif self.use_checkpoint:
                src = checkpoint(layer, src, pos, ref, shapes, starts, use_reentrant=False)
            else:
                src = layer(src, pos, ref, shapes, starts)
        return src


def reference_points(shapes, n, device):
    pts = []
    for h, w in shapes:
        ys = (torch.arange(h, device=device, dtype=torch.float32) + 0.5) / h
        xs = (torch.arange(w, device=device, dtype=torch.float32) + 0.5) / w
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        pts.append(torch.stack([xx.reshape(-1), yy.reshape(-1)], -1))
    ref = torch.cat(pts, 0)
    return ref[None, :, None, :].expand(n, -1, len(shapes), -1).contiguous()


def install_impl(name):
    """Swap the op under the module (the module code itself is unchanged)."""
    MSDeformAttn.fuse_pre_op = (name == "ours")
    if name in ("ours", "ours-unfused"):
        return
    if name == "ref-op":
        from oracle import ref_cuda
        ref = ref_cuda.load()

        class RefFn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, value, sh, st, loc, aw, step):
                ctx.save_for_backward(value, sh, st, loc, aw)
                return ref.ms_deform_attn_forward(value, sh, st, loc, aw, step)

            @staticmethod
            def backward(ctx, g):
                value, sh, st, loc, aw = ctx.saved_tensors
                gv, gl, ga = ref.ms_deform_attn_backward(value, sh, st, loc, aw, g.contiguous(), 64)
                return gv, None, None, gl, ga, None

        class Shim:   # the reference's AMP branch: widen, run, narrow (ops/modules/ms_deform_attn.py:114-120)
            @staticmethod
            def apply(value, sh, st, loc, aw, step):
                return RefFn.apply(value.float(), sh, st, loc.float(), aw.float(), step).to(value.dtype)
        mm.MSDeformAttnFunction = Shim
    elif name == "grid-sample":
        from oracle import torch_port

        class Shim:   # what the PAVE fork runs (models/GroundingDINO/ms_deform_attn.py:339-341)
            @staticmethod
            def apply(value, sh, st, loc, aw, step):
                return torch_port.msda_grid_sample(value.float(), SHAPES, loc.float(), aw.float()).to(value.dtype)
        mm.MSDeformAttnFunction = Shim
    else:
        raise SystemExit(f"unknown impl {name}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--checkpoint", action="store_true", help="activation checkpointing per layer (config/cfg_odvg.py:42)")
    ap.add_argument("--save-op-output", action="store_true",
                    help="with --checkpoint: selective checkpointing that keeps the MSDeformAttn output (row f3)")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    install_impl(a.impl)
    torch.manual_seed(1234)
    if a.save_op_output:
        MSDeformAttn.use_dispatcher_ops = True
    model = Encoder(use_checkpoint=a.checkpoint, save_op_output=a.save_op_output).to(dev)
    if world > 1:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-4)
    sh = torch.tensor(SHAPES, dtype=torch.long, device=dev)
    st = torch.cat((sh.new_zeros(1), sh.prod(1).cumsum(0)[:-1]))
    S = int(sh.prod(1).sum())
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    src = torch.randn(a.batch, S, 256, device=dev, generator=g)
    pos = torch.randn(a.batch, S, 256, device=dev, generator=g)
    ref = reference_points(SHAPES, a.batch, dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(src, pos, ref, sh, st)
        loss = out.float().square().mean()
        loss.backward()
        opt.step()
        return loss

    for _ in range(a.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"impl": a.impl, "n_gpus": world, "batch_per_gpu": a.batch, "checkpoint": a.checkpoint, "save_op_output": a.save_op_output,
                          "ms_per_step": float(t.item()), "images_per_s": world * a.batch / (float(t.item()) * 1e-3),
                          "loss": float(loss), "params": sum(p.numel() for p in model.parameters())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
