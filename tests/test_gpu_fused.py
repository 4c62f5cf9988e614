"""Fused pre-op path: softmax + sampling-location arithmetic inside the kernels must give
what the op-by-op module gives -- same pixel coordinates bit for bit, same outputs/gradients within tolerance."""
import numpy as np
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu
SHAPES = [(16, 22), (8, 11), (4, 6), (2, 3)]


def _tables(shapes):
    sh = torch.as_tensor(shapes, dtype=torch.long, device="cuda")
    st = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    return sh, st


def _unfused(value, sh, st, off, logit, ref, P):
    from msda_b200 import MSDeformAttnFunction
    from msda_b200.modules import sampling_locations_from
    N, Lq, M, L, _, _ = off.shape
    aw = torch.softmax(logit.float(), -1).view(N, Lq, M, L, P)
    loc = sampling_locations_from(ref, off, sh, P)
    return MSDeformAttnFunction.apply(value, sh, st, loc, aw, 64), loc


@pytest.mark.parametrize("ref_dim", [2, 4])
@pytest.mark.parametrize("vdt,qdt", [(torch.float32, torch.float32), (torch.bfloat16, torch.float32), (torch.bfloat16, torch.bfloat16),
                                     (torch.float16, torch.float16), (torch.float16, torch.float32)])
# encoder shapes (query-tile order): default kernels / row-group backward / row-group both ways; the decoder shape has one path
@pytest.mark.parametrize("encoder,fused_variant", [(True, 0), (True, 0x2000), (True, 0x400), (False, 0)])
def test_fused_function_matches_unfused(ref_dim, vdt, qdt, encoder, fused_variant, monkeypatch):
    from msda_b200 import functions
    from msda_b200.functions import MSDeformAttnFusedFunction, fused_supported
    monkeypatch.setattr(functions, "fused_variant", fused_variant)
    torch.manual_seed(ref_dim * 7 + (vdt == torch.bfloat16) + 2 * (qdt == torch.bfloat16) + 4 * (vdt == torch.float16))
    sh, st = _tables(SHAPES)
    S = int(sh.prod(1).sum())
    N, M, D, L, P = 2, 8, 32, 4, 4
    Lq = S if encoder else 77
    value = torch.randn(N, S, M, D, device="cuda").to(vdt)
    off = (torch.randn(N, Lq, M, L, P, 2, device="cuda") * 3).to(qdt)
    logit = torch.randn(N, Lq, M, L * P, device="cuda").to(qdt)
    ref = torch.rand(N, Lq, L, ref_dim, device="cuda")
    if ref_dim == 4:
        ref[..., 2:] = ref[..., 2:] * 0.3 + 0.02
    gout = torch.randn(N, Lq, M * D, device="cuda").to(vdt)
    assert fused_supported(value, off, logit, ref, L, P)

    a = [t.detach().clone().requires_grad_(True) for t in (value, off, logit)]
    out_u, loc = _unfused(a[0], sh, st, a[1], a[2], ref, P)
    out_u.backward(gout)
    b = [t.detach().clone().requires_grad_(True) for t in (value, off, logit)]
    out_f = MSDeformAttnFusedFunction.apply(b[0], sh, st, b[1], b[2], ref)
    out_f.backward(gout)
    torch.cuda.synchronize()

    tol = 1e-5 if vdt == torch.float32 else 1e-2
    n = lambda x, y: oracle.normwise(x.detach().float().cpu().numpy(), y.detach().float().cpu().numpy())
    assert n(out_f, out_u) <= tol
    assert n(b[0].grad, a[0].grad) <= tol
    qtol = 2e-5 if qdt == torch.float32 else 2e-2         # gradients of the raw linear outputs
    assert n(b[2].grad, a[2].grad) <= qtol
    # d(loss)/d(offset) is discontinuous where a sample sits on an integer pixel coordinate; both paths use the
    # same coordinates, so they agree there too
    assert n(b[1].grad, a[1].grad) <= qtol
    assert b[1].grad.dtype == qdt and b[2].grad.dtype == qdt


def test_fused_coordinates_bit_identical_to_pytorch_expression():
    """Run the fused forward with value == pixel-index planes so the output reveals which cells were used."""
    from msda_b200.functions import MSDeformAttnFusedFunction, ms_deform_attn_forward
    from msda_b200.modules import sampling_locations_from
    torch.manual_seed(5)
    sh, st = _tables(SHAPES)
    S = int(sh.prod(1).sum())
    N, M, D, L, P = 1, 8, 32, 4, 4
    for qdt, ref_dim in ((torch.float32, 2), (torch.bfloat16, 2), (torch.float32, 4), (torch.bfloat16, 4), (torch.float16, 2), (torch.float16, 4)):
        value = torch.randn(N, S, M, D, device="cuda")
        off = (torch.randn(N, S, M, L, P, 2, device="cuda") * 2).to(qdt)
        logit = torch.zeros(N, S, M, L * P, device="cuda", dtype=qdt)
        ref = torch.rand(N, S, L, ref_dim, device="cuda")
        loc = sampling_locations_from(ref, off, sh, P).float().contiguous()
        aw = torch.full((N, S, M, L, P), 1.0 / (L * P), device="cuda")
        v = value.to(qdt) if qdt != torch.float32 else value
        want = ms_deform_attn_forward(v, sh, st, loc, aw, 64)
        got = MSDeformAttnFusedFunction.apply(v, sh, st, off, logit, ref)
        # identical coordinates and exactly uniform weights => identical arithmetic => identical bits
        assert torch.equal(got, want), (qdt, ref_dim)


@pytest.mark.parametrize("autocast", [False, True])
def test_module_fused_equals_module_unfused(autocast):
    import copy
    from msda_b200 import MSDeformAttn
    torch.manual_seed(3)
    mod = MSDeformAttn(256, 4, 8, 4).cuda()
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.05)
    ref_mod = copy.deepcopy(mod)
    ref_mod.fuse_pre_op = False
    sh, st = _tables(SHAPES)
    S = int(sh.prod(1).sum())
    q = torch.randn(2, S, 256, device="cuda")
    src = torch.randn(2, S, 256, device="cuda")
    ref_pts = torch.rand(2, S, 4, 2, device="cuda")
    outs = []
    for m in (mod, ref_mod):
        x = src.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            o = m(q, ref_pts, x, sh, st)
        o.float().square().mean().backward()
        outs.append((o.detach().float(), x.grad, m.sampling_offsets.bias.grad, m.attention_weights.weight.grad))
    tol = 3e-2 if autocast else 2e-5
    for got, want in zip(*outs):
        assert oracle.normwise(got.cpu().numpy(), want.cpu().numpy()) <= tol


# ---- fused kernels against an INDEPENDENT judge ------------------------------------------------------------------
# The tests above compare the fused path with this repo's own unfused path.  Here the judge is the grid_sample
# formulation (oracle/torch_port.py, pinned to the reference by tests/golden) evaluated on the module's own PyTorch
# expressions (ops/modules/ms_deform_attn.py:98-108) with autograd for the three gradients -- nothing of this repo's
# CUDA is on the judge's side.  16-bit offsets / logits follow PyTorch's promotion rules (`bf16 / int64` stays bf16,
# `fp32 + bf16` is fp32), exactly what the module would compute under autocast.
def _judge(value, sh, shapes, off, logit, ref, P):
    from oracle import torch_port
    N, Lq, M, L, _, _ = off.shape
    v = value.detach().float().requires_grad_(True)
    o = off.detach().clone().requires_grad_(True)
    lg = logit.detach().clone().requires_grad_(True)
    aw = torch.softmax(lg.float(), -1).view(N, Lq, M, L, P)
    if ref.shape[-1] == 2:
        wh = torch.stack([sh[..., 1], sh[..., 0]], -1)
        loc = ref[:, :, None, :, None, :] + o / wh[None, None, None, :, None, :]
    else:
        loc = ref[:, :, None, :, None, :2] + o / P * ref[:, :, None, :, None, 2:] * 0.5
    out = torch_port.msda_grid_sample(v, shapes, loc.float(), aw)
    return out, v, o, lg, loc.detach().float()


@pytest.mark.parametrize("ref_dim", [2, 4])
@pytest.mark.parametrize("vdt,qdt", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float32),
                                     (torch.float16, torch.float16)])
@pytest.mark.parametrize("encoder", [True, False])
def test_fused_vs_grid_sample_port(ref_dim, vdt, qdt, encoder):
    """PRE = 1/2/3 forward and backward, query-tile order (encoder=True: one query per pixel) and linear order."""
    from conftest import generic_coordinate_mask
    from msda_b200.functions import MSDeformAttnFusedFunction
    torch.manual_seed(100 + ref_dim + 3 * (vdt == torch.bfloat16) + 5 * (qdt == torch.bfloat16) + 7 * encoder)
    shapes = [(24, 37), (12, 19), (6, 10), (3, 5)]
    sh, st = _tables(shapes)
    S = int(sh.prod(1).sum())
    N, M, D, L, P = 2, 8, 32, 4, 4
    Lq = S if encoder else 131
    value = torch.randn(N, S, M, D, device="cuda").to(vdt)
    off = (torch.randn(N, Lq, M, L, P, 2, device="cuda") * 2.5).to(qdt)
    logit = torch.randn(N, Lq, M, L * P, device="cuda").to(qdt)
    ref = torch.rand(N, Lq, L, ref_dim, device="cuda")
    if ref_dim == 4:
        ref[..., 2:] = ref[..., 2:] * 0.3 + 0.02
    gout = torch.randn(N, Lq, M * D, device="cuda").to(vdt)

    want_out, jv, jo, jl, loc = _judge(value, sh, shapes, off, logit, ref, P)
    want_out.backward(gout.float())
    b = [t.detach().clone().requires_grad_(True) for t in (value, off, logit)]
    got = MSDeformAttnFusedFunction.apply(b[0], sh, st, b[1], b[2], ref)
    got.backward(gout)
    torch.cuda.synchronize()

    tol = 1e-5 if vdt == torch.float32 else (1e-2 if vdt == torch.bfloat16 else 2e-3)
    qtol = 2e-5 if qdt == torch.float32 else (2e-2 if qdt == torch.bfloat16 else 4e-3)   # gradients stored in qdt
    n = lambda x, y: oracle.normwise(x.detach().double().cpu().numpy(), y.detach().double().cpu().numpy())
    assert n(got, want_out) <= tol, "output"
    assert n(b[0].grad, jv.grad) <= tol, "grad_value"
    assert n(b[2].grad, jl.grad) <= qtol, "grad_attention_logits"
    # grad wrt offsets is discontinuous at integer pixel coordinates and grid_sample rounds its coordinates differently
    # (SURVEY.md 7.3): compare where the sample is farther than 1e-4 px from an integer coordinate
    m = torch.from_numpy(np.ascontiguousarray(generic_coordinate_mask(loc.cpu().numpy(), np.asarray(shapes)))).cuda()
    assert n(torch.where(m, b[1].grad.float(), 0.0), torch.where(m, jo.grad.float(), 0.0)) <= qtol, "grad_sampling_offsets"
