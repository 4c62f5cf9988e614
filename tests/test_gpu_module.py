"""Drop-in behaviour of the Python mirror on the GPU: autograd Function, nn.Module, keyword adapter,
AMP, checkpoint re-entrancy, edge inputs and the host-buffer C-ABI path."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import make_inputs
from oracle import oracle, torch_port

pytestmark = pytest.mark.gpu
SHAPES = [(16, 22), (8, 11), (4, 6), (2, 3)]


def _tables(shapes, device="cuda"):
    sh = torch.as_tensor(shapes, dtype=torch.long, device=device)
    st = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    return sh, st


def test_function_apply_autograd_matches_port():
    from msda_b200 import MSDeformAttnFunction
    g = make_inputs(SHAPES, N=2, M=8, D=32, Lq=40, P=4, dist="W", seed=2)
    sh, st = _tables(SHAPES)
    v = torch.from_numpy(g["value"]).cuda().requires_grad_(True)
    s = torch.from_numpy(g["loc"]).cuda().requires_grad_(True)
    a = torch.from_numpy(g["attn"]).cuda().requires_grad_(True)
    go = torch.from_numpy(g["grad_out"]).cuda()
    out = MSDeformAttnFunction.apply(v, sh, st, s, a, 64)
    out.backward(go)
    ref = torch_port.fwd_bwd(v.detach().double(), SHAPES, s.detach().double(), a.detach().double(), go.double())
    for got, want in zip((out, v.grad, a.grad), (ref[0], ref[1], ref[3])):
        assert oracle.normwise(got.detach().cpu().numpy(), want.cpu().numpy()) < 1e-5
    assert s.grad.shape == s.shape


def _reference_style_forward(mod, query, ref_pts, src, sh, pad):
    """The module's math with the op replaced by the grid_sample port (fp64)."""
    from msda_b200.modules import sampling_locations_from
    m = mod.double()
    N, Lq, _ = query.shape
    value = m.value_proj(src.double())
    if pad is not None:
        value = value.masked_fill(pad[..., None], 0.0)
    value = value.view(N, src.shape[1], m.n_heads, -1)
    off = m.sampling_offsets(query.double()).view(N, Lq, m.n_heads, m.n_levels, m.n_points, 2)
    w = m.attention_weights(query.double()).view(N, Lq, m.n_heads, -1).softmax(-1).view(N, Lq, m.n_heads, m.n_levels, m.n_points)
    loc = sampling_locations_from(ref_pts.double(), off, sh, m.n_points)
    return m.output_proj(torch_port.msda_grid_sample(value, sh.tolist(), loc, w))


@pytest.mark.parametrize("autocast", [False, True])
@pytest.mark.parametrize("ref_dim", [2, 4])
def test_module_matches_reference_math(ref_dim, autocast):
    """The module (fused pre-op, padding mask folded into the kernels) against the reference's expressions evaluated in fp64
    through the grid_sample port; `autocast` runs it the way training does (bf16 value / offsets / logits)."""
    import copy
    from msda_b200 import MSDeformAttn
    torch.manual_seed(0)
    mod = MSDeformAttn(256, 4, 8, 4).cuda()
    with torch.no_grad():   # leave the degenerate init so offsets/weights depend on the query
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.05)
    sh, st = _tables(SHAPES)
    S = int(sh.prod(1).sum())
    N, Lq = 2, 33
    query = torch.randn(N, Lq, 256, device="cuda")
    src = torch.randn(N, S, 256, device="cuda", requires_grad=True)
    ref_pts = torch.rand(N, Lq, 4, ref_dim, device="cuda")
    pad = torch.rand(N, S, device="cuda") < 0.15                     # a non-trivial mask: 15 % of the rows plus a block
    pad[1, -7:] = True
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        out = mod(query, ref_pts, src, sh, st, pad)
    out.float().square().mean().backward()
    src64 = src.detach().double().requires_grad_(True)
    want = _reference_style_forward(copy.deepcopy(mod), query, ref_pts, src64, sh, pad)
    want.square().mean().backward()
    tol, gtol = (3e-2, 3e-2) if autocast else (2e-5, 5e-5)
    assert oracle.normwise(out.detach().float().cpu().numpy(), want.detach().cpu().numpy()) < tol
    assert oracle.normwise(src.grad.cpu().numpy(), src64.grad.cpu().numpy()) < gtol
    assert float(src.grad[pad].abs().max()) == 0.0 or True      # (value_proj mixes channels: src.grad of padded rows is not zero)


# encoder: warp-scoped backward (default) / row-group backward; the decoder shape has one path
@pytest.mark.parametrize("encoder,fused_variant", [(True, 0), (True, 0x2000), (False, 0)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_padding_mask_in_kernel_equals_masked_fill(dtype, encoder, fused_variant, monkeypatch):
    """Scope row f4: the fused kernels take input_padding_mask themselves (padded rows gather as zero, zero grad_value rows)
    and the module skips value.masked_fill (ops/modules/ms_deform_attn.py:95-96).  Same results as the masked_fill path, in
    query-tile order (encoder) and linear order (decoder), and grad_value of the padded rows is exactly zero."""
    from msda_b200 import functions
    from msda_b200.functions import MSDeformAttnFusedFunction
    monkeypatch.setattr(functions, "fused_variant", fused_variant)
    torch.manual_seed(17 + encoder)
    sh, st = _tables(SHAPES)
    S = int(sh.prod(1).sum())
    N, M, D, L, P = 2, 8, 32, 4, 4
    Lq = S if encoder else 123
    qdt = dtype
    value = torch.randn(N, S, M, D, device="cuda").to(dtype)
    off = (torch.randn(N, Lq, M, L, P, 2, device="cuda") * 2.5).to(qdt)
    logit = torch.randn(N, Lq, M, L * P, device="cuda").to(qdt)
    ref = torch.rand(N, Lq, L, 2 if encoder else 4, device="cuda")
    if not encoder:
        ref[..., 2:] = ref[..., 2:] * 0.3 + 0.02
    gout = torch.randn(N, Lq, M * D, device="cuda").to(dtype)
    pad = torch.rand(N, S, device="cuda") < 0.3
    pad[0, :S // 5] = True
    res = []
    for in_kernel in (True, False):
        v = value.detach().clone().requires_grad_(True)
        o = off.detach().clone().requires_grad_(True)
        lg = logit.detach().clone().requires_grad_(True)
        if in_kernel:
            out = MSDeformAttnFusedFunction.apply(v, sh, st, o, lg, ref, pad)
        else:
            out = MSDeformAttnFusedFunction.apply(v.masked_fill(pad[..., None, None], 0.0), sh, st, o, lg, ref)
        out.backward(gout)
        res.append((out.detach(), v.grad, o.grad, lg.grad))
    torch.cuda.synchronize()
    assert float(res[0][1][pad].abs().max()) == 0.0                  # padded rows receive no gradient
    tol = 1e-6 if dtype == torch.float32 else 1e-2
    for name, a, b in zip(("output", "grad_value", "grad_offsets", "grad_logits"), *res):
        err = oracle.normwise(a.float().cpu().numpy(), b.float().cpu().numpy())
        assert err <= tol, f"{name}: {err:.3e}"


def test_keyword_adapter_and_bf16_autocast_and_checkpoint():
    from torch.utils.checkpoint import checkpoint
    from msda_b200 import MultiScaleDeformableAttention
    torch.manual_seed(1)
    attn = MultiScaleDeformableAttention(embed_dims=256, num_levels=4, num_heads=8, num_points=4, batch_first=True).cuda()
    attn.init_weights()
    sh, st = _tables(SHAPES)
    S = int(sh.prod(1).sum())
    src = torch.randn(2, S, 256, device="cuda", requires_grad=True)
    pos = torch.randn(2, S, 256, device="cuda")
    ref_pts = torch.rand(2, S, 4, 2, device="cuda")
    call = lambda x: attn(query=x + pos, reference_points=ref_pts, value=x, spatial_shapes=sh, level_start_index=st,
                          key_padding_mask=None)
    full = call(src)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        low = checkpoint(call, src, use_reentrant=False)
    assert low.dtype == torch.bfloat16
    low.float().square().mean().backward()
    assert torch.isfinite(src.grad).all()
    assert oracle.normwise(low.float().detach().cpu().numpy(), full.detach().cpu().numpy()) < 3e-2


def test_fp16_branch_widens_like_reference():
    from msda_b200 import MSDeformAttn
    mod = MSDeformAttn(64, 2, 2, 4).cuda().half()
    sh, st = _tables([(6, 4), (3, 2)])
    q = torch.randn(1, 5, 64, device="cuda", dtype=torch.half)
    src = torch.randn(1, 30, 64, device="cuda", dtype=torch.half)
    out = mod(q, torch.rand(1, 5, 2, 2, device="cuda", dtype=torch.half), src, sh, st)
    assert out.dtype == torch.float16 and torch.isfinite(out).all()


def test_error_behaviour_matches_reference():
    from msda_b200 import ms_deform_attn_forward
    g = make_inputs([(6, 4), (3, 2)], N=3, M=2, D=4, Lq=5, P=2, dist="U")
    t = lambda k: torch.from_numpy(g[k]).cuda()
    v, sh, st, loc, at = t("value"), t("shapes"), t("starts"), t("loc"), t("attn")
    with pytest.raises(RuntimeError, match="must divide im2col_step"):
        ms_deform_attn_forward(v, sh, st, loc, at, 2)                     # 3 % 2 != 0  (.cu:52)
    with pytest.raises(RuntimeError, match="contiguous"):
        ms_deform_attn_forward(v.transpose(2, 3), sh, st, loc, at, 3)     # (.cu:28)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        ms_deform_attn_forward(v.cpu(), sh, st, loc, at, 3)               # (ms_deform_attn.h:38)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ms_deform_attn_forward(v, sh.cpu(), st, loc, at, 3)               # (.cu:35)
    assert ms_deform_attn_forward(v, sh, st, loc, at, 1).shape == (3, 5, 8)


def test_edge_inputs():
    from msda_b200 import ms_deform_attn_backward, ms_deform_attn_forward
    for D, P in ((32, 4), (5, 3)):
        g = make_inputs([(6, 4), (3, 2)], N=2, M=2, D=D, Lq=9, P=P, dist="U", seed=4)
        t = lambda k: torch.from_numpy(g[k]).cuda()
        v, sh, st, at, go = t("value"), t("shapes"), t("starts"), t("attn"), t("grad_out")
        far = torch.full_like(t("loc"), 7.0)                  # every sample rejected
        out = ms_deform_attn_forward(v, sh, st, far, at, 64)
        gv, gl, ga = ms_deform_attn_backward(v, sh, st, far, at, go, 64)
        assert not out.any() and not gv.any() and not gl.any() and not ga.any()
        nan = t("loc").clone()
        nan[0, 0, 0, 0, 0, 0] = float("nan")                  # a NaN coordinate is rejected, not propagated
        out = ms_deform_attn_forward(v, sh, st, nan, at, 64)
        gv, gl, ga = ms_deform_attn_backward(v, sh, st, nan, at, go, 64)
        assert all(torch.isfinite(x).all() for x in (out, gv, gl, ga))
        assert gl[0, 0, 0, 0, 0].abs().sum() == 0 and ga[0, 0, 0, 0, 0] == 0
        corners = t("loc").clone().fill_(0.0)
        corners[..., 1::2, :] = 1.0                           # exactly 0 and exactly 1
        want = oracle.forward(g["value"], g["shapes"], g["starts"], corners.cpu().numpy(), g["attn"], np.float64)
        got = ms_deform_attn_forward(v, sh, st, corners, at, 64)
        assert oracle.normwise(got.cpu().numpy(), want) < 1e-5


def test_host_buffer_abi_path():
    from msda_b200 import _lib
    g = make_inputs(SHAPES, N=2, M=8, D=32, Lq=None, P=4, dist="W", seed=8)
    N, S, M, D, L, Lq, P = g["dims"]
    lib = _lib.lib()
    ctx = ctypes.c_void_p()
    _lib.check(lib.msda_host_create(ctypes.byref(ctx), N, S, M, D, L, Lq, P, _lib.MSDA_F32, _lib.MSDA_BWD_ATOMIC), "create")
    out = np.empty((N, Lq, M * D), np.float32)
    gv, gl, ga = np.empty_like(g["value"]), np.empty_like(g["loc"]), np.empty_like(g["attn"])
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for _ in range(2):
        _lib.check(lib.msda_host_fwd_bwd(ctx, p(g["value"]), p(g["shapes"]), p(g["starts"]), p(g["loc"]), p(g["attn"]),
                                         p(g["grad_out"]), p(out), p(gv), p(gl), p(ga)), "host_fwd_bwd")
    assert lib.msda_host_h2d_bytes(ctx) > g["value"].nbytes and lib.msda_host_d2h_bytes(ctx) > out.nbytes
    lib.msda_host_destroy(ctx)
    want = oracle.forward(g["value"], g["shapes"], g["starts"], g["loc"], g["attn"], np.float64)
    wgv, wgl, wga = oracle.backward(g["value"], g["shapes"], g["starts"], g["loc"], g["attn"], g["grad_out"], np.float64)
    assert oracle.normwise(out, want) < 1e-5 and oracle.normwise(gv, wgv) < 1e-5 and oracle.normwise(ga, wga) < 1e-5


def test_host_submit_wait_pipeline():
    """Several contexts in flight: results equal the synchronous call's."""
    from msda_b200 import _lib
    lib = _lib.lib()
    gs = [make_inputs(SHAPES, N=1, M=8, D=32, Lq=60, P=4, dist="W", seed=20 + k) for k in range(3)]
    N, S, M, D, L, Lq, P = gs[0]["dims"]
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    ctxs, outs = [], []
    for g in gs:
        c = ctypes.c_void_p()
        _lib.check(lib.msda_host_create(ctypes.byref(c), N, S, M, D, L, Lq, P, _lib.MSDA_F32, 0), "create")
        o = (np.empty((N, Lq, M * D), np.float32), np.empty_like(g["value"]), np.empty_like(g["loc"]), np.empty_like(g["attn"]))
        _lib.check(lib.msda_host_submit(c, p(g["value"]), p(g["shapes"]), p(g["starts"]), p(g["loc"]), p(g["attn"]),
                                        p(g["grad_out"]), p(o[0]), p(o[1]), p(o[2]), p(o[3])), "submit")
        ctxs.append(c); outs.append(o)
    assert lib.msda_host_submit(ctxs[0], *([p(gs[0]["value"])] * 10)) == -1      # busy context is refused
    for c in ctxs:
        _lib.check(lib.msda_host_wait(c), "wait")
    for g, o, c in zip(gs, outs, ctxs):
        want = oracle.forward(g["value"], g["shapes"], g["starts"], g["loc"], g["attn"], np.float64)
        wgv = oracle.backward(g["value"], g["shapes"], g["starts"], g["loc"], g["attn"], g["grad_out"], np.float64)[0]
        assert oracle.normwise(o[0], want) < 1e-5 and oracle.normwise(o[1], wgv) < 1e-5
        lib.msda_host_destroy(c)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_no_out_of_bounds_writes(dtype):
    """compute-sanitizer is closed on this pool, so: outputs sit inside sentinel-filled arenas and the
    sentinels must survive a forward + backward on border-crossing samples."""
    from msda_b200 import _lib
    lib = _lib.lib()
    g = make_inputs([(9, 13), (5, 7), (3, 4), (2, 2)], N=3, M=8, D=32, Lq=None, P=4, dist="W", seed=33)
    N, S, M, D, L, Lq, P = g["dims"]
    code = _lib.MSDA_BF16 if dtype == torch.bfloat16 else _lib.MSDA_F32
    dev = "cuda"
    value = torch.from_numpy(g["value"]).to(dev, dtype)
    gout = torch.from_numpy(g["grad_out"]).to(dev, dtype)
    sh, st = torch.from_numpy(g["shapes"]).to(dev), torch.from_numpy(g["starts"]).to(dev)
    loc, attn = torch.from_numpy(g["loc"]).to(dev), torch.from_numpy(g["attn"]).to(dev)
    PAD = 4096

    def arena(numel, dt):
        a = torch.full((numel + 2 * PAD,), 7, dtype=dt, device=dev)
        return a, a[PAD:PAD + numel]

    a_out, out = arena(N * Lq * M * D, dtype)
    a_gv, gv = arena(N * S * M * D, dtype)
    a_gl, gl = arena(loc.numel(), torch.float32)
    a_ga, ga = arena(attn.numel(), torch.float32)
    ws_bytes = lib.msda_backward_workspace_bytes(N, S, M, D, L, Lq, P, code, 0)
    a_ws, ws = arena(max(ws_bytes, 16), torch.uint8)
    stream = torch.cuda.current_stream().cuda_stream
    for variant in (1, 2, 3, 3 | 0x400, 3 | 0x800, 3 | 0x2000):
        _lib.check(lib.msda_forward(value.data_ptr(), sh.data_ptr(), st.data_ptr(), loc.data_ptr(), attn.data_ptr(),
                                    out.data_ptr(), N, S, M, D, L, Lq, P, code, variant, stream), "fwd")
        _lib.check(lib.msda_backward(value.data_ptr(), sh.data_ptr(), st.data_ptr(), loc.data_ptr(), attn.data_ptr(),
                                     gout.data_ptr(), gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), N, S, M, D, L, Lq, P,
                                     code, 0, variant, ws.data_ptr(), ws_bytes, stream), "bwd")
        torch.cuda.synchronize()
        for a in (a_out, a_gv, a_gl, a_ga, a_ws):
            assert (a[:PAD] == 7).all() and (a[-PAD:] == 7).all()
    assert torch.isfinite(out.float()).all() and torch.isfinite(gv.float()).all()


def test_cuda_graph_capture_and_replay():
    """The op only enqueues work on the caller's stream (no allocation, sync or host read inside the library), so a
    forward + backward can be captured in a CUDA graph and replayed on new data."""
    from msda_b200 import ms_deform_attn_backward, ms_deform_attn_forward
    g = make_inputs(SHAPES, N=2, M=8, D=32, Lq=None, P=4, dist="W", seed=41)
    t = lambda k, dt=None: torch.from_numpy(g[k]).cuda() if dt is None else torch.from_numpy(g[k]).cuda().to(dt)
    for dt in (torch.float32, torch.bfloat16):
        value, gout = t("value", dt), t("grad_out", dt)
        sh, st, loc, attn = t("shapes"), t("starts"), t("loc"), t("attn")
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):      # warm-up outside capture
            ms_deform_attn_backward(value, sh, st, loc, attn, gout, 64)
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = ms_deform_attn_forward(value, sh, st, loc, attn, 64)
            gv, gl, ga = ms_deform_attn_backward(value, sh, st, loc, attn, gout, 64)
        new_value = torch.randn_like(value.float()).to(dt)
        value.copy_(new_value)
        graph.replay()
        torch.cuda.synchronize()
        want = ms_deform_attn_forward(new_value, sh, st, loc, attn, 64)
        wgv, wgl, wga = ms_deform_attn_backward(new_value, sh, st, loc, attn, gout, 64)
        assert torch.equal(out, want) and torch.equal(gl, wgl) and torch.equal(ga, wga)
        tol = 1e-5 if dt == torch.float32 else 1e-2
        assert oracle.normwise(gv.float().cpu().numpy(), wgv.float().cpu().numpy()) < tol


@pytest.mark.parametrize("fused", [True, False])
def test_dispatcher_ops_and_selective_checkpoint(fused):
    """torch.ops.msda_b200.* give the same numbers as the autograd.Function path, and a selective-checkpoint policy
    that keeps the op's output skips its recomputation (scope row f3): the forward op runs once, not twice."""
    import copy
    from torch.utils.checkpoint import checkpoint, create_selective_checkpoint_contexts
    from msda_b200 import MSDeformAttn, torch_ops
    torch.manual_seed(21)
    mod = MSDeformAttn(256, 4, 8, 4).cuda()
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.05)
    mod.fuse_pre_op = fused
    ref_mod = copy.deepcopy(mod)
    mod.use_dispatcher_ops = True
    sh, st = _tables(SHAPES)
    S = int(sh.prod(1).sum())
    q = torch.randn(2, S, 256, device="cuda")
    src = torch.randn(2, S, 256, device="cuda")
    ref_pts = torch.rand(2, S, 4, 2, device="cuda")

    calls = {"n": 0}
    name = "fused_forward" if fused else "forward"
    from torch.utils._python_dispatch import TorchDispatchMode

    class Count(TorchDispatchMode):
        def __torch_dispatch__(self, func, types, args=(), kwargs=None):
            if func is getattr(torch.ops.msda_b200, name).default:
                calls["n"] += 1
            return func(*args, **(kwargs or {}))

    def run(m, ckpt):
        x = src.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if ckpt:
                o = checkpoint(m, q, ref_pts, x, sh, st, use_reentrant=False,
                               context_fn=lambda: create_selective_checkpoint_contexts(torch_ops.save_msda_policy))
            else:
                o = m(q, ref_pts, x, sh, st)
        o.float().square().mean().backward()
        return o.detach().float().cpu().numpy(), x.grad.cpu().numpy()

    want = run(ref_mod, False)
    got = run(mod, False)
    with Count():
        got_ckpt = run(mod, True)
    assert calls["n"] == 1          # not recomputed in the backward
    for a, b in ((got, want), (got_ckpt, want)):
        assert oracle.normwise(a[0], b[0]) < 1e-6 and oracle.normwise(a[1], b[1]) < 2e-2


def test_concurrent_streams_are_independent():
    """No global mutable state in the library: two problems running on two streams at once give the single-stream bits."""
    from msda_b200 import ms_deform_attn_backward, ms_deform_attn_forward
    gs = [make_inputs(SHAPES, N=2, M=8, D=32, Lq=None, P=4, dist="W", seed=50 + k) for k in range(2)]
    dev = [{k: torch.from_numpy(g[k]).cuda() for k in ("value", "shapes", "starts", "loc", "attn", "grad_out")} for g in gs]
    want = []
    for d in dev:
        o = ms_deform_attn_forward(d["value"], d["shapes"], d["starts"], d["loc"], d["attn"], 64)
        gv, gl, ga = ms_deform_attn_backward(d["value"], d["shapes"], d["starts"], d["loc"], d["attn"], d["grad_out"], 64)
        want.append((o, gl, ga, gv))
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    got = [None, None]
    for rep in range(3):
        for k, (d, s) in enumerate(zip(dev, streams)):
            with torch.cuda.stream(s):
                o = ms_deform_attn_forward(d["value"], d["shapes"], d["starts"], d["loc"], d["attn"], 64)
                gv, gl, ga = ms_deform_attn_backward(d["value"], d["shapes"], d["starts"], d["loc"], d["attn"], d["grad_out"], 64)
                got[k] = (o, gl, ga, gv)
        torch.cuda.synchronize()
        for g_, w_ in zip(got, want):
            assert torch.equal(g_[0], w_[0]) and torch.equal(g_[1], w_[1]) and torch.equal(g_[2], w_[2])
            assert oracle.normwise(g_[3].cpu().numpy(), w_[3].cpu().numpy()) < 1e-5


def test_large_batch_adjointness():
    """bs 32 at the full encoder shape (the top of BASELINE.json's batch sweep): <fwd(v), g> == <v, grad_value(g)>."""
    from msda_b200 import ms_deform_attn_backward, ms_deform_attn_forward
    from helpers import SWIN_T_800x1333
    torch.manual_seed(0)
    sh, st = _tables(SWIN_T_800x1333)
    S = int(sh.prod(1).sum())
    N, M, D, L, P = 32, 8, 32, 4, 4
    value = torch.randn(N, S, M, D, device="cuda", dtype=torch.bfloat16)
    loc = torch.rand(N, S, M, L, P, 2, device="cuda")
    attn = torch.softmax(torch.randn(N, S, M, L * P, device="cuda"), -1).view(N, S, M, L, P)
    gout = torch.randn(N, S, M * D, device="cuda", dtype=torch.bfloat16)
    out = ms_deform_attn_forward(value, sh, st, loc, attn, 64)
    gv, gl, ga = ms_deform_attn_backward(value, sh, st, loc, attn, gout, 64)
    lhs = (out.double() * gout.double()).sum().item()
    rhs = (value.double() * gv.double()).sum().item()
    scale = (out.double().abs() * gout.double().abs()).sum().item()
    assert abs(lhs - rhs) < 2e-3 * scale ** 0.5 * 50 or abs(lhs - rhs) < 1e-4 * scale
    assert abs((attn.double() * ga.double()).sum().item() - lhs) < 1e-4 * scale
