"""CPU checks of msda_b200.inject (no GPU): the mmcv shim resolves to the adapter, and the source rewrites of
patch_sync_sites leave the reference model's forward unchanged in value.  The reference sources come from the staged copy
(baseline/_ref/src, written by tools/refmodel.py where /root/reference is mounted)."""
import importlib
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_mmcv_shim_resolves_to_the_adapter():
    from msda_b200 import inject
    from msda_b200.modules import MultiScaleDeformableAttention
    saved = {k: sys.modules.get(k) for k in ("mmcv", "mmcv.ops", "mmcv.ops.multi_scale_deform_attn")}
    try:
        for k in saved:
            sys.modules.pop(k, None)
        inject.install_mmcv_shim()
        mod = importlib.import_module("mmcv.ops.multi_scale_deform_attn")
        assert mod.MultiScaleDeformableAttention is MultiScaleDeformableAttention
        from mmcv.ops.multi_scale_deform_attn import MultiScaleDeformableAttention as Imported   # the reference's import line
        m = Imported(embed_dims=256, num_levels=4, num_heads=8, num_points=4, batch_first=True)   # transformer.py:761-767
        assert m.batch_first and m.embed_dims == 256 and hasattr(m, "init_weights")
        assert set(dict(m.named_parameters())) == {f"{n}.{p}" for n in ("sampling_offsets", "attention_weights", "value_proj", "output_proj")
                                                   for p in ("weight", "bias")}
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.timeout(600)
def test_source_rewrites_keep_the_reference_forward():
    """text_mask loop -> slice assignment, per-layer NaN check -> device counter: same outputs on the same weights and
    inputs (CPU, the reference's own attention module)."""
    import refmodel
    if not refmodel.available():
        pytest.skip("no reference sources here (neither /root/reference nor baseline/_ref/src)")
    transformer, gdino = refmodel.prepare("reference")
    from msda_b200 import inject
    torch.manual_seed(0)
    model, criterion, args = refmodel.build_model("swin_T_224_1k", device="cpu", msda="reference", use_checkpoint=False,
                                                   max_text_len=48)
    model.eval()
    images, targets, cats, caps = refmodel.make_batch(2, 224, 256, "cpu", max_text_len=48, seed=3)   # >= 900 image tokens for the top-900 proposal selection
    with torch.no_grad():
        before = model(images, captions=caps)
        inject.patch_sync_sites(gdino, transformer, criterion=False)
        assert getattr(gdino.GroundingDINO.forward, "__msda_b200_patched__", False)
        assert getattr(transformer.TransformerDecoder.forward, "__msda_b200_patched__", False)
        after = model(images, captions=caps)
    for key in ("pred_logits", "pred_boxes", "text_mask"):
        assert torch.equal(before[key], after[key]), key
    assert int(model.transformer.decoder.nonfinite_layers) == 0          # the device-side replacement of the NaN check
    # and the patch refuses a reference whose text differs
    with pytest.raises(RuntimeError):
        inject._recompile(gdino.SetCriterion, "loss_boxes", "    this statement does not exist\n", "    pass\n", gdino, 4)
