"""Drop-in installation into the reference's model code WITHOUT editing it (INTEGRATION.md routes 2 and 4).

``install_mmcv_shim()``  registers a module ``mmcv.ops.multi_scale_deform_attn`` whose ``MultiScaleDeformableAttention`` is
    this package's keyword adapter, so that ``from mmcv.ops.multi_scale_deform_attn import MultiScaleDeformableAttention``
    at models/GroundingDINO/transformer.py:29 binds to the B200 kernels.  Call it before importing the reference's
    ``transformer`` module.  A real mmcv, if installed, is left alone unless ``force=True``.

``patch_sync_sites(groundingdino_module, transformer_module)``  replaces the host synchronisations the reference performs in
    every training step by device-side equivalents (scope row f2; same values, no ``.item()`` / ``.cpu()`` / Python loops
    over device tensors):
      * models/GroundingDINO/groundingdino.py:342-349   text_mask filled by a bs x 256 double loop of device reads
      * models/GroundingDINO/transformer.py:711-718     per-decoder-layer NaN / inf check that branches on the host
      * models/GroundingDINO/groundingdino.py:538-567,573-576,587-593,613-619   positive-map Python loops, the per-image CPU
        Hungarian matcher for the 7 output sets, the CPU ``one_hot`` builder and ``num_boxes.item()``
      * models/GroundingDINO/groundingdino.py:479-481 ``torch.masked_select`` (data-dependent size: a hidden host read) and
        groundingdino/util/box_ops.py:51-52 the two ``assert (...).all()`` of the GIoU on device tensors
    See the functions below for what each replacement keeps identical.
"""
from __future__ import annotations

import sys
import types


def install_mmcv_shim(force: bool = False, **adapter_defaults) -> types.ModuleType:
    """Make ``mmcv.ops.multi_scale_deform_attn.MultiScaleDeformableAttention`` resolve to msda_b200's adapter."""
    from .modules import MultiScaleDeformableAttention as Adapter
    name = "mmcv.ops.multi_scale_deform_attn"
    if name in sys.modules and not force and not getattr(sys.modules[name], "__msda_b200_shim__", False):
        return sys.modules[name]
    cls = Adapter
    if adapter_defaults:
        class MultiScaleDeformableAttention(Adapter):          # noqa: D401 - same name as the class it replaces
            def __init__(self, *a, **k):
                for key, val in adapter_defaults.items():
                    k.setdefault(key, val)
                super().__init__(*a, **k)
        cls = MultiScaleDeformableAttention
    shim = types.ModuleType(name)
    shim.MultiScaleDeformableAttention = cls
    shim.__msda_b200_shim__ = True
    import importlib.machinery
    shim.__spec__ = importlib.machinery.ModuleSpec(name, None)
    for pkg in ("mmcv", "mmcv.ops"):
        if pkg not in sys.modules:
            sys.modules[pkg] = types.ModuleType(pkg)
            sys.modules[pkg].__spec__ = importlib.machinery.ModuleSpec(pkg, None)
            sys.modules[pkg].__path__ = []
    sys.modules[name] = shim
    sys.modules["mmcv"].ops = sys.modules["mmcv.ops"]
    sys.modules["mmcv.ops"].multi_scale_deform_attn = shim
    return shim


# ---- host-synchronisation removal (scope row f2) ---------------------------------------------------------------------
# Two of the sites sit in the middle of long `forward` methods.  They are replaced by re-compiling THE REFERENCE'S OWN
# source of that method with the blocking statement swapped for its device-side equivalent: the text to replace must be
# present exactly once (the patch fails loudly when the reference changes), everything else of the method stays the
# reference's code, and no file of the reference is edited.
_TEXT_MASK_LOOP = """        for b in range(bs):
            for j in range(len_td):
                if text_dict['text_token_mask'][b][j] == True:
                    out['text_mask'][b][j] = True
"""
_TEXT_MASK_VECTOR = """        out['text_mask'][:, :len_td] = text_dict['text_token_mask']      # msda_b200.inject: one device copy, no host read
"""
_NAN_CHECK = """            if output.isnan().any() | output.isinf().any():
                print(f"output layer_id {layer_id} is nan")
                try:
                    num_nan = output.isnan().sum().item()
                    num_inf = output.isinf().sum().item()
                    print(f"num_nan {num_nan}, num_inf {num_inf}")
                except Exception as e:
                    print(e)
"""
_NAN_CHECK_DEVICE = """            # msda_b200.inject: the non-finite check stays on the device (read it with `decoder.nonfinite_layers` when wanted)
            _bad = (~torch.isfinite(output)).any()
            self.nonfinite_layers = _bad.to(torch.int32) if not hasattr(self, 'nonfinite_layers') or self.nonfinite_layers is None \\
                else self.nonfinite_layers + _bad.to(torch.int32)
"""


_GIOU_ASSERTS = """    assert (boxes1[:, 2:] >= boxes1[:, :2]).all()
    assert (boxes2[:, 2:] >= boxes2[:, :2]).all()
"""
_GIOU_NO_ASSERTS = """    # msda_b200.inject: the two `assert (...).all()` on device tensors were host reads; degenerate boxes give the same
    # inf / nan below as they always did
"""


def _recompile(owner, name: str, old: str, new: str, module, indent: int) -> None:
    """Replace `old` by `new` in the source of `owner.name` (a method: indent 4, a module-level function: indent 0) and bind
    the recompiled function in place of the original.  Globals are the reference module's own."""
    import inspect
    import textwrap
    fn = getattr(owner, name)
    if getattr(fn, "__msda_b200_patched__", False):
        return
    src = textwrap.dedent(inspect.getsource(fn))

    def shift(text):      # the snippets are written at their indentation in the file; the dedented source lost `indent` columns
        return "\n".join(line[indent:] if line.startswith(" " * indent) else line for line in text.splitlines()) + "\n"

    old_s, new_s = shift(old), shift(new)
    if src.count(old_s) != 1:
        raise RuntimeError(f"msda_b200.inject: the statement to replace was not found exactly once in "
                           f"{getattr(owner, '__name__', owner)}.{name}; the reference differs from the version this patch "
                           "was written for")
    code = compile(src.replace(old_s, new_s), inspect.getsourcefile(fn) or "<msda_b200.inject>", "exec")
    scope: dict = {}
    exec(code, module.__dict__, scope)
    patched = scope[name]
    patched.__msda_b200_patched__ = True
    setattr(owner, name, patched)


def _sync_free_focal_loss(self, outputs, targets, indices, num_boxes):
    """SetCriterion.token_sigmoid_binary_focal_loss (groundingdino.py:467-499) without ``torch.masked_select``, whose
    data-dependent output size forces a device->host read: the element-wise focal loss is evaluated on the whole
    (bs, Q, K) tensor and the padded tokens are multiplied by zero.  Same sum, same normaliser."""
    import torch
    import torch.nn.functional as F
    pred_logits = outputs["pred_logits"]
    new_targets = outputs["one_hot"].to(pred_logits.device).float()
    text_mask = outputs["text_mask"]
    assert new_targets.dim() == 3 and pred_logits.dim() == 3
    if text_mask is not None:
        # logits of padded tokens are -inf (ContrastiveEmbed masks them): keep them out of the arithmetic, 0 * inf is NaN
        keep = text_mask[:, None, :].expand_as(pred_logits)
        pred_logits = torch.where(keep, pred_logits, torch.zeros_like(pred_logits))
    p = torch.sigmoid(pred_logits)
    ce_loss = F.binary_cross_entropy_with_logits(pred_logits, new_targets, reduction="none")
    p_t = p * new_targets + (1 - p) * (1 - new_targets)
    loss = ce_loss * ((1 - p_t) ** self.focal_gamma)
    if self.focal_alpha >= 0:
        loss = (self.focal_alpha * new_targets + (1 - self.focal_alpha) * (1 - new_targets)) * loss
    if text_mask is not None:
        loss = loss * keep.to(loss.dtype)
    total_num_pos = sum(len(b[0]) for b in indices)               # shapes only
    return {"loss_ce": loss.sum() / max(total_num_pos, 1.0)}


def _make_sync_free_criterion_forward(gdino_module):
    import torch
    from matcher_b200 import HungarianMatcher, one_hot_targets

    label_map_cache: dict = {}

    def label_maps_for(token, cat_list, caption):
        """groundingdino.py:538-547, memoised per (caption, categories): host tokeniser work only."""
        maps = []
        for j in range(len(cat_list)):
            key = (caption[j], tuple(cat_list[j]))
            if key not in label_map_cache:
                rows = [gdino_module.create_positive_map(token[j], torch.tensor([i]), cat_list[j], caption[j])
                        for i in range(len(cat_list[j]))]
                label_map_cache[key] = torch.stack(rows, dim=0).squeeze(1)
            maps.append(label_map_cache[key])
        return maps

    def forward(self, outputs, targets, cat_list, caption, return_indices=False):
        """SetCriterion.forward (groundingdino.py:525-639) with the same losses and keys, and no host synchronisation:
        all 7 output sets are matched in one device call (matcher_b200.match_sets), the one_hot targets are built on the
        device, and num_boxes stays a device scalar (every loss only divides by it)."""
        device = next(iter(outputs.values())).device
        label_map_list = [m.to(device, non_blocking=True) for m in label_maps_for(outputs["token"], cat_list, caption)]
        if not isinstance(self.matcher, HungarianMatcher) or self.matcher.indices_device is not None:
            ref = self.matcher
            self.matcher = HungarianMatcher(cost_class=ref.cost_class, cost_bbox=ref.cost_bbox, cost_giou=ref.cost_giou,
                                            focal_alpha=ref.focal_alpha, indices_device=None)
        sets = [outputs] + list(outputs.get("aux_outputs", [])) + ([outputs["interm_outputs"]] if "interm_outputs" in outputs else [])
        pairs = self.matcher.match_sets(sets, targets, label_map_list)
        bs, Q, K = outputs["pred_logits"].shape
        for s, out in enumerate(sets):
            out["one_hot"] = one_hot_targets(pairs[s], targets, label_map_list, Q, K)
            out["text_mask"] = outputs["text_mask"]

        # (torch.full is a fill kernel; torch.as_tensor([...], device=cuda) would be a blocking host-to-device copy)
        num_boxes = torch.full((1,), float(sum(len(t["labels"]) for t in targets)), dtype=torch.float, device=device)
        if gdino_module.is_dist_avail_and_initialized():
            torch.distributed.all_reduce(num_boxes)
        num_boxes = torch.clamp(num_boxes / gdino_module.get_world_size(), min=1)[0]       # a device scalar, never read back

        losses = {}
        for loss in self.losses:
            losses.update(self.get_loss(loss, outputs, targets, pairs[0], num_boxes))
        n_aux = len(outputs.get("aux_outputs", []))
        for idx in range(n_aux):
            for loss in self.losses:
                l_dict = self.get_loss(loss, sets[1 + idx], targets, pairs[1 + idx], num_boxes)
                losses.update({k + f"_{idx}": v for k, v in l_dict.items()})
        if "interm_outputs" in outputs:
            for loss in self.losses:
                l_dict = self.get_loss(loss, sets[-1], targets, pairs[-1], num_boxes)
                losses.update({k + "_interm": v for k, v in l_dict.items()})
        if return_indices:
            return losses, [pairs[1 + i] for i in range(n_aux)] + ([pairs[-1]] if "interm_outputs" in outputs else []) + [pairs[0]]
        return losses

    forward.__msda_b200_patched__ = True
    return forward


def patch_sync_sites(groundingdino_module, transformer_module, criterion: bool = True) -> None:
    """Apply the replacements listed in the module docstring to the reference's classes (idempotent)."""
    _recompile(groundingdino_module.GroundingDINO, "forward", _TEXT_MASK_LOOP, _TEXT_MASK_VECTOR, groundingdino_module, 4)
    _recompile(transformer_module.TransformerDecoder, "forward", _NAN_CHECK, _NAN_CHECK_DEVICE, transformer_module, 4)
    if criterion:
        box_ops = groundingdino_module.box_ops            # groundingdino.util.box_ops: GIoU of the box loss (:453-456)
        _recompile(box_ops, "generalized_box_iou", _GIOU_ASSERTS, _GIOU_NO_ASSERTS, box_ops, 0)
        cls = groundingdino_module.SetCriterion
        cls.token_sigmoid_binary_focal_loss = _sync_free_focal_loss
        if not getattr(cls.forward, "__msda_b200_patched__", False):
            cls.forward = _make_sync_free_criterion_forward(groundingdino_module)
