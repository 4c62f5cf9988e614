"""Harness around the REFERENCE's own model code (benchmark / test scaffolding -- nothing of the product imports this).

What it does
  * stage():   copies the reference's Python model sources (models/, groundingdino/util/, util/, config/) from
               /root/reference into the git-ignored baseline/_ref/src, which travels to the GPU box with the snapshot
               (the reference is not pip-installable: no setup.py at its root).  Nothing is edited.
  * prepare(): puts that copy on sys.path and registers the few third-party modules this image lacks, as stand-ins with
               the semantics the reference uses: timm.models.layers (DropPath, to_2tuple, trunc_normal_), matplotlib-
               backed visualisers (never called in training), and -- INTEGRATION.md route 2 -- the module
               `mmcv.ops.multi_scale_deform_attn`, which models/GroundingDINO/transformer.py:29 imports its attention from.
               `msda="ours"` installs this repo's adapter there (msda_b200.inject), `msda="reference"` the reference's own
               in-repo module of the same name (models/GroundingDINO/ms_deform_attn.py, the grid_sample path).
  * build_model(): the reference builder `build_groundingdino(args)` (models/GroundingDINO/groundingdino.py:726-806) on
               args taken from the reference's config file (config/cfg_odvg.py) -- SwinB for BASELINE configs[3].
               No network: the text encoder is a randomly initialised BertModel(BertConfig()) (bert-base geometry,
               110 M parameters) and the tokenizer a WordPiece tokenizer over a synthetic vocabulary; both are handed to
               the reference through its own util.get_tokenlizer hooks.  transformers 5.x removed BertModel.get_head_mask
               and the encoder's head_mask argument, which models/GroundingDINO/bertwarper.py still uses: two one-line
               compatibility shims below.
  * make_batch() / TrainStep: synthetic images, captions tokenising to max_text_len tokens, ~10 boxes per image; one
               training step as engine_multi.py:499-557 does it (autocast, criterion, weighted loss sum, backward,
               clip_grad_norm_(0.1), AdamW step).
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = "/root/reference"
STAGE = os.path.join(ROOT, "baseline", "_ref", "src")
PKG_DIR = os.path.join(ROOT, "open-groundingdino_b200")

_WANTED = ("models", "groundingdino", "util", "config")


def stage(force: bool = False) -> str:
    """Copy the reference's model sources into baseline/_ref/src (only where /root/reference is mounted)."""
    marker = os.path.join(STAGE, ".staged")
    if os.path.exists(marker) and not force:
        return STAGE
    if not os.path.isdir(REF_ROOT):
        if os.path.isdir(STAGE):
            return STAGE
        raise RuntimeError("reference sources are neither at /root/reference nor staged in baseline/_ref/src")
    os.makedirs(STAGE, exist_ok=True)
    for name in _WANTED:
        dst = os.path.join(STAGE, name)
        shutil.rmtree(dst, ignore_errors=True)
        shutil.copytree(os.path.join(REF_ROOT, name), dst,
                        ignore=shutil.ignore_patterns("*.so", "*.o", "build", "__pycache__", "csrc", "*.cu", "*.cuh", "*.cpp", "*.h"))
    open(marker, "w").write("staged from /root/reference (unmodified)\n")
    return STAGE


def available() -> bool:
    return os.path.isdir(os.path.join(STAGE, "models")) or os.path.isdir(REF_ROOT)


# ---- stand-ins for third-party modules the image lacks -------------------------------------------------------------
def _timm_layers() -> types.ModuleType:
    import torch
    from torch import nn

    class DropPath(nn.Module):
        """Stochastic depth per sample (timm.models.layers.DropPath semantics)."""

        def __init__(self, drop_prob: float = 0.0):
            super().__init__()
            self.drop_prob = float(drop_prob or 0.0)

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1.0 - self.drop_prob
            mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
            return x * mask / keep

    def to_2tuple(v):
        return tuple(v) if isinstance(v, (tuple, list)) else (v, v)

    def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

    m = types.ModuleType("timm.models.layers")
    m.DropPath, m.to_2tuple, m.trunc_normal_ = DropPath, to_2tuple, trunc_normal_
    return m


def _register(name: str, module: types.ModuleType) -> None:
    import importlib.machinery
    if getattr(module, "__spec__", None) is None:        # importlib.util.find_spec() rejects spec-less modules
        module.__spec__ = importlib.machinery.ModuleSpec(name, None)
    sys.modules[name] = module
    parent, _, child = name.rpartition(".")
    if parent:
        if parent not in sys.modules:
            try:                                          # the real parent package, when there is one
                importlib.import_module(parent)
            except Exception:
                _register(parent, types.ModuleType(parent))
        setattr(sys.modules[parent], child, module)


_prepared = None


def prepare(msda: str = "ours"):
    """Make `models.GroundingDINO.*` importable with the chosen attention implementation; returns the imported
    reference modules (transformer, groundingdino).  One implementation per process (the class is bound at import)."""
    global _prepared
    if _prepared is not None:
        if _prepared[0] != msda:
            raise RuntimeError(f"reference modules already imported with msda={_prepared[0]!r}")
        return _prepared[1]
    src = stage()
    for p in (src, PKG_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    import transformers  # noqa: F401  (probes for optional packages at import: before any stand-in is registered)
    if "timm" not in sys.modules:
        try:
            importlib.import_module("timm.models.layers")
        except Exception:
            _register("timm.models.layers", _timm_layers())
    # the config loader (never used here: args come from reference_args()) needs addict and yapf
    try:
        importlib.import_module("addict")
    except Exception:
        addict = types.ModuleType("addict")

        class Dict(dict):
            __getattr__ = dict.get
            __setattr__ = dict.__setitem__

        addict.Dict = Dict
        _register("addict", addict)
    try:
        importlib.import_module("yapf.yapflib.yapf_api")
    except Exception:
        yapf_api = types.ModuleType("yapf.yapflib.yapf_api")
        yapf_api.FormatCode = lambda text, **_k: (text, False)
        _register("yapf.yapflib.yapf_api", yapf_api)
    # visualisers pull matplotlib / pycocotools; training never calls them
    for name in ("groundingdino.util.visualizer", "util.visualizer"):
        try:
            importlib.import_module(name)
        except Exception:
            stub = types.ModuleType(name)
            stub.COCOVisualizer = type("COCOVisualizer", (), {})
            _register(name, stub)
    if msda == "ours":
        from msda_b200 import inject
        inject.install_mmcv_shim()
    elif msda == "reference":
        # the in-repo module of the same name; its ctor spells the width `embed_dim`, the call sites say `embed_dims`
        # (loaded by file path: importing it as models.GroundingDINO.ms_deform_attn would run the package __init__, which
        # imports transformer.py before the attention module it wants exists)
        spec = importlib.util.spec_from_file_location("_ref_ms_deform_attn", os.path.join(src, "models", "GroundingDINO", "ms_deform_attn.py"))
        ref_mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_mod)

        class MultiScaleDeformableAttention(ref_mod.MultiScaleDeformableAttention):
            def __init__(self, embed_dims=256, num_levels=4, num_heads=8, num_points=4, batch_first=False, **kw):
                super().__init__(embed_dim=embed_dims, num_levels=num_levels, num_heads=num_heads, num_points=num_points,
                                 batch_first=batch_first, **kw)

        shim = types.ModuleType("mmcv.ops.multi_scale_deform_attn")
        shim.MultiScaleDeformableAttention = MultiScaleDeformableAttention
        _register("mmcv.ops.multi_scale_deform_attn", shim)
    else:
        raise ValueError(msda)
    transformer = importlib.import_module("models.GroundingDINO.transformer")
    gdino = importlib.import_module("models.GroundingDINO.groundingdino")
    _prepared = (msda, (transformer, gdino))
    return _prepared[1]


# ---- text side without a network ------------------------------------------------------------------------------------
LABELS = ["dent", "scratch", "missing", "scraped", "broken", "others"]          # config/cfg_odvg.py label_list


def synthetic_tokenizer(n_fill: int = 400):
    from tokenizers import Tokenizer, models, normalizers, pre_tokenizers, processors
    from transformers import PreTrainedTokenizerFast
    words = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", ".", "?"] + LABELS + ["w%03d" % i for i in range(n_fill)]
    vocab = {w: i for i, w in enumerate(words)}
    tk = Tokenizer(models.WordPiece(vocab, unk_token="[UNK]"))
    tk.normalizer = normalizers.BertNormalizer(lowercase=True)
    tk.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
    tk.post_processor = processors.TemplateProcessing(single="[CLS] $A [SEP]", pair="[CLS] $A [SEP] $B:1 [SEP]:1",
                                                      special_tokens=[("[CLS]", 2), ("[SEP]", 3)])
    return PreTrainedTokenizerFast(tokenizer_object=tk, unk_token="[UNK]", pad_token="[PAD]", cls_token="[CLS]",
                                   sep_token="[SEP]", mask_token="[MASK]")


def caption_with_tokens(n_tokens: int) -> str:
    """A caption that tokenises to exactly n_tokens ids ([CLS] ... [SEP] included): the label phrases, each followed by
    '.', then filler words."""
    parts = []
    for w in LABELS:
        parts += [w, "."]
    k = 0
    while len(parts) + 2 < n_tokens - 1:
        parts.append("w%03d" % (k % 400))
        k += 1
    parts = parts[:n_tokens - 3] + ["."]
    return " ".join(parts)


def _random_bert():
    import torch
    from transformers import BertConfig, BertModel
    cfg = BertConfig()                                   # bert-base-uncased geometry: 12 x 768, 110 M parameters
    cfg.vocab_size = 30522
    bert = BertModel(cfg)

    # transformers 5.x: BertModel.get_head_mask is gone and BertEncoder.forward lost `head_mask`; the reference's wrapper
    # (bertwarper.py:29,141-152) still passes both.  Head masks are never used (always None).
    bert.get_head_mask = lambda head_mask, n_layers, *a, **k: [None] * n_layers
    # ... and get_extended_attention_mask's third positional argument became a dtype (the wrapper passes a device)
    ext_mask = bert.get_extended_attention_mask
    bert.get_extended_attention_mask = lambda mask, shape, device=None: ext_mask(mask, shape)
    enc_forward = bert.encoder.forward

    def forward(hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                past_key_values=None, use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None):
        out = enc_forward(hidden_states, attention_mask=attention_mask, encoder_hidden_states=encoder_hidden_states,
                          encoder_attention_mask=encoder_attention_mask, past_key_values=past_key_values, use_cache=use_cache)
        for name in ("past_key_values", "hidden_states", "attentions", "cross_attentions"):
            if not hasattr(out, name):
                setattr(out, name, None)
        return out

    bert.encoder.forward = forward
    if not hasattr(type(cfg), "use_return_dict"):
        cfg.use_return_dict = True
    return bert


def reference_args(backbone: str = "swin_B_384_22k", device: str = "cuda", use_checkpoint: bool = True, **override):
    """args namespace with the fields of the reference's config file (config/cfg_odvg.py), plus what main.py adds."""
    src = stage()
    ns: dict = {}
    with open(os.path.join(src, "config", "cfg_odvg.py")) as f:
        exec(compile(f.read(), "cfg_odvg.py", "exec"), ns)
    cfg = {k: v for k, v in ns.items() if not k.startswith("__")}
    cfg.update(backbone=backbone, device=device, use_checkpoint=use_checkpoint, use_transformer_ckpt=use_checkpoint)
    cfg.update(override)
    return types.SimpleNamespace(**cfg)


def build_model(backbone: str = "swin_B_384_22k", device: str = "cuda", msda: str = "ours", use_checkpoint: bool = True,
                **override):
    """-> (model, criterion, args) from the reference's own builder."""
    import torch
    transformer, gdino = prepare(msda)
    args = reference_args(backbone, device, use_checkpoint, **override)
    from groundingdino.util import get_tokenlizer
    get_tokenlizer.get_tokenlizer = lambda *_a, **_k: synthetic_tokenizer()
    get_tokenlizer.get_pretrained_language_model = lambda *_a, **_k: _random_bert()
    model, criterion, _post = gdino.build_groundingdino(args)
    model.to(torch.device(device))
    return model, criterion, args


def make_batch(batch: int, height: int, width: int, device, max_text_len: int = 256, boxes_per_image: int = 10, seed: int = 0):
    import torch
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, height, width, generator=g)
    caption = caption_with_tokens(max_text_len)
    targets, cat_lists, captions = [], [], []
    for b in range(batch):
        n = boxes_per_image + (b % 3) - 1
        cxcy = torch.rand(n, 2, generator=g) * 0.6 + 0.2
        wh = torch.rand(n, 2, generator=g) * 0.3 + 0.05
        targets.append({"boxes": torch.cat([cxcy, wh], -1).to(device), "labels": torch.randint(0, len(LABELS), (n,), generator=g).to(device),
                        "caption": caption})
        cat_lists.append(list(LABELS))
        captions.append(caption)
    return images, targets, cat_lists, captions


class TrainStep:
    """One optimisation step the way engine_multi.py:499-557 runs it (crop construction and logging left out)."""

    def __init__(self, model, criterion, args, amp_dtype=None, lr: float = 1e-5):
        import torch
        self.torch, self.model, self.criterion, self.args = torch, model, criterion, args
        params = [p for p in model.parameters() if p.requires_grad]
        self.optimizer = torch.optim.AdamW(params, lr=lr, weight_decay=args.weight_decay)
        self.amp_dtype = amp_dtype
        self.max_norm = args.clip_max_norm

    def __call__(self, images, targets, cat_lists, captions):
        torch = self.torch
        dev = next(self.model.parameters()).device
        images = images.to(dev, non_blocking=True)
        with torch.autocast("cuda", dtype=self.amp_dtype, enabled=self.amp_dtype is not None and dev.type == "cuda"):
            outputs = self.model(images, captions=captions)
            loss_dict = self.criterion(outputs, targets, cat_lists, captions)
            weight_dict = self.criterion.weight_dict
            losses = sum(loss_dict[k] * weight_dict[k] for k in loss_dict.keys() if k in weight_dict)
        self.optimizer.zero_grad()
        losses.backward()
        if self.max_norm > 0:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.max_norm)
        self.optimizer.step()
        return losses.detach()
