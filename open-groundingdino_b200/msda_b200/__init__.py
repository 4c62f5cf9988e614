"""msda_b200 -- B200-native multi-scale deformable attention behind the reference's operator API.

Mirrors models/GroundingDINO/ops of PAVE-ComputerVision/Open-GroundingDino:
  functions.MSDeformAttnFunction   <- ops/functions/ms_deform_attn_func.py:21-38
  modules.MSDeformAttn             <- ops/modules/ms_deform_attn.py:30-126
  modules.MultiScaleDeformableAttention  (mmcv-style keyword adapter used by
                                          models/GroundingDINO/transformer.py:761-767,794-801)
The compute lives in libmsda_b200.so (C ABI, include/msda_b200.h); there is no CPU or eager fallback.
"""
from .functions import MSDeformAttnFunction, ms_deform_attn_forward, ms_deform_attn_backward  # noqa: F401
from .modules import MSDeformAttn, MultiScaleDeformableAttention  # noqa: F401

__all__ = ["MSDeformAttnFunction", "MSDeformAttn", "MultiScaleDeformableAttention",
           "ms_deform_attn_forward", "ms_deform_attn_backward"]
