"""Stand-in for the reference's compiled extension of the same name.

The reference builds a pybind module ``MultiScaleDeformableAttention`` (ops/setup.py:55) and imports it
as ``MSDA`` (ops/functions/ms_deform_attn_func.py:18, models/GroundingDINO/ms_deform_attn.py:30).  Putting
this directory on ``sys.path`` makes ``import MultiScaleDeformableAttention`` resolve here, so the
reference's own functions/modules run unmodified on the B200 kernels.  Same two entry points, same
signatures (ops/src/vision.cpp:13-16).
"""
from msda_b200.functions import ms_deform_attn_backward as _backward
from msda_b200.functions import ms_deform_attn_forward as _forward


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    return _forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step):
    return _backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, im2col_step)
