"""matcher_b200 -- the reference's Hungarian matcher with the cost matrix and the assignment solved on the device.

Mirrors models/GroundingDINO/matcher.py of PAVE-ComputerVision/Open-GroundingDino:
  HungarianMatcher   <- matcher.py:25-119 (same constructor, same ``forward(outputs, targets, label_map)``)
  build_matcher      <- matcher.py:204-217
The compute lives in libmsda_b200.so (C ABI, include/hungarian_b200.h); there is no CPU or scipy fallback.
"""
from .matcher import HungarianMatcher, build_matcher, hungarian_assign, hungarian_match, one_hot_targets  # noqa: F401

__all__ = ["HungarianMatcher", "build_matcher", "hungarian_assign", "hungarian_match", "one_hot_targets"]
