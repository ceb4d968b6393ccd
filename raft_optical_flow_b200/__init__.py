"""B200-native (sm_100a) implementation of RAFT's correlation hot path.

Public surface (mirrors the reference's core/corr.py and alt_cuda_corr extension):
    from raft_optical_flow_b200 import CorrBlock, AlternateCorrBlock, alt_cuda_corr
"""
from . import alt_cuda_corr  # noqa: F401
from .corr import AlternateCorrBlock, CorrBlock, PackedConvC1, PackedFmaps  # noqa: F401
from .upsample import patch_raft, upsample_flow  # noqa: F401  (next row of the scope table: core/raft.py:112-142)

__all__ = ["CorrBlock", "AlternateCorrBlock", "alt_cuda_corr", "upsample_flow", "patch_raft", "PackedConvC1", "PackedFmaps"]
