"""`import alt_cuda_corr` (reference core/corr.py:5-9) resolves here when dropin/ is on PYTHONPATH."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raft_optical_flow_b200.alt_cuda_corr import backward, forward  # noqa: E402,F401
