"""Put this directory first on PYTHONPATH and the reference's core/raft.py (`from corr import CorrBlock,
AlternateCorrBlock`, core/raft.py:8) picks up the B200-native blocks instead of core/corr.py -- the
reference scripts append 'core' to sys.path (demo.py:2, evaluate.py:2, train.py:3), so an earlier entry wins."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raft_optical_flow_b200.corr import AlternateCorrBlock, CorrBlock  # noqa: E402,F401
