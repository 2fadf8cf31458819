"""ugaitnet_b200 -- B200-native (sm_100a) implementation of UGaitNet's training /
descriptor-extraction step and open-world k-NN.

Importing this package loads the in-tree libugaitnet_b200.so; there is no CPU fallback.
"""
from .config import (ACT_LEAKY, ACT_LINEAR, ACT_RELU, MERGE_AVG, MERGE_MAX, MERGE_SIGNMAX, NetConfig)

__all__ = ["NetConfig", "MERGE_MAX", "MERGE_AVG", "MERGE_SIGNMAX", "ACT_LINEAR", "ACT_RELU", "ACT_LEAKY"]
