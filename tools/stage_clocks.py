#!/usr/bin/env python
"""Per-stage cycle counts of the fused rollout pipeline (needs a library built with -DIA2C_STAGE_CLOCKS; pass it with
IA2C_B200_LIB=...).  Runs a few headline episodes; block 0's four stage warps print their work / total cycles."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ia2c_b200.trainer import IA2CTrainer, reference_init

tr = IA2CTrainer(4096, n_agents=2, init=reference_init(2, 5, seed=0), seed=1)
for _ in range(3):
    tr.train_episode()
torch.cuda.synchronize()
