#!/usr/bin/env python
"""Where a headline episode's kernels spend their time (needs a library built with -DIA2C_STAGE_CLOCKS: `python
tools/stage_clocks.py --build` writes var_tmp/libclk.so; run with IA2C_B200_LIB=var_tmp/libclk.so).  The first / last block of
the rollout (each stage warp), of the actor-gradient kernel and of the reduce + Adam kernels print %globaltimer stamps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if "--build" in sys.argv:
    import subprocess
    from ia2c_b200 import build as B
    B.build()
    out = os.path.join(B.ROOT, "var_tmp")
    os.makedirs(out, exist_ok=True)
    objs = []
    for src in B.SOURCES:
        obj = os.path.join(out, src.replace(".cu", ".o"))
        subprocess.run([B.NVCC, *B.FLAGS, "-DIA2C_STAGE_CLOCKS", "-diag-suppress", "177", "-c", os.path.join(B.CSRC, src), "-o", obj], check=True)
        objs.append(obj)
    subprocess.run([B.NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", os.path.join(out, "libclk.so"), *objs, "-lcudart"], check=True)
    for o in objs:
        os.remove(o)
    sys.exit(0)
import torch

from ia2c_b200.trainer import IA2CTrainer, reference_init

tr = IA2CTrainer(4096, n_agents=2, init=reference_init(2, 5, seed=0), seed=1)
for _ in range(7):   # the kernels print during episode 5
    tr.train_episode()
torch.cuda.synchronize()
