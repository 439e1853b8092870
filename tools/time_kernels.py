import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from ia2c_b200.trainer import IA2CTrainer, reference_init
for (E,N,fused) in ((4096,2,True),(1024,64,False)):
    tr=IA2CTrainer(E,n_agents=N,init=reference_init(N,5,seed=0),seed=7,fused_rollout=fused)
    for _ in range(5): tr.train_episode()
    acc={}
    n=30
    for _ in range(n):
        r=tr.train_episode_timed()
        for k,v in r.items(): acc[k]=acc.get(k,0)+v/n
    print(E,N,{k:round(v*1000,2) for k,v in acc.items()}, "us")
