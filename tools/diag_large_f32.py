"""Dev tool: the 65536-human crowd in fp32 on one GPU (whole crowd, or one rank's slice of a `ways`-way split), ms per sub-step."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from social_navigation_pyenvs_b200 import scenarios
from social_navigation_pyenvs_b200.large import LargeCrowd
ways, rank = int(sys.argv[1]), int(sys.argv[2])
sc = scenarios.jittered_grid_crowd(256, pitch=2.0, jitter=0.5, seed=0)
perm = scenarios.spatial_order(sc["states"][0, :, 0:2])
perm = scenarios.deal_tiles(perm, ways)
S, G = np.ascontiguousarray(sc["states"][0, perm]), np.ascontiguousarray(sc["goals"][0, perm])
n = S.shape[0] // ways
crowd = LargeCrowd("hsfm_farina", S, G, dtype=torch.float32, shard=(rank * n, n))
for _ in range(2):
    crowd.step(0.0125, 1)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    crowd.step(0.0125, 1)
b.record(); torch.cuda.synchronize()
print(f"fp32 {ways}-way shard, rank {rank}: {a.elapsed_time(b) / 20:.4f} ms per sub-step")
