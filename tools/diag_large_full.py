"""Dev tool: ms per sub-step of the whole 65536 crowd on one GPU (fused run), culled and all-pairs."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from social_navigation_pyenvs_b200 import scenarios
from social_navigation_pyenvs_b200.large import LargeCrowd
sc = scenarios.jittered_grid_crowd(256, pitch=2.0, jitter=0.5, seed=0)
perm = scenarios.spatial_order(sc["states"][0, :, 0:2])
S, G = np.ascontiguousarray(sc["states"][0, perm]), np.ascontiguousarray(sc["goals"][0, perm])
for dt_ in (torch.float64, torch.float32):
    crowd = LargeCrowd("hsfm_farina", S, G, dtype=dt_)
    for cull in (True, False):
        crowd.culling = cull
        k = 10 if cull else 2
        crowd.step(0.0125, k); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); crowd.step(0.0125, k); b.record(); torch.cuda.synchronize()
        print(f"{dt_} culling={cull}: {a.elapsed_time(b) / k:.4f} ms per sub-step")
