"""Dev tool: health of the bench crowd after a long run (NaNs, contacts, speeds) -- explains state-dependent kernel time."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import torch
from social_navigation_pyenvs_b200 import CrowdEngine

inp = bench.build_inputs("4096x25_hsfm_ccso_walls_robot", 2000)
eng = CrowdEngine.from_reference_arrays(inp["model"], inp["states"], inp["goals"], walls=inp["walls"], safety=inp["safety"],
                                        consider_robot=True, all_params_equal=True, dtype=torch.float64)
E, N = inp["E"], inp["N"]
eng.action.copy_(torch.tensor(np.tile([[0.0], [1.0]], (1, E))))
for k in range(0, 1201, 100):
    st = eng.rows(inp["states"])
    p = st[:, :N, :2]
    bad = ~np.isfinite(st[:, :N, :8]).all(axis=(1, 2))
    d = np.linalg.norm(p[:, :, None] - p[:, None], axis=-1)
    r = st[:, :N, 8]
    rr = r[:, :, None] + r[:, None]
    iu = np.triu_indices(N, 1)
    contact = (d[:, iu[0], iu[1]] < rr[:, iu[0], iu[1]])
    rob = st[:, N, :2]
    print(f"step {k}: envs with non-finite state {int(bad.sum())}, envs with a human-human contact {int(contact.any(1).sum())}, "
          f"contact pairs {int(contact.sum())}, |p| max {np.nanmax(np.abs(p)):.2f}, robot y {rob[0,1]:.1f}, speed mean {np.nanmean(np.linalg.norm(st[:, :N, 3:5], axis=-1)):.3f}")
    for _ in range(100):
        eng.step(None, bench.DT, n_substeps=20, pre_checks=True, post_checks=False, track_touch=True)
torch.cuda.synchronize()
