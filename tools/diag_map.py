import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle
from oracle import OracleConfig
from helpers import rel_err
from test_gpu_sizes_large import _random_crowd
from social_navigation_pyenvs_b200 import CrowdEngine, scenarios
N, E = int(sys.argv[1]), int(sys.argv[2])
S, G = _random_crowd(E, N, seed=N + E, spread=2.0)
rob = np.zeros((E, 13)); rob[:, 0:2] = S[:, 0, 0:2] + 1.0; rob[:, 8] = 0.3; rob[:, 9] = 80; rob[:, 10:12] = 5.0
S1 = np.concatenate([S, rob[:, None]], 1)
walls = scenarios.pack_walls([[[-2.0, -1.0], [-1.2, -1.0], [-1.2, 6.0], [-2.0, 6.0]]])
act = np.tile([0.2, -0.1], (E, 1))
cfg = OracleConfig(oracle.type_code("hsfm_new_guo"), True, True, False)
params = np.tile(oracle.default_params("hsfm_new_guo"), (E, N, 1))
for k in (1, 2, 3, 5, 10):
    ref, _, _ = oracle.update_humans(cfg, S1, G, walls, params, np.zeros((E, N + 1)), np.zeros((E, N, 2)), 0.0125, k, robot_vel=act, n_threads=4)
    for mapping in (1, 2):
        eng = CrowdEngine.from_reference_arrays("hsfm_new_guo", S1, G, walls=walls, consider_robot=True, all_params_equal=True)
        eng.mapping = mapping
        eng.step(act, 0.0125, n_substeps=k, pre_checks=True, post_checks=True, track_touch=True)
        got = eng.rows(S1)
        e = rel_err(got[:, :N, :8], ref[:, :N, :8])
        bad = np.argwhere(e.max(axis=(1, 2)) > 1e-9).ravel()
        worst = int(np.argmax(e.max(axis=(1, 2))))
        p_, r_ = ref[worst, :N, 0:2], ref[worst, :N, 8]
        gap = np.linalg.norm(p_[:, None] - p_[None], axis=-1) - r_[:, None] - r_[None]; gap[np.arange(N), np.arange(N)] = 9
        print("   worst env", worst, "min gap", gap.min(), "robot gap", (np.linalg.norm(p_ - ref[worst, N, 0:2], axis=1) - r_ - 0.3).min())
        print("substeps", k, "mapping", mapping, "max err", e.max(), "bad envs", bad[:10], len(bad))
