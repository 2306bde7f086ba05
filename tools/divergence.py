"""Multi-step divergence of the CUDA path from the CPU oracle (= the reference's serial algorithm), reported rather than hidden.

Config 1 of BASELINE.json (single env, circular crossing, 5 humans, HSFM, dt = 0.0125, 4000 steps = the 50 s time limit) plus the
25-human static-obstacle crowd with walls; fp64 and fp32 engines against the fp64 oracle, checkpoint every 100 sub-steps.
Writes profiles/<round>_divergence.csv.  Run on the GPU box:  python tools/divergence.py r01
"""
import csv
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import OracleConfig  # noqa: E402
from social_navigation_pyenvs_b200 import CrowdEngine, scenarios  # noqa: E402

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
DT, EVERY, TOTAL = 0.0125, 100, 4000


def run(tag, model, states, goals, walls, robot_visible, rows):
    E, N = goals.shape[0], goals.shape[1]
    cfg = OracleConfig(oracle.type_code(model), robot_visible, True, False)
    params = np.tile(oracle.default_params(model), (E, N, 1))
    safety = np.zeros(states.shape[:2])
    act = np.tile([0.0, 0.35], (E, 1))
    S, G, D = states.copy(), goals.copy(), np.zeros((E, N, 2))
    eng = {d: CrowdEngine.from_reference_arrays(model, states, goals, walls=walls, consider_robot=robot_visible, all_params_equal=True, dtype=d)
           for d in (torch.float64, torch.float32)}
    first_contact = None
    for step in range(EVERY, TOTAL + 1, EVERY):
        S, G, D = oracle.update_humans(cfg, S, G, walls, params, safety, D, DT, EVERY, robot_vel=act if robot_visible else None, n_threads=8)
        p, r = S[:, :N, 0:2], S[:, :N, 8]
        gap = np.linalg.norm(p[:, :, None] - p[:, None], axis=-1) - r[:, :, None] - r[:, None]
        gap[:, np.arange(N), np.arange(N)] = np.inf
        if first_contact is None and gap.min() < 0:
            first_contact = step
        row = [tag, step, float(gap.min())]
        for d, e in eng.items():
            if robot_visible:
                e.step(act, DT, n_substeps=EVERY, pre_checks=False)
            else:
                e.update_humans(0.0, DT, n_substeps=EVERY)
            got = e.rows(states)
            err = np.abs(got[:, :N, :8] - S[:, :N, :8])
            err[:, :, 2] = np.abs((err[:, :, 2] + np.pi) % (2 * np.pi) - np.pi)  # headings as angles
            rel = err / np.maximum(np.abs(S[:, :N, :8]), 1.0)
            row += [float(rel.max()), float(np.median(rel.reshape(E, -1).max(1)))]
        rows.append(row)
    print(f"{tag}: first body contact (oracle) at step {first_contact}; final max rel err fp64 {rows[-1][3]:.3e} fp32 {rows[-1][5]:.3e}")


rows = []
sc = scenarios.circular_crossing(1, 5, seed0=1002)
run("config1_cc5_hsfm_farina", "hsfm_farina", sc["states"], sc["goals"], None, False, rows)
run("config1_cc5_hsfm_new_guo", "hsfm_new_guo", sc["states"], sc["goals"], None, False, rows)
sc = scenarios.ccso_synthetic(64, 25, seed0=2000)
walls = scenarios.pack_walls(scenarios.EXAMPLE_WALLS)
run("config3b_64x25_hsfm_farina_walls_robot", "hsfm_farina", np.concatenate([sc["states"], sc["robot"][:, None]], 1), sc["goals"], walls, True, rows)
out = os.path.join(ROOT, "profiles", f"{R}_divergence.csv")
with open(out, "w") as f:
    w = csv.writer(f)
    w.writerow(["case", "sub_step", "min_body_gap_m(oracle)", "fp64_max_rel_err", "fp64_median_env_max_rel_err", "fp32_max_rel_err",
                "fp32_median_env_max_rel_err"])
    w.writerows(rows)
print("wrote", out)
