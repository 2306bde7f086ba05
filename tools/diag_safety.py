import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, oracle
from oracle import OracleConfig
from social_navigation_pyenvs_b200 import CrowdEngine, scenarios, _lib as L
from helpers import rel_err
np.set_printoptions(precision=6, linewidth=200)
E, N, s, dt, k = 6, 7, 0.15, 0.0125, 20
sc = scenarios.circular_crossing(E, N, seed0=77)
sc["states"][:, :, 0:2] *= 0.5; sc["goals"] *= 0.5
robot = sc["robot"].copy(); robot[:, 1] = -1.0; robot[:, 11] = 1.0
params = np.tile(oracle.default_params("hsfm_new_guo"), (E, N, 1))
action = np.tile([0.0, 1.0], (E, 1))
states = np.concatenate([sc["states"], robot[:, None]], 1)
for kk in (1, 5, 10, 20):
    eng = CrowdEngine.from_reference_arrays("hsfm_new_guo", states, sc["goals"], consider_robot=True, all_params_equal=True)
    eng.set_safety_space(s)
    eng.step(action, dt, n_substeps=kk, pre_checks=True)
    safety = np.full((E, N + 1), 0.01 + s); safety[:, N] = 0.0
    cfg = OracleConfig(oracle.type_code("hsfm_new_guo"), True, True, False)
    ref, _, _ = oracle.update_humans(cfg, states, sc["goals"], None, params, safety, np.zeros((E, N, 2)), dt, kk, robot_vel=action)
    got = eng.rows(states)
    err = rel_err(got[:, :N, :8], ref[:, :N, :8])
    idx = np.unravel_index(err.argmax(), err.shape)
    print(kk, "max err", err.max(), "at", idx, "got", got[idx[0], idx[1], :8], "ref", ref[idx[0], idx[1], :8])
