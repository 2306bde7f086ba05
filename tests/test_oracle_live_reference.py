"""The oracle against the LIVE reference, on seeds no committed fixture holds (build container only: /root/reference does not
exist on the GPU box, where this module is skipped).  tests/golden/*.npz pin the oracle on recorded runs; this re-derives the pin
from the reference itself every time the CPU suite runs here, on fresh circular-crossing crowds of every model family."""
import os
import sys

import numpy as np
import pytest

import oracle
from oracle import OracleConfig
from helpers import rel_err

REF = "/root/reference"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "social_gym")), reason="the live reference exists in the build container only")


@pytest.fixture(scope="module")
def mg():
    sys.path.insert(0, GOLDEN)
    import make_golden  # installs the pygame / gymnasium / rvo2 stubs (ref_shim) and imports the reference
    return make_golden


@pytest.mark.parametrize("model,seed,n,visible", [("hsfm_farina", 4101, 6, True), ("sfm_guo", 4102, 9, False), ("hsfm_new_guo", 4103, 7, True),
                                                 ("sfm_helbing", 4104, 12, True), ("hsfm_guo", 4105, 5, False)])
def test_serial_update_matches_the_live_reference(mg, model, seed, n, visible):
    """MotionModelManager.update_humans (serial Euler path, mmm:369-373) x 80 with the robot moved like RobotAgent.step does."""
    sim = mg.cc_sim(model, seed, n, robot_visible=visible)
    mm, humans = sim.motion_model_manager, sim.humans
    rv, dt = np.array([0.2, 0.5]), mg.DT
    S = np.array([h.get_safe_state() for h in humans])
    if visible:
        S = np.concatenate([S, sim.robot.get_safe_state()[None]], 0)
    G = mg.pack_goals(humans)[None]
    params = np.array([h.get_parameters(model) for h in humans])[None]
    safety = np.array([h.safety_space for h in humans] + ([sim.robot.safety_space] if visible else []), np.float64)[None]
    cfg = OracleConfig(mg.SFMS.index(model), visible, bool(mm.all_equal_humans), False)
    S, D = S[None], np.zeros((1, n, 2))
    worst = 0.0
    for step in range(80):
        sim.robot.position = sim.robot.position + rv * dt
        sim.robot.linear_velocity = rv.copy()
        mm.update_humans(0.0, dt)
        S, G, D = oracle.update_humans(cfg, S, G, None, params, safety, D, dt, 1, robot_vel=rv[None] if visible else None)
        ref = np.array([mg.human_row(h) for h in humans])
        got = np.concatenate([S[0, :n, :8], S[0, :n, 10:12], D[0]], 1)
        worst = max(worst, rel_err(got, ref).max())
    assert worst < 1e-9, worst
