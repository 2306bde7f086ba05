"""Host-side scenario generators against what the live reference produced (recorded in tests/golden/gym_step.npz and the
trajectory fixtures).  CPU only."""
import os

import numpy as np

from helpers import GOLDEN, load_traj
from social_navigation_pyenvs_b200 import scenarios


def test_circular_crossing_replays_the_reference_generator():
    """SocialNavGym.reset(phase='test', test_case=3) seeds np.random with 1003 (gym:135-137) and calls
    generate_circular_crossing_setting (sim:200-299): same humans, bit for bit."""
    z = np.load(os.path.join(GOLDEN, "gym_step.npz"))
    sc = scenarios.circular_crossing(1, 5, seed0=1003)
    assert np.array_equal(sc["states"][0], z["hsfm_farina_0_states0"])
    assert np.array_equal(sc["goals"][0], z["hsfm_farina_0_goals0"])
    # and the seed-1002 / seed-2000 crowds of the trajectory fixtures (SocialNavSim(..., scenario="circular_crossing"))
    d = load_traj("cc5_hsfm_farina")
    sc = scenarios.circular_crossing(1, 5, seed0=1002)
    assert np.array_equal(sc["states"][0], d["states0"]) and np.array_equal(sc["goals"][0], d["goals0"])
    d = load_traj("cc25_robot_hsfm_farina")
    sc = scenarios.circular_crossing(1, 25, seed0=2000)
    assert np.array_equal(sc["states"][0], d["states0"])
    assert np.array_equal(sc["robot"][0, [0, 1, 2, 8, 9, 10, 11, 12]], d["robot0"][[0, 1, 2, 8, 9, 10, 11, 12]])


def test_parallel_traffic_replays_the_reference_generator():
    for name, seed, n in [("pt5_robot_hsfm_farina", 2004, 5), ("pt7_sfm_helbing", 2011, 7), ("pt10_robot_sfm_guo", 77, 10)]:
        d = load_traj(name)
        sc = scenarios.parallel_traffic(1, n, seed0=seed)
        assert np.array_equal(sc["states"][0], d["states0"]) and np.array_equal(sc["goals"][0], d["goals0"]), name
        assert sc["respawn_bounds"] == d["respawn_bounds"]
        assert np.array_equal(sc["robot"][0, [0, 1, 2, 8, 9, 10, 11, 12]], d["robot0"][[0, 1, 2, 8, 9, 10, 11, 12]])


def test_batches_are_seeded_per_env_and_pool_path_matches_serial_path():
    a = scenarios.circular_crossing(3, 5, seed0=2000)
    b = scenarios.circular_crossing(1, 5, seed0=2002)
    assert np.array_equal(a["states"][2], b["states"][0])
    big = scenarios.ccso_synthetic(260, 9, seed0=50)          # >= 256 envs -> worker pool
    small = scenarios.ccso_synthetic(2, 9, seed0=50 + 258)
    assert np.array_equal(big["states"][258:260], small["states"])


def test_ccso_synthetic_respects_clearances():
    sc = scenarios.ccso_synthetic(8, 25, seed0=2000)
    S = sc["states"]
    assert np.all(S[:, :3, 12] == 0) and np.all(S[:, 3:, 12] == 1.0)            # three static obstacles per env
    assert np.all((S[:, :3, 8] > 0.6) & (S[:, :3, 8] <= 1.0)) and np.all(S[:, 3:, 8] == 0.3)
    for e in range(8):
        p, r = S[e, :, 0:2], S[e, :, 8]
        d = np.linalg.norm(p[:, None] - p[None], axis=-1) - r[:, None] - r[None]
        assert (d[np.triu_indices(25, 1)] >= 0.2 - 1e-12).all()
    crowd = scenarios.jittered_grid_crowd(16)
    p = crowd["states"][0, :, 0:2]
    d = np.linalg.norm(p[:, None] - p[None], axis=-1)
    assert d[np.triu_indices(256, 1)].min() > 0.6                                # nobody overlaps at the start
