"""Host-side scenario generators against what the live reference produced (recorded in tests/golden/gym_step.npz and the
trajectory fixtures).  CPU only."""
import os

import numpy as np
import pytest

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


def _golden_cases():
    z = np.load(os.path.join(GOLDEN, "scenarios.npz"))
    return z, [("cc5", 0), ("cc25", 0), ("cc7_randattr", 0), ("pt5", 1), ("pt12_randattr", 1), ("ccso6", 2), ("ccso8", 2)]


def test_host_generators_match_the_reference_on_many_seeds():
    """scenarios.py against tests/golden/scenarios.npz (recorded from the live reference's generators, social_nav_sim.py:200-431):
    every seed, bit for bit -- including randomized attributes and the reference's own static-obstacle generator."""
    z, cases = _golden_cases()
    for key, scen in cases:
        S, G, seeds = z[key + "_states"], z[key + "_goals"], z[key + "_seeds"]
        n, rand = S.shape[1], "randattr" in key
        for e, seed in enumerate(seeds):
            if scen == 0:
                sc = scenarios.circular_crossing(1, n, seed0=int(seed), randomize_attributes=rand)
            elif scen == 1:
                sc = scenarios.parallel_traffic(1, n, seed0=int(seed), randomize_attributes=rand)
            else:
                sc = scenarios.circular_crossing_with_static_obstacles(1, n, seed0=int(seed))
            assert np.array_equal(sc["states"][0], S[e]), (key, seed)
            assert np.array_equal(sc["goals"][0], G[e], equal_nan=True), (key, seed)
    assert [scenarios.hybrid_choice(s) for s in z["hybrid_seeds"]] == list(z["hybrid_choice"])


def test_reset_core_replays_numpy_stream_and_reference_generators(tmp_path):
    """csrc/snp_reset_core.h -- the code every thread of the CUDA reset kernel runs -- compiled for the host: MT19937 + random_double
    as NumPy's legacy RandomState, and the generators consuming exactly the reference's draws (count recorded from the live
    reference) with bit-identical results."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "reset_core_host"
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-I", os.path.join(root, "social_navigation_pyenvs_b200", "csrc"),
                    os.path.join(root, "tests", "reset_core_host.cpp"), "-o", str(exe)], check=True)

    def run(scen, n, rand, seed0, count):
        out = subprocess.run([str(exe), str(scen), str(n), str(int(rand)), str(seed0), str(count)], capture_output=True, text=True, check=True).stdout.split("\n")
        envs, k = [], 0
        for _ in range(count):
            _, _, sc, dr = out[k].split()
            rows = np.array([[float(x) for x in out[k + 1 + i].split()] for i in range(n)])
            k += n + 1
            envs.append((int(sc), int(dr), rows))
        return envs

    z, cases = _golden_cases()
    for key, scen in cases:
        S, G, D, seeds = z[key + "_states"], z[key + "_goals"], z[key + "_draws"], z[key + "_seeds"]
        envs = run(scen, S.shape[1], "randattr" in key, int(seeds[0]), len(seeds))
        for e, (sc, draws, rows) in enumerate(envs):
            assert sc == scen and draws == D[e], (key, e)
            ref = np.stack([S[e, :, 0], S[e, :, 1], S[e, :, 2], S[e, :, 8], S[e, :, 12], G[e, :, 0, 0], G[e, :, 0, 1]], 1)
            assert np.array_equal(rows[:, :7], ref), (key, e)
            if G.shape[2] > 1:
                assert np.array_equal(rows[:, 7:9], G[e, :, 1]), (key, e)
    assert [e[0] for e in run(4, 5, False, 3000, 64)] == list(z["hybrid_choice"])
    sc = scenarios.ccso_synthetic(3, 25, seed0=2000)
    for e, (_, _, rows) in enumerate(run(3, 25, False, 2000, 3)):
        assert np.array_equal(rows[:, :2], sc["states"][e, :, 0:2]) and np.array_equal(rows[:, 3], sc["states"][e, :, 8])


def test_spatial_order_is_a_permutation_with_compact_runs():
    """scenarios.spatial_order: every human exactly once, and runs of 256 consecutive humans of the 65536-human benchmark crowd
    cover a ~31 m x 31 m patch (row-by-row numbering: a 510 m x 1 m strip) -- what the exact far-tile culling feeds on.  Also on
    a non-uniform crowd and on sizes that do not divide evenly."""
    sc = scenarios.jittered_grid_crowd(256, pitch=2.0, jitter=0.5, seed=0)
    pos = sc["states"][0, :, 0:2]
    perm = scenarios.spatial_order(pos)
    assert np.array_equal(np.sort(perm), np.arange(len(pos)))
    runs = pos[perm].reshape(-1, 256, 2)
    ext = runs.max(1) - runs.min(1)
    assert ext.max() < 32.0
    rows = pos.reshape(-1, 256, 2)
    assert (rows.max(1) - rows.min(1))[:, 1].min() > 500.0
    inv = np.argsort(perm)
    assert np.array_equal(pos[perm][inv], pos)
    rng = np.random.RandomState(0)
    for n in (1, 7, 513, 5000):
        p = rng.normal(0.0, 30.0, (n, 2)) ** 3 / 900.0   # heavy-tailed, clustered near the origin
        q = scenarios.spatial_order(p)
        assert np.array_equal(np.sort(q), np.arange(n))


def test_deal_tiles_gives_every_rank_whole_tiles_from_all_over_the_crowd():
    """scenarios.deal_tiles: a permutation; rank r's contiguous slice (parallel.agent_shard) is exactly the tiles r, r + world, ... of
    the spatial order, each kept whole and in order; every slice reaches across the crowd (all strips in x, at least two distant
    patches of every strip in y) instead of being one compact block."""
    from social_navigation_pyenvs_b200.parallel import agent_shard
    sc = scenarios.jittered_grid_crowd(128, pitch=2.0, jitter=0.5, seed=1)  # 16384 humans: 8 strips of 16 tiles
    pos = sc["states"][0, :, 0:2]
    perm = scenarios.spatial_order(pos)
    n = len(perm)
    assert np.array_equal(scenarios.deal_tiles(perm, 1), perm)
    for world in (2, 4, 8):
        dealt = scenarios.deal_tiles(perm, world)
        assert np.array_equal(np.sort(dealt), np.arange(n))
        tiles = perm.reshape(-1, 128)
        for r in range(world):
            off, cnt = agent_shard(n, r, world)
            assert np.array_equal(dealt[off:off + cnt].reshape(-1, 128), tiles[r::world])
            span = np.ptp(pos[dealt[off:off + cnt]], axis=0)
            full = np.ptp(pos, axis=0)
            assert span[0] > 0.8 * full[0] and span[1] > 0.4 * full[1]
    with pytest.raises(ValueError):
        scenarios.deal_tiles(np.arange(130), 2)
