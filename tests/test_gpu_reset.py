"""On-device reset (snp_reset): the reference's scenario generators replayed per env on NumPy's MT19937 stream.
Against the outputs recorded from the live reference (tests/golden/scenarios.npz), against the host generators at batch size,
and the gym-level reset / restart of finished episodes."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN

pytestmark = pytest.mark.gpu

CASES = [("cc5", "circle_crossing"), ("cc25", "circle_crossing"), ("cc7_randattr", "circle_crossing"), ("pt5", "parallel_traffic"),
         ("pt12_randattr", "parallel_traffic"), ("ccso6", "circular_crossing_with_static_obstacles"),
         ("ccso8", "circular_crossing_with_static_obstacles")]


def _engine(E, N, dtype=torch.float64, model="hsfm_farina"):
    from social_navigation_pyenvs_b200 import CrowdEngine
    return CrowdEngine(model, E, N, G=2, dtype=dtype, has_robot=True)


def _rows(eng):
    """Engine state after a reset in the reference's row layout [E,N,13] + goals [E,N,2,2] + goal counts."""
    from social_navigation_pyenvs_b200 import _lib as L
    d, s = eng.dyn.double().cpu().numpy(), eng.stat.double().cpu().numpy()
    g = eng.goals.double().cpu().numpy()  # [G,2,E,N]
    rows = np.zeros((eng.E, eng.N, 13))
    rows[..., 0], rows[..., 1], rows[..., 2] = d[L.DYN_PX], d[L.DYN_PY], d[L.DYN_TH]
    rows[..., 3], rows[..., 4], rows[..., 5], rows[..., 6], rows[..., 7] = d[L.DYN_VX], d[L.DYN_VY], d[L.DYN_BVX], d[L.DYN_BVY], d[L.DYN_OM]
    rows[..., 8], rows[..., 9], rows[..., 12] = s[L.STAT_R], s[L.STAT_M], s[L.STAT_VD]
    rows[..., 10], rows[..., 11] = g[0, 0], g[0, 1]
    return rows, g.transpose(2, 3, 0, 1), eng.goal_cnt.cpu().numpy()


@pytest.mark.parametrize("key,scenario", CASES)
def test_device_reset_vs_reference_golden(key, scenario):
    """Same seeds as the live reference: identical number of draws (same accept / reject decisions), attributes bit-exact,
    positions / headings / goals to 1e-13 (device cos / sin vs libm)."""
    z = np.load(os.path.join(GOLDEN, "scenarios.npz"))
    S, G, D, seeds = z[key + "_states"], z[key + "_goals"], z[key + "_draws"], z[key + "_seeds"]
    E, N = S.shape[:2]
    eng = _engine(E, N)
    eng.time_now.fill_(3.0)
    scen, draws = eng.reset_scenario(scenario, seeds=seeds, randomize_attributes="randattr" in key)
    assert np.array_equal(draws.cpu().numpy(), D)
    rows, goals, cnt = _rows(eng)
    assert np.array_equal(rows[..., [8, 9, 12]], S[..., [8, 9, 12]])                 # radius, mass, desired speed
    assert np.abs(rows - S).max() < 1e-13
    assert np.abs(goals[:, :, :G.shape[2]] - G).max() < 1e-13 and (cnt == G.shape[2]).all()
    assert (eng.time_now == 0).all() and (eng.goal_idx == 0).all()
    from social_navigation_pyenvs_b200 import scenarios
    ref_robot = scenarios.parallel_traffic(1, 2)["robot"][0] if scenario == "parallel_traffic" else scenarios.robot_rows(1)[0]
    rr, _ = eng.robot_rows()
    assert np.array_equal(rr[:, [0, 1, 2, 8, 9, 10, 11, 12]], np.tile(ref_robot[[0, 1, 2, 8, 9, 10, 11, 12]], (E, 1)))
    assert np.all(rr[:, 3:8] == 0)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_device_reset_vs_host_generators_at_batch_size(dtype):
    from social_navigation_pyenvs_b200 import scenarios
    tol = 1e-12 if dtype == torch.float64 else 1e-6
    for name, host, E, N in [("circle_crossing", scenarios.circular_crossing, 600, 25), ("ccso_synthetic", scenarios.ccso_synthetic, 300, 25),
                             ("parallel_traffic", scenarios.parallel_traffic, 600, 10)]:
        eng = _engine(E, N, dtype)
        eng.reset_scenario(name, seed0=2000)
        sc = host(E, N, 2000)
        rows, goals, cnt = _rows(eng)
        assert np.abs(rows - sc["states"]).max() < tol, name
        assert np.abs(goals[:, :, :sc["goals"].shape[2]] - sc["goals"]).max() < tol, name
        if dtype == torch.float32:   # the generator runs in double on the device too: the fp32 state is the ROUNDED double result
            assert np.array_equal(rows[..., 0].astype(np.float32), sc["states"][..., 0].astype(np.float32)) or \
                np.mean(rows[..., 0].astype(np.float32) == sc["states"][..., 0].astype(np.float32)) > 0.999


def test_hybrid_scenario_coin_respawn_and_masked_reset():
    """social_nav_gym.py:155-167: np.random.choice picks the generator per env, then the stream is re-seeded.  A mixed batch steps
    like the two pure batches it is made of (parallel-traffic envs respawn, the others do not); a masked reset leaves the other
    envs untouched bit for bit."""
    z = np.load(os.path.join(GOLDEN, "scenarios.npz"))
    seeds = z["hybrid_seeds"]
    E, N = len(seeds), 5
    hyb, cc, pt = _engine(E, N), _engine(E, N), _engine(E, N)
    scen, _ = hyb.reset_scenario("hybrid_scenario", seeds=seeds)
    coin = scen.cpu().numpy()
    assert np.array_equal(coin, z["hybrid_choice"]) and 0 < coin.sum() < E
    cc.reset_scenario("circle_crossing", seeds=seeds)
    pt.reset_scenario("parallel_traffic", seeds=seeds)
    assert hyb.respawn_envs is not None and cc.respawn_bounds is None and pt.respawn_bounds == (7.0, 1.5)
    act = torch.zeros((2, E), dtype=torch.float64, device="cuda")
    act[0] = 0.3
    for eng in (hyb, cc, pt):
        eng.action.copy_(act)
    for _ in range(40):   # 40 x 20 sub-steps = 10 s: parallel-traffic humans reach the left end and respawn
        for eng in (hyb, cc, pt):
            eng.step(None, 0.0125, n_substeps=20, pre_checks=True)
    sel = torch.as_tensor(coin == 1, device="cuda")
    assert torch.equal(hyb.dyn[:, sel], pt.dyn[:, sel]) and torch.equal(hyb.dyn[:, ~sel], cc.dyn[:, ~sel])
    assert torch.equal(hyb.flags[sel], pt.flags[sel]) and torch.equal(hyb.flags[~sel], cc.flags[~sel])
    assert (pt.goal_cnt == 1).all() and (pt.dyn[0].max() > 6.0)       # somebody respawned at the right end
    # masked reset: every third env restarts with a new seed, the rest is untouched
    before = {k: getattr(hyb, k).clone() for k in ("dyn", "stat", "goals", "goal_idx", "goal_cnt", "robot", "time_now")}
    mask = torch.arange(E, device="cuda") % 3 == 0
    hyb.reset_scenario("circle_crossing", seeds=seeds + 500, mask=mask)
    fresh = _engine(E, N)
    fresh.reset_scenario("circle_crossing", seeds=seeds + 500)
    for k, old in before.items():
        new = getattr(hyb, k)
        if k in ("time_now", "goal_idx", "goal_cnt"):
            assert torch.equal(new.reshape(E, -1)[~mask], old.reshape(E, -1)[~mask]), k
        elif k == "robot":
            assert torch.equal(new[:, ~mask], old[:, ~mask])
        else:
            assert torch.equal(new[..., ~mask, :], old[..., ~mask, :]), k
    assert torch.equal(hyb.dyn[..., mask, :], fresh.dyn[..., mask, :]) and (hyb.time_now[mask] == 0).all() and (hyb.time_now[~mask] > 0).all()


def test_gym_reset_on_device_matches_host_reset_and_restarts_finished_envs():
    from social_navigation_pyenvs_b200.social_nav_gym import BatchedSocialNavGym
    for sim, n in [("circle_crossing", 5), ("parallel_traffic", 6), ("circular_crossing_with_static_obstacles", 8),
                   ("circular_crossing_with_static_obstacles", 25)]:
        a, b = BatchedSocialNavGym(64), BatchedSocialNavGym(64)
        for g in (a, b):
            g.configure(dict(human_policy="hsfm_new_guo", human_num=n, test_sim=sim, train_val_sim=sim))
        oa, _ = a.reset("test", test_case=7, on_device=True)
        ob, _ = b.reset("test", test_case=7, on_device=False)
        assert np.abs(oa - ob).max() < 1e-12 and a.case_counter["test"] == b.case_counter["test"] == 71
        act = np.tile([0.0, 1.0], (64, 1))
        for _ in range(3):
            ra, rb = a.step(act), b.step(act)
        assert np.abs(ra[0] - rb[0]).max() < 1e-9 and np.array_equal(ra[4], rb[4])
    # restart of finished episodes: the k-th finished env gets case counter + k
    g = BatchedSocialNavGym(32)
    g.configure(dict(human_policy="sfm_helbing", human_num=5))
    g.reset("train", test_case=100)
    fin = np.zeros(32, bool)
    fin[[3, 4, 20]] = True
    keep = g.engine.dyn.clone()
    g.reset_finished(fin, "train")
    assert g.case_counter["train"] == 100 + 32 + 3
    ref = BatchedSocialNavGym(3)
    ref.configure(dict(human_policy="sfm_helbing", human_num=5))
    ref.reset("train", test_case=132)
    assert torch.equal(g.engine.dyn[:, [3, 4, 20]], ref.engine.dyn) and torch.equal(g.engine.dyn[:, ~torch.as_tensor(fin)], keep[:, ~torch.as_tensor(fin)].to(g.engine.dyn.device))
