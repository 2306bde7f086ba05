"""The CUDA path next to the LIVE reference on the GPU box (-m gpu): fresh crowds that no committed fixture holds are stepped by
the reference's own MotionModelManager.update_humans (serial Python / NumPy path, mmm:354-373) and by snp_step side by side.
The reference is the copy staged under oracle/_ref by oracle/build.py::stage_reference (or /root/reference in the build
container); the oracle port is not involved."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from oracle import reference

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not reference.available(), reason="live reference not staged (oracle/build.py)")]

DT = 0.0125


def _crowd(seed, n, walls, model):
    from social_navigation_pyenvs_b200 import scenarios
    sc = scenarios.ccso_synthetic(1, n, seed) if walls else scenarios.circular_crossing(1, n, seed)
    if not walls:  # bring the crowd together so that the pair forces matter from the first step
        sc["states"][:, :, 0:2] *= 0.45
        sc["goals"] *= 0.45
    robot = sc["robot"][0].copy()
    robot[1] = -3.0
    return sc["states"][0], sc["goals"][0], robot, (scenarios.EXAMPLE_WALLS if walls else None)


@pytest.mark.parametrize("model,seed,n,walls,visible", [("hsfm_farina", 9101, 25, True, True), ("sfm_helbing", 9102, 5, False, False),
                                                        ("hsfm_new_guo", 9106, 6, False, True), ("sfm_guo", 9104, 12, False, True),
                                                        ("hsfm_guo", 9105, 7, True, False)])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_step_by_step_against_the_live_reference(model, seed, n, walls, visible, dtype):
    """60 updates; before EVERY update the engine is loaded with the reference's current state, so each comparison is a single
    step from identical inputs (north_star's parity bar: 1e-9 relative in fp64, 1e-4 in fp32).  A free-running engine is
    compared as well (multi-step divergence reported through the looser bound).  (hsfm_new* crowds with zero-speed static humans
    are left out: the reference's own explicit torque update overshoots there -- |omega| reaches 1e91 within 60 updates -- so a
    multi-step comparison measures the instability, not the implementation; tests/golden/sim_update.npz holds such a run.)"""
    from social_navigation_pyenvs_b200 import CrowdEngine, scenarios
    S0, G0, robot, wl = _crowd(seed, n, walls, model)
    sim = reference.sim_from_arrays(model, S0, G0, wl, robot, visible, DT)
    mm = sim.motion_model_manager
    rv = np.array([0.3, 0.8])
    packed = None if wl is None else scenarios.pack_walls(wl)
    rows = lambda: np.concatenate([np.array([h.get_safe_state() for h in sim.humans]), sim.robot.get_safe_state()[None]], 0)[None]
    st = rows()
    kw = dict(walls=packed, consider_robot=visible, all_params_equal=bool(mm.all_equal_humans), dtype=dtype)
    eng = CrowdEngine.from_reference_arrays(model, st if visible else st[:, :n], G0[None], robot=None if visible else st[:, n], **kw)
    free = CrowdEngine.from_reference_arrays(model, st if visible else st[:, :n], G0[None], robot=None if visible else st[:, n], **kw)
    tol = 1e-9 if dtype == torch.float64 else 1e-4
    worst, worst_free = 0.0, 0.0
    for step in range(60):
        st = rows()
        df = np.array([h.desired_force for h in sim.humans])[None]
        eng.load_rows(st if visible else st[:, :n])
        if not visible:
            eng.set_robot_rows(st[:, n])
        eng.load_goals(np.array([[g for g in h.goals] for h in sim.humans], np.float64)[None])
        eng.set_desired_force(df)
        reference.step_like_gym(sim, rv, DT, 1)
        for e in (eng, free):
            e.step(rv[None], DT, n_substeps=1, pre_checks=False)
        ref = reference.human_rows(sim)
        got = eng.rows(st if visible else st[:, :n])[0]
        cmp_ = lambda g: rel_err(np.concatenate([g[:n, :8], g[:n, 10:12]], 1), ref[:, :10]).max()
        worst = max(worst, cmp_(got))
        worst_free = max(worst_free, cmp_(free.rows(st if visible else st[:, :n])[0]))
        if dtype == torch.float64:
            assert rel_err(eng.desired_force()[0], ref[:, 10:12], scale=100.0).max() < tol
    assert worst < tol, worst
    assert worst_free < (1e-7 if dtype == torch.float64 else 5e-2), worst_free


def test_gym_checks_against_the_live_reference():
    """collision_detection_and_reaching_goal + compute_reward_and_infos (sim:949-1029) and check_actual_collisions_and_goal
    (gym:107-118) of the live reference vs snp_checks on 60 random robot placements: every flag, dmin and reward bit-identical."""
    from social_navigation_pyenvs_b200 import CrowdEngine
    import social_gym.social_nav_gym as gym_mod
    S0, G0, robot, _ = _crowd(9201, 8, False, "sfm_guo")
    sim = reference.sim_from_arrays("sfm_guo", S0, G0, None, robot, False, DT)
    sim.time_limit, sim.collision_penalty, sim.success_reward, sim.discomfort_dist, sim.discomfort_penalty_factor = 50, -0.25, 1.0, 0.2, 0.5
    rng = np.random.RandomState(5)
    code = {"Timeout": 1, "Collision": 2, "Reaching goal": 3, "Too close": 4, "": 0}
    seen = set()
    for k in range(60):
        reference.step_like_gym(sim, (0.0, 0.0), DT, 3)
        h0 = sim.humans[rng.randint(8)]
        sim.robot.position = (h0.position + rng.uniform(-1.0, 1.0, 2)) if k % 2 else (np.array(sim.robot.goals[0], np.float64) + rng.uniform(-0.5, 0.5, 2))
        a = rng.uniform(-1.0, 1.0, 2)
        t_now = 49.5 if k % 17 == 0 else float(rng.uniform(0, 40))
        col, dmin, goal = sim.collision_detection_and_reaching_goal(a, 0.25)
        reward, term, trunc, info = sim.compute_reward_and_infos(col, dmin, goal, t_now, 0.25)
        acol, admin, agoal = gym_mod.SocialNavGym.check_actual_collisions_and_goal(sim)
        H = np.array([h.get_safe_state() for h in sim.humans])[None]
        eng = CrowdEngine.from_reference_arrays("sfm_guo", H, G0[None], consider_robot=False, robot=sim.robot.get_safe_state()[None])
        eng.time_now.fill_(t_now)
        f = eng.run_checks(a[None], pre=True, post=True)
        got = [bool(f["collision"][0]), float(f["dmin"][0]), bool(f["reaching_goal"][0]), float(f["reward"][0]), bool(f["terminated"][0]),
               bool(f["truncated"][0]), int(f["info"][0]), bool(f["actual_collision"][0]), float(f["actual_dmin"][0]), bool(f["actual_goal"][0])]
        ref = [bool(col), float(dmin), bool(goal), float(reward), bool(term), bool(trunc), code[str(info)], bool(acol), float(admin), bool(agoal)]
        assert got == ref, (k, got, ref)
        seen.add(code[str(info)])
    assert len(seen) >= 3
