"""The object-level and gym-level host mirrors on the GPU: MotionModelManager on reference-shaped agent objects, and
BatchedSocialNavGym against the recorded SocialNavGym.step sequence."""
import os
import types

import numpy as np
import pytest
import torch

from helpers import GOLDEN, load_traj, rel_err

pytestmark = pytest.mark.gpu


def _agents(d, with_robot):
    humans = []
    for i in range(d["n"]):
        s = d["states0"][i]
        goals = [list(g) for g in d["goals0"][i] if not np.isnan(g[0])]
        humans.append(types.SimpleNamespace(position=s[0:2].copy(), yaw=float(s[2]), linear_velocity=s[3:5].copy(), body_velocity=s[5:7].copy(),
                                            angular_velocity=float(s[7]), radius=float(s[8]), mass=float(s[9]), desired_speed=float(s[12]),
                                            goals=goals, safety_space=float(d["safety"][i]), desired_force=np.zeros(2)))
    r = d["robot0"]
    robot = types.SimpleNamespace(position=r[0:2].copy(), yaw=float(r[2]), linear_velocity=r[3:5].copy(), body_velocity=r[5:7].copy(),
                                  angular_velocity=float(r[7]), radius=float(r[8]), mass=float(r[9]), desired_speed=float(r[12]),
                                  goals=[[float(r[10]), float(r[11])]], safety_space=float(d["safety"][-1])) if with_robot else None
    return humans, robot


@pytest.mark.parametrize("name", ["cc6_robot_hsfm_new_guo", "walls7_sfm_guo", "corridor_hsfm_new", "jym_sfm_helbing"])
def test_motion_model_manager_on_agent_objects(name):
    from social_navigation_pyenvs_b200.motion_model_manager import MotionModelManager
    from social_navigation_pyenvs_b200 import SFMS
    d = load_traj(name)
    humans, robot = _agents(d, d["consider_robot"])
    walls = [types.SimpleNamespace(segments={k: [list(s[0]), list(s[1])] for k, s in enumerate(w) if not np.isnan(s[0, 0])}) for w in d["walls"]]
    mm = MotionModelManager(SFMS[int(d["type"])], d["consider_robot"], False, humans, robot, walls)
    assert mm.all_equal_humans == d["all_equal"] and mm.headed == (int(d["type"]) >= 3)
    rv, dt, cur = d["robot_vel"], float(d["dt"]), 0
    for k, s in enumerate(d["steps"][:14]):
        while cur < s:
            if robot is not None:
                robot.position = robot.position + rv * dt
                robot.linear_velocity = rv.copy()
            mm.update_humans(0.0, dt)
            cur += 1
        got = mm.get_human_states(include_goal=True, headed=False)
        ref = d["traj"][k][:, [0, 1, 2, 3, 4, 7, 8, 9]]
        assert rel_err(got, ref).max() < 1e-9, (name, int(s))
    # peek leaves pose / velocity / goal untouched (mmm:691-709)
    before = mm.get_human_states(include_goal=True, headed=mm.headed)
    nxt = mm.get_next_human_observable_states(0.25)
    assert nxt.shape == (d["n"], 4) and np.array_equal(mm.get_human_states(include_goal=True, headed=mm.headed), before)
    with pytest.raises(NotImplementedError):
        MotionModelManager("orca", False, False, humans, robot, walls)
    with pytest.raises(Exception):
        MotionModelManager("no_such_model", False, False, humans, robot, walls)


def test_batched_gym_env0_equals_recorded_reference_episode():
    """reset(phase='test', test_case=3) + 60 x step(action): env 0 of a 3-env batch must reproduce the reference's recorded
    episode (same scenario from the same seed, same rewards / info codes, same observations)."""
    from social_navigation_pyenvs_b200.social_nav_gym import BatchedSocialNavGym
    z = np.load(os.path.join(GOLDEN, "gym_step.npz"))
    for key, model, visible in [("hsfm_farina_0", "hsfm_farina", False), ("sfm_helbing_1", "sfm_helbing", True)]:
        env = BatchedSocialNavGym(3)
        env.configure(dict(human_policy=model, human_num=5, robot_visible=visible))
        ob, info = env.reset(phase="test", test_case=3)
        assert np.array_equal(ob[0, :, :2], z[key + "_states0"][:, 0:2])
        for k, a in enumerate(z[key + "_actions"]):
            acts = np.tile(a, (3, 1))
            acts[1:] *= 0.5                                                     # the other envs do something else
            ob, reward, term, trunc, info = env.step(acts)
            ref = z[key + "_result"][k]
            assert term[0] == bool(ref[1]) and trunc[0] == bool(ref[2]) and info[0] == int(ref[3]) and abs(reward[0] - ref[0]) < 1e-9, (key, k)
            assert rel_err(ob[0], z[key + "_obs"][k]).max() < 1e-9, (key, k)
        assert abs(env.global_time[0] - 15.0) < 1e-9
        col, dmin, goal = env.check_actual_collisions_and_goal()
        assert col.shape == (3,) and dmin.shape == (3,)


def test_robot_driven_by_motion_model_vs_reference_golden():
    """imitation-learning sub-step loop (update_robot then update_humans, gym:260-265; mmm:593-653) inside ONE launch per block of
    sub-steps, against 6 runs recorded from the live reference: robot and human models may differ, the robot may be invisible to
    the humans, walls, and one run where the robot's goal list rotates."""
    from social_navigation_pyenvs_b200 import CrowdEngine, SFMS
    z = np.load(os.path.join(GOLDEN, "il_robot.npz"))
    keys = sorted(k[:-8] for k in z.files if k.endswith("_states0"))
    assert len(keys) == 6
    for key in keys:
        vis, equal = (bool(v) for v in z[key + "_flags"])
        S, G, rb = z[key + "_states0"], z[key + "_goals0"][None], z[key + "_robot0"][None]
        n = S.shape[0]
        S1 = (np.concatenate([S, rb], 0) if vis else S)[None]
        eng = CrowdEngine.from_reference_arrays(SFMS[int(z[key + "_type"])], S1, G, walls=z[key + "_walls"], consider_robot=vis,
                                                all_params_equal=equal, robot=None if vis else rb)
        eng.set_robot_motion_model(SFMS[int(z[key + "_robot_type"])], goals=z[key + "_robot_goals"][None])
        assert np.array_equal(eng.robot_params, z[key + "_robot_params"])
        moussaid = int(z[key + "_type"]) % 3 == 2 or int(z[key + "_robot_type"]) % 3 == 2
        cur = 0
        for k, s_ in enumerate(z[key + "_steps"]):
            if s_ > cur:
                eng.imitation_learning_step(0.0125, n_substeps=int(s_ - cur))
                cur = s_
            got_h = eng.rows(S1)[0]
            rr, rdf = eng.robot_rows()
            ref_h, ref_r = z[key + "_traj"][k], z[key + "_robot_traj"][k]
            tol = 1e-5 if moussaid else 1e-8
            assert rel_err(np.concatenate([got_h[:n, :8], got_h[:n, 10:12]], 1), ref_h[:, :10]).max() < tol, (key, int(s_))
            got_r = np.concatenate([rr[0, :8], rr[0, 10:12]])
            got_r[2] = ref_r[2] + (got_r[2] - ref_r[2] + np.pi) % (2 * np.pi) - np.pi
            assert rel_err(got_r, ref_r[:10]).max() < tol, (key, int(s_))
            assert rel_err(rdf[0], ref_r[10:12], scale=100.0).max() < tol
        f = eng.decode_flags()
        assert f["info"][0] in (0, 2, 3, 4)
