"""The oracle against the LIVE reference, on seeds no committed fixture holds (the reference is imported from /root/reference in the
build container and from the copy staged under oracle/_ref -- oracle/build.py::stage_reference -- on the GPU box).  tests/golden/*.npz pin the oracle on recorded runs; this re-derives the pin
from the reference itself every time the CPU suite runs here, on fresh circular-crossing crowds of every model family."""
import os
import sys

import numpy as np
import pytest

import oracle
from oracle import OracleConfig
from helpers import rel_err

from oracle import reference

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
pytestmark = pytest.mark.skipif(not reference.available(), reason="the live reference is neither at /root/reference nor staged under oracle/_ref")


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


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_laser_matches_the_live_reference(mg, seed):
    """LaserSensor.get_laser_measurements (sensors.py:53-69) on a walled scene after a random number of updates, from random sensor
    poses: ranges bit-identical, hit indices equal to the replayed first-strict-minimum of the reference loop."""
    rng = np.random.RandomState(seed)
    sim = mg.custom_sim(mg.dense_example_data(), "hsfm_farina", True)
    mm = sim.motion_model_manager
    for _ in range(int(rng.randint(10, 200))):
        mm.update_humans(0.0, mg.DT)
    humans = np.array([[h.position[0], h.position[1], h.radius] for h in sim.humans])
    walls = mg.pack_walls(mm.walls)
    for rep in range(4):
        pos, yaw = rng.uniform(-5.0, 5.0, 2), float(rng.uniform(-np.pi, np.pi))
        samples, span, maxd = int(rng.choice([61, 180, 360])), float(rng.choice([np.pi, 2 * np.pi])), float(rng.choice([6.0, 10.0]))
        sensor = mg.LaserSensor(pos, yaw, span, samples, maxd, uncertainty=None)
        sensor.uncertainty = None
        meas = sensor.get_laser_measurements(sim.humans, mm.walls)
        ranges, hits = oracle.laser(humans[None], walls, np.array([[pos[0], pos[1], yaw]]), span, samples, maxd)
        assert np.array_equal(ranges[0], np.array(list(meas.values())))
        assert np.array_equal(hits[0], mg.laser_hits(sensor, sim.humans, mm.walls))


def test_checks_match_the_live_reference(mg):
    """collision_detection_and_reaching_goal + compute_reward_and_infos (sim:949-1029) and check_actual_collisions_and_goal
    (gym:107-118) on 150 fresh random robot placements / actions / times: every flag, dmin and reward bit-identical."""
    rng = np.random.RandomState(77)
    sim = mg.cc_sim("sfm_guo", 4201, 7, robot_visible=False)
    sim.time_limit, sim.collision_penalty, sim.success_reward, sim.discomfort_dist, sim.discomfort_penalty_factor = 50, -0.25, 1.0, 0.2, 0.5
    consts = np.array([50, -0.25, 1.0, 0.2, 0.5, 0.25])
    mm = sim.motion_model_manager
    code = {"Timeout": 1, "Collision": 2, "Reaching goal": 3, "Too close": 4, "": 0}
    seen = set()
    for k in range(150):
        for _ in range(3):
            mm.update_humans(0.0, mg.DT)
        h0 = sim.humans[rng.randint(7)]
        if k % 3 == 0:
            sim.robot.position = h0.position + rng.uniform(-1.2, 1.2, 2)
        elif k % 3 == 1:
            sim.robot.position = np.array(sim.robot.goals[0], np.float64) + rng.uniform(-0.5, 0.5, 2)
        else:
            sim.robot.position = rng.uniform(-7, 7, 2)
        a = rng.uniform(-1.0, 1.0, 2)
        t_now = 49.5 if k % 29 == 0 else float(rng.uniform(0, 40))
        col, dmin, goal = sim.collision_detection_and_reaching_goal(a, 0.25)
        reward, term, trunc, info = sim.compute_reward_and_infos(col, dmin, goal, t_now, 0.25)
        acol, admin, agoal = mg.gym_mod.SocialNavGym.check_actual_collisions_and_goal(sim)
        H = np.array([h.get_safe_state() for h in sim.humans])[None]
        out = oracle.checks(H, 7, sim.robot.get_safe_state()[None], a[None], np.array([t_now]), consts)[0]
        ref = [float(col), dmin, float(goal), reward, float(term), float(trunc), code[str(info)], float(acol), admin, float(agoal)]
        assert np.array_equal(out[:10], np.array(ref, np.float64)), (k, out[:10], ref)
        seen.add(code[str(info)])
    assert seen == {0, 1, 2, 3, 4}


@pytest.mark.parametrize("model,robot_model,seed,n,visible,robot_dt", [("hsfm_new_guo", "hsfm_farina", 4301, 6, True, 0.1),
                                                                      ("sfm_guo", "sfm_helbing", 4302, 8, True, 0.25),
                                                                      ("hsfm_farina", "hsfm_new", 4303, 5, False, 0.0125),
                                                                      ("sfm_helbing", "hsfm_guo", 4304, 7, True, 0.05)])
def test_sim_update_matches_the_live_reference(mg, model, robot_model, seed, n, visible, robot_dt):
    """SocialNavSim.update (sim:476-529) x 120 with a model-driven robot at its own sampling time, on crowds / model pairs /
    sampling ratios (8, 20, 1, 4) that tests/golden/sim_update.npz does not hold."""
    sim = mg.cc_sim(model, seed, n, robot_visible=visible)
    mm, humans, robot = sim.motion_model_manager, sim.humans, sim.robot
    if len(robot.goals) == 1:
        robot.goals = [list(robot.goals[0]), [float(robot.position[0]), float(robot.position[1])]]
    sim.set_time_step(mg.DT)
    sim.set_robot_time_step(robot_dt)
    sim.set_robot_policy(policy_name=robot_model, runge_kutta=False)
    every = int(round(robot_dt / mg.DT))
    S = np.array([h.get_safe_state() for h in humans])
    rb = robot.get_safe_state()[None].copy()
    if visible:
        S = np.concatenate([S, rb], 0)
    S, G = S[None], mg.pack_goals(humans)[None]
    params = np.array([h.get_parameters(model) for h in humans])[None]
    cfg = OracleConfig(mg.SFMS.index(model), visible, bool(mm.all_equal_humans), False)
    D, rD, rG = np.zeros((1, n, 2)), np.zeros((1, 2)), np.array(robot.goals, np.float64)[None]
    rp, rtype = robot.get_parameters(robot_model), mg.SFMS.index(robot_model)
    worst = 0.0
    for step in range(120):
        sim.update()
        S, G, D, rb, rG, rD = oracle.sim_update_steps(cfg, S, G, None, params, np.zeros((1, S.shape[1])), D, mg.DT, 1, rb, rG, rD, rp, rtype,
                                                      every, robot_dt, phase=step)
        ref_h, ref_r = np.array([mg.human_row(h) for h in humans]), mg.human_row(robot)
        got_h = np.concatenate([S[0, :n, :8], S[0, :n, 10:12], D[0]], 1)
        got_r = np.concatenate([rb[0, :8], rb[0, 10:12], rD[0]])
        worst = max(worst, rel_err(got_h, ref_h).max(), rel_err(got_r, ref_r).max())
    assert worst < 1e-9, worst


def test_unicycle_robot_step_matches_the_live_reference(mg):
    """RobotAgent.step / compute_position with unicycle kinematics (robot_agent.py:116-136) vs tests/helpers.unicycle_step, which the
    GPU test of robot_mode 3 steps the oracle with: bit-identical over 200 random (v, r, dt) calls.  (The reference's own swept
    check cannot be run with an ActionRot: social_nav_sim.py:973 reads robot.theta, which is never assigned -- TypeError.)"""
    from crowd_nav.utils.action import ActionRot
    from helpers import unicycle_step
    sim = mg.cc_sim("sfm_helbing", 4401, 5, robot_visible=True)
    robot = sim.robot
    robot.kinematics = "unicycle"
    rng = np.random.RandomState(3)
    pos, yaw = np.array(robot.position, np.float64), float(robot.yaw)
    for _ in range(200):
        v, r, dt = float(rng.uniform(0, 1.2)), float(rng.uniform(-0.6, 0.6)), float(rng.choice([0.0125, 0.25]))
        want_pos = robot.compute_position(ActionRot(v, r), dt)
        robot.step(ActionRot(v, r), dt)
        pos, yaw, vel = unicycle_step(pos, yaw, v, r, dt)
        assert np.array_equal(want_pos, robot.position) and np.array_equal(pos, robot.position)
        assert yaw == robot.yaw and np.array_equal(vel, robot.linear_velocity)


def test_constant_velocity_propagation_matches_the_live_reference(mg):
    """propagate_humans_state_with_constant_velocity_model (cadrl.py:92-105, the policies' query_env = False branch) vs the NumPy
    restatement the GPU test of CrowdEngine.lookahead(query_env=False) feeds the oracle with: bit-identical."""
    from crowd_nav.policy.cadrl import propagate_humans_state_with_constant_velocity_model as ref_fn
    from helpers import constant_velocity_next
    rng = np.random.RandomState(9)
    for vis in (False, True):
        cur = rng.uniform(-3, 3, (11, 7 if vis else 5))
        assert np.array_equal(ref_fn(cur, 0.25, theta_and_omega_visible=vis), constant_velocity_next(cur, 0.25, vis))
