#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/*.npz by RUNNING THE LIVE REFERENCE.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every fixture is produced by the reference's own code through its own public entry points
(`SocialNavSim`, `MotionModelManager.update_humans`, `update_humans_parallel`, `SocialNavGym.step`,
`collision_detection_and_reaching_goal`, `compute_reward_and_infos`, `LaserSensor`), imported via
`ref_shim.install()`.  Nothing from the reference is copied; only numeric inputs/outputs are
stored.  The fixtures pin the oracle (oracle/) and, through it, the CUDA path.

Fixture layout (all float64 unless noted):
  trajectory cases  traj_<name>.npz
    type            int   index into SFMS (motion_model_manager.py:15-17)
    states0         [N,13]  Agent.get_safe_state rows  (agent.py:256)
    robot0          [13]    robot row (NaN when the env has no robot)
    robot_vel       [2]     constant robot velocity applied as robot.step does (robot_agent.py:126)
    goals0          [N,G,2] NaN padded goal lists
    walls           [W,S,2,2] NaN padded, endpoints sorted as obstacle.py:31-32
    params          [N,20]  Agent.get_parameters (agent.py:268)
    safety          [N+1]
    flags           int[4]  consider_robot, all_equal_humans, n_steps, save_every
    dt              scalar
    steps           int[K]  number of updates performed before each saved row
    traj            [K,N,12] px,py,yaw,vx,vy,bvx,bvy,omega,gx,gy,desired_fx,desired_fy
    robot_traj      [K,2]
    forces1         [N,9]   desired(2) obstacle(2) social(2) torque global(2) after the FIRST update
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()

from social_gym.social_nav_sim import SocialNavSim  # noqa: E402
from social_gym.src.motion_model_manager import SFMS  # noqa: E402
from social_gym.src.sensors import LaserSensor  # noqa: E402
from social_gym.src.forces_parallel import update_humans_parallel  # noqa: E402
import social_gym.social_nav_gym as gym_mod  # noqa: E402
from social_gym.custom_config import config_example, config_corridor, config_socialjym_cc  # noqa: E402
from crowd_nav.utils.action import ActionXY  # noqa: E402
from crowd_nav.policy_no_train.policy_factory import policy_factory  # noqa: E402
from social_gym.src.robot_agent import RobotAgent  # noqa: E402
import copy  # noqa: E402
import configparser  # noqa: E402

DT = 0.0125


def pack_goals(humans):
    gmax = max(len(h.goals) for h in humans)
    out = np.full((len(humans), gmax, 2), np.nan)
    for i, h in enumerate(humans):
        for j, g in enumerate(h.goals):
            out[i, j] = g
    return out


def pack_walls(walls):
    if len(walls) == 0:
        return np.zeros((0, 1, 2, 2))
    smax = max(len(w.segments) for w in walls)
    out = np.full((len(walls), smax, 2, 2), np.nan)
    for i, w in enumerate(walls):
        for j, seg in w.segments.items():
            out[i, j, 0] = seg[0]
            out[i, j, 1] = seg[1]
    return out


def human_row(h):
    return np.array([h.position[0], h.position[1], h.yaw, h.linear_velocity[0], h.linear_velocity[1],
                     h.body_velocity[0], h.body_velocity[1], h.angular_velocity,
                     h.goals[0][0], h.goals[0][1], h.desired_force[0], h.desired_force[1]], np.float64)


def force_row(h):
    return np.array([*h.desired_force, *h.obstacle_force, *h.social_force, h.torque_force, *h.global_force], np.float64)


def robot_row(sim):
    r = sim.robot
    if not getattr(sim, "insert_robot", False) or len(r.goals) == 0:
        return np.full(13, np.nan)
    return r.get_safe_state()


def run_traj(sim, name, n_steps, save_every, robot_vel=(0.0, 0.0), dense_first=20, dt=DT):
    mm = sim.motion_model_manager
    humans = sim.humans
    model = mm.motion_model_title
    rv = np.array(robot_vel, np.float64)
    case = dict(
        type=np.int64(SFMS.index(model)),
        states0=np.array([h.get_safe_state() for h in humans]),
        robot0=robot_row(sim),
        robot_vel=rv,
        goals0=pack_goals(humans),
        walls=pack_walls(mm.walls),
        params=np.array([h.get_parameters(model) for h in humans]),
        safety=np.array([h.safety_space for h in humans] + [sim.robot.safety_space], np.float64),
        flags=np.array([int(mm.consider_robot), int(mm.all_equal_humans), n_steps, save_every], np.int64),
        dt=np.float64(dt),
    )
    steps, traj, rtraj = [0], [np.array([human_row(h) for h in humans])], [sim.robot.position.copy()]
    forces1 = None
    for s in range(1, n_steps + 1):
        if getattr(sim, "insert_robot", False):
            # what RobotAgent.step does for a holonomic action (robot_agent.py:126-131)
            sim.robot.position = sim.robot.position + rv * dt
            sim.robot.linear_velocity = rv.copy()
        mm.update_humans(0.0, dt)
        if s == 1:
            forces1 = np.array([force_row(h) for h in humans])
        if s <= dense_first or s % save_every == 0 or s == n_steps:
            steps.append(s)
            traj.append(np.array([human_row(h) for h in humans]))
            rtraj.append(sim.robot.position.copy())
    case.update(steps=np.array(steps, np.int64), traj=np.array(traj), robot_traj=np.array(rtraj), forces1=forces1)
    path = os.path.join(HERE, f"traj_{name}.npz")
    np.savez_compressed(path, **case)
    print(f"{name}: N={len(humans)} type={int(case['type'])} robot={int(mm.consider_robot)} equal={int(mm.all_equal_humans)} "
          f"walls={case['walls'].shape} steps={n_steps} -> {os.path.getsize(path)} B")


def cc_sim(model, seed, n, robot_visible, randomize_attr=False, parallel=False):
    np.random.seed(seed)
    sim = SocialNavSim({"insert_robot": True, "human_policy": model, "headless": True, "runge_kutta": False,
                        "robot_visible": robot_visible, "robot_radius": 0.3, "circle_radius": 7, "n_actors": n,
                        "randomize_human_positions": True, "randomize_human_attributes": randomize_attr},
                       scenario="circular_crossing", parallelize_humans=parallel)
    sim.set_time_step(DT)
    return sim


def ccso_sim(model, seed, n, robot_visible):
    np.random.seed(seed)
    sim = SocialNavSim({"insert_robot": True, "human_policy": model, "headless": True, "runge_kutta": False,
                        "robot_visible": robot_visible, "robot_radius": 0.3, "circle_radius": 7, "n_actors": n,
                        "randomize_human_positions": True},
                       scenario="circular_crossing_with_static_obstacles", parallelize_humans=False)
    sim.set_time_step(DT)
    return sim


def custom_sim(data, model, robot_visible=None, parallel=False):
    d = copy.deepcopy(data)
    d["headless"] = True
    d["motion_model"] = model
    if robot_visible is not None:
        d["robot_visible"] = robot_visible
    sim = SocialNavSim(d, scenario="custom_config", parallelize_humans=parallel)
    sim.set_time_step(DT)
    return sim


def dense_example_data():
    """config_example.py walls + a denser crowd squeezed between them, so wall and body-contact terms fire."""
    d = copy.deepcopy(config_example.data)
    d["humans"] = {
        0: {"pos": [-0.84, -0.84], "yaw": -np.pi, "goals": [[-2.5, -2.5], [0.5, -5.5]]},
        1: {"pos": [-5.0, -5.0], "yaw": 0.0, "goals": [[-2.5, -2.5], [0.5, -5.5]]},
        2: {"pos": [-5.5, -2.5], "yaw": -np.pi, "goals": [[-4.5, -0.5], [0.5, 0.5]], "radius": 0.35},
        3: {"pos": [-5.5, -4.17], "yaw": 0.0, "goals": [[-4.5, -0.5], [0.5, 0.5]], "radius": 0.4},
        4: {"pos": [-3.0, 1.0], "yaw": 0.3, "goals": [[1.0, -1.3], [-3.0, 1.0]], "des_speed": 1.2},
        5: {"pos": [2.0, -1.2], "yaw": 2.0, "goals": [[-4.0, 2.6], [2.0, -1.2]], "mass": 60},
        6: {"pos": [-6.4, -5.2], "yaw": 1.0, "goals": [[-6.4, -3.9], [-6.4, -5.2]]},
    }
    d["robot"] = {"pos": [-2.0, -2.0], "yaw": 0.0, "radius": 0.25, "goals": [[-7.5, -7.5]]}
    return d


def walls7eq_cases():
    """Same crowd with the last pair identical, so all_equal_humans is True (motion_model_manager.py:279-283
    only keeps the last pair's verdict) and the symmetric path runs with unequal radii and walls."""
    for model in ["sfm_guo", "hsfm_farina", "hsfm_new_moussaid", "hsfm_new_guo"]:
        d = dense_example_data()
        del d["humans"][5]["mass"]
        sim = custom_sim(d, model, robot_visible=True)
        run_traj(sim, f"walls7eq_{model}", 320, 40, robot_vel=(-0.3, -0.2))


def traj_cases():
    for model in SFMS:
        sim = cc_sim(model, 1002, 5, robot_visible=False)
        run_traj(sim, f"cc5_{model}", 400, 40, robot_vel=(0.0, 1.0))
    for model in SFMS:
        sim = cc_sim(model, 2003, 6, robot_visible=True)
        run_traj(sim, f"cc6_robot_{model}", 240, 40, robot_vel=(0.0, 1.0))
    for model in SFMS:
        sim = custom_sim(dense_example_data(), model, robot_visible=True)
        run_traj(sim, f"walls7_{model}", 320, 40, robot_vel=(-0.3, -0.2))
    walls7eq_cases()
    for model in ["sfm_guo", "hsfm_farina", "hsfm_new_moussaid"]:
        sim = custom_sim(config_example.data, model)
        run_traj(sim, f"example_{model}", 200, 40)
    for model in ["sfm_guo", "hsfm_new"]:
        sim = custom_sim(config_corridor.data, model)
        run_traj(sim, f"corridor_{model}", 1600, 100, dense_first=5)
    for model in ["hsfm_new_guo", "sfm_helbing", "hsfm_moussaid"]:
        sim = custom_sim(config_socialjym_cc.data, model)
        run_traj(sim, f"jym_{model}", 1600, 100, robot_vel=(0.0, 0.5), dense_first=5)
    for model in ["sfm_helbing", "hsfm_farina", "hsfm_new_guo", "sfm_moussaid"]:
        sim = ccso_sim(model, 7, 8, robot_visible=True)
        run_traj(sim, f"ccso8_{model}", 400, 40, robot_vel=(0.0, 1.0))
    for model in ["sfm_helbing", "sfm_guo", "hsfm_moussaid", "hsfm_new_guo"]:
        sim = cc_sim(model, 11, 7, robot_visible=True, randomize_attr=True)
        run_traj(sim, f"cc7_randattr_{model}", 240, 40, robot_vel=(0.1, 0.9))
    # a 25-human crowd (the headline shape), short
    for model in ["hsfm_farina", "hsfm_new_guo"]:
        sim = cc_sim(model, 2000, 25, robot_visible=True)
        run_traj(sim, f"cc25_robot_{model}", 60, 20, robot_vel=(0.0, 1.0), dense_first=5)


def pt_sim(model, seed, n, robot_visible):
    np.random.seed(seed)
    sim = SocialNavSim({"insert_robot": True, "human_policy": model, "headless": True, "runge_kutta": False, "robot_visible": robot_visible,
                        "robot_radius": 0.3, "traffic_length": 14, "traffic_height": 3, "n_actors": n, "randomize_human_attributes": False},
                       scenario="parallel_traffic", parallelize_humans=False)
    sim.set_time_step(DT)
    return sim


def pt_cases():
    """Parallel-traffic scenario with the respawn of humans that reach the left end (motion_model_manager.py:407-422,
    social_nav_sim.py:301-362).  The fixtures carry `respawn` = respawn_bounds; the goal columns change at every respawn."""
    for model, seed, n, vis in [("hsfm_farina", 2004, 5, True), ("sfm_helbing", 2011, 7, False), ("hsfm_new_guo", 1003, 5, True),
                                ("sfm_guo", 77, 10, True)]:
        sim = pt_sim(model, seed, n, vis)
        assert sim.motion_model_manager.parallel_traffic_humans_respawn
        name = f"pt{n}_{'robot_' if vis else ''}{model}"
        run_traj(sim, name, 1600, 50, robot_vel=(0.5, 0.0), dense_first=5)
        path = os.path.join(HERE, f"traj_{name}.npz")
        d = dict(np.load(path))
        d["respawn"] = np.array(sim.motion_model_manager.respawn_bounds, np.float64)
        np.savez_compressed(path, **d)
        gy = d["traj"][:, :, 9]
        print("   respawns seen:", int((np.diff(gy, axis=0) != 0).sum()))


def _near_goal(sim):
    sim.robot.position = np.array([0.3, 6.0])  # one metre from its goal (0, 7): the robot's goal list rotates within the run
    return sim


def robot_model_cases():
    """Robot driven by a human motion model, as SocialNavGym.imitation_learning_step does per sub-step (social_nav_gym.py:260-265):
    update_robot (motion_model_manager.py:593-653) then update_humans.  Fixture il_robot.npz."""
    out = {}
    confs = [("cc5_hsfm_farina__hsfm_farina", lambda: cc_sim("hsfm_farina", 1002, 5, True), "hsfm_farina", 600),
             ("cc5_sfm_helbing__hsfm_new_guo", lambda: cc_sim("sfm_helbing", 2003, 5, True), "hsfm_new_guo", 600),
             ("cc6_hsfm_new_guo__sfm_guo_invisible", lambda: cc_sim("hsfm_new_guo", 2003, 6, False), "sfm_guo", 400),
             ("walls7_hsfm_farina__sfm_moussaid", lambda: custom_sim(dense_example_data(), "hsfm_farina", True), "sfm_moussaid", 300),
             ("cc25_hsfm_farina__hsfm_farina", lambda: cc_sim("hsfm_farina", 2000, 25, True), "hsfm_farina", 100),
             ("cc5_near_goal_hsfm_guo__hsfm_guo", lambda: _near_goal(cc_sim("hsfm_guo", 31, 5, True)), "hsfm_guo", 400)]
    for key, mk, robot_model, n_steps in confs:
        sim = mk()
        mm = sim.motion_model_manager
        if len(sim.robot.goals) == 1:
            sim.robot.goals = [list(sim.robot.goals[0]), [float(sim.robot.position[0]), float(sim.robot.position[1])]]
        mm.set_robot_motion_model(robot_model, False)
        humans, robot = sim.humans, sim.robot
        model = mm.motion_model_title
        out[key + "_type"] = np.int64(SFMS.index(model))
        out[key + "_robot_type"] = np.int64(SFMS.index(robot_model))
        out[key + "_states0"] = np.array([h.get_safe_state() for h in humans])
        out[key + "_goals0"] = pack_goals(humans)
        out[key + "_walls"] = pack_walls(mm.walls)
        out[key + "_params"] = np.array([h.get_parameters(model) for h in humans])
        out[key + "_robot_params"] = robot.get_parameters(robot_model)
        out[key + "_robot0"] = robot.get_safe_state()
        out[key + "_robot_goals"] = np.array(robot.goals, np.float64)
        out[key + "_flags"] = np.array([int(mm.consider_robot), int(mm.all_equal_humans)], np.int64)
        steps, traj, rtraj = [0], [np.array([human_row(h) for h in humans])], [human_row(robot)]
        for s in range(1, n_steps + 1):
            mm.update_robot(0.0, DT)
            mm.update_humans(0.0, DT)
            if s <= 10 or s % 20 == 0:
                steps.append(s)
                traj.append(np.array([human_row(h) for h in humans]))
                rtraj.append(human_row(robot))
        out[key + "_steps"] = np.array(steps, np.int64)
        out[key + "_traj"] = np.array(traj)
        out[key + "_robot_traj"] = np.array(rtraj)
        rt = np.array(rtraj)
        print("il", key, "robot goal switches:", int((np.abs(np.diff(rt[:, 8:10], axis=0)).sum(1) > 0).sum()), "final robot pos", rt[-1, :2])
    path = os.path.join(HERE, "il_robot.npz")
    np.savez_compressed(path, **out)
    print("il_robot ->", os.path.getsize(path), "B")


def sim_update_cases():
    """Robot driven by a human motion model inside SocialNavSim.update (social_nav_sim.py:476-492 -> control_robot :500-529), the
    loop behind run_k_steps: every update the robot pose advances with its last velocity (update_robot_pose, mmm:655), every
    ROBOT_SAMPLING_TIME its velocities are refreshed by update_robot(..., just_velocities=True) with dt = ROBOT_SAMPLING_TIME
    (mmm:615; euler_*_single_agent_update mmm:72-85), or -- equal sampling times -- update_robot(dt) moves it; the humans are then
    updated seeing the robot's PREVIOUS state (sim:484-491).  Produced by calling sim.update() itself.  Fixture sim_update.npz."""
    out = {}
    confs = [("cc5_hsfm_farina__hsfm_new_guo_rt20", lambda: cc_sim("hsfm_farina", 1002, 5, True), "hsfm_new_guo", 0.25, 300),
             ("cc6_sfm_helbing__sfm_guo_invisible_rt4", lambda: cc_sim("sfm_helbing", 2003, 6, False), "sfm_guo", 0.05, 300),
             ("walls7_hsfm_farina__hsfm_farina_rt1", lambda: custom_sim(dense_example_data(), "hsfm_farina", True), "hsfm_farina", DT, 200),
             ("cc5_near_goal_hsfm_guo__hsfm_guo_rt20", lambda: _near_goal(cc_sim("hsfm_guo", 31, 5, True)), "hsfm_guo", 0.25, 400)]
    for key, mk, robot_model, robot_dt, n_steps in confs:
        sim = mk()
        mm = sim.motion_model_manager
        if len(sim.robot.goals) == 1:
            sim.robot.goals = [list(sim.robot.goals[0]), [float(sim.robot.position[0]), float(sim.robot.position[1])]]
        sim.set_time_step(DT)
        sim.set_robot_time_step(robot_dt)
        sim.set_robot_policy(policy_name=robot_model, runge_kutta=False)
        assert sim.robot_env_same_timestep == (robot_dt == DT)
        humans, robot = sim.humans, sim.robot
        model = mm.motion_model_title
        out[key + "_type"] = np.int64(SFMS.index(model))
        out[key + "_robot_type"] = np.int64(SFMS.index(robot_model))
        out[key + "_robot_dt"] = np.float64(robot_dt)
        out[key + "_every"] = np.int64(round(robot_dt / DT))
        out[key + "_states0"] = np.array([h.get_safe_state() for h in humans])
        out[key + "_goals0"] = pack_goals(humans)
        out[key + "_walls"] = pack_walls(mm.walls)
        out[key + "_params"] = np.array([h.get_parameters(model) for h in humans])
        out[key + "_robot_params"] = robot.get_parameters(robot_model)
        out[key + "_robot0"] = robot.get_safe_state()
        out[key + "_robot_goals"] = np.array(robot.goals, np.float64)
        out[key + "_flags"] = np.array([int(mm.consider_robot), int(mm.all_equal_humans)], np.int64)
        steps, traj, rtraj = [0], [np.array([human_row(h) for h in humans])], [human_row(robot)]
        for s in range(1, n_steps + 1):
            sim.update()
            if s <= 25 or s % 10 == 0:
                steps.append(s)
                traj.append(np.array([human_row(h) for h in humans]))
                rtraj.append(human_row(robot))
        out[key + "_steps"] = np.array(steps, np.int64)
        out[key + "_traj"] = np.array(traj)
        out[key + "_robot_traj"] = np.array(rtraj)
        rt = np.array(rtraj)
        print("sim_update", key, "robot goal switches:", int((np.abs(np.diff(rt[:, 8:10], axis=0)).sum(1) > 0).sum()), "final robot pos", rt[-1, :2],
              "yaw", rt[-1, 2])
    path = os.path.join(HERE, "sim_update.npz")
    np.savez_compressed(path, **out)
    print("sim_update ->", os.path.getsize(path), "B")


def lookahead_case():
    """The policy-side operator the CrowdNav value-network policies call once per decision (crowd_nav/policy/cadrl.py:42-83
    compute_rotated_states_and_reward, used by CADRL.predict :235-276): peek of the humans at dt = 0.25
    (get_next_human_observable_states), then rewards and agent-centric rotated states for all 81 actions."""
    from crowd_nav.policy.cadrl import compute_rotated_states_and_reward
    speeds = [(np.exp((i + 1) / 5) - 1) / (np.e - 1) * 1.0 for i in range(5)]            # cadrl.py build_action_space (holonomic)
    rotations = np.linspace(0, 2 * np.pi, 16, endpoint=False)
    actions = np.array([[0.0, 0.0]] + [[sp * np.cos(r), sp * np.sin(r)] for r in rotations for sp in speeds])
    out = {"actions": actions}
    for model, seed, n, steps in [("hsfm_farina", 1002, 5, 330), ("sfm_helbing", 2003, 6, 200), ("hsfm_new_guo", 2000, 25, 120)]:
        sim = cc_sim(model, seed, n, robot_visible=False)
        mm = sim.motion_model_manager
        for k in range(steps):
            sim.robot.position = sim.robot.position + np.array([0.0, 1.0]) * DT
            mm.update_humans(0.0, DT)
        rng = np.random.RandomState(seed)
        for rep in range(3):
            h0 = sim.humans[rng.randint(n)]
            sim.robot.position = h0.position + rng.uniform(-1.2, 1.2, 2) if rep < 2 else np.array(sim.robot.goals[0], float) - [0.1, 0.2]
            sim.robot.linear_velocity = rng.uniform(-1, 1, 2)
            r = sim.robot
            robot_state = np.array([r.position[0], r.position[1], r.linear_velocity[0], r.linear_velocity[1], r.radius, r.goals[0][0], r.goals[0][1],
                                    r.desired_speed, r.yaw])
            for vis in (False, True):
                if vis:
                    cur = np.array([[h.position[0], h.position[1], h.linear_velocity[0], h.linear_velocity[1], h.radius, h.yaw, h.angular_velocity]
                                    for h in sim.humans])
                    nxt = mm.get_next_human_observable_states(0.25, theta_and_omega_visible=True)[:, :6]
                else:
                    cur = np.array([[h.position[0], h.position[1], h.linear_velocity[0], h.linear_velocity[1], h.radius] for h in sim.humans])
                    nxt = mm.get_next_human_observable_states(0.25)
                rot, rew = compute_rotated_states_and_reward(actions, nxt, cur, robot_state, 0.25, theta_and_omega_visible=vis)
                key = f"{model}_{rep}_{int(vis)}"
                out[key + "_cur"], out[key + "_next"], out[key + "_robot"] = cur, nxt, robot_state
                out[key + "_rotated"], out[key + "_rewards"] = rot, rew
            out[f"{model}_{rep}_states"] = np.array([h.get_safe_state() for h in sim.humans])
            out[f"{model}_{rep}_desired"] = np.array([h.desired_force for h in sim.humans])
            out[f"{model}_{rep}_goals"] = pack_goals(sim.humans)
            rw = out[f"{model}_{rep}_0_rewards"]
            print("lookahead", model, rep, "collisions", int((rw == -0.25).sum()), "goal", int((rw == 1).sum()), "discomfort", int(((rw < 0) & (rw > -0.25)).sum()))
    path = os.path.join(HERE, "lookahead.npz")
    np.savez_compressed(path, **out)
    print("lookahead ->", os.path.getsize(path), "B")


def scenario_cases():
    """The reference's scenario generators (social_nav_sim.py:200-431) on many seeds, through SocialNavSim's own entry point, with the
    number of uniforms each one consumed (read back from the global MT19937 position) and the hybrid scenario's coin
    (social_nav_gym.py:155-156 np.random.choice).  Pins scenarios.py and the on-device reset (snp_reset)."""
    out = {}

    def rows(sim):
        st = np.array([h.get_safe_state() for h in sim.humans])
        return st, pack_goals(sim.humans)

    counter = [0]
    real_random, real_uniform = np.random.random, np.random.uniform

    def counting_random(*a, **k):
        counter[0] += 1
        return real_random(*a, **k)

    def counting_uniform(*a, **k):
        counter[0] += 1
        return real_uniform(*a, **k)

    np.random.random, np.random.uniform = counting_random, counting_uniform   # the generators draw through these two only

    cases = [("cc", 5, False), ("cc", 25, False), ("cc", 7, True), ("pt", 5, False), ("pt", 12, True), ("ccso", 6, False), ("ccso", 8, False)]
    for kind, n, rand in cases:
        seeds = np.arange(3000, 3000 + (6 if n == 25 else 16))
        S, G, D = [], [], []
        for seed in seeds:
            np.random.seed(int(seed))
            counter[0] = 0
            base = {"insert_robot": True, "human_policy": "hsfm_farina", "headless": True, "runge_kutta": False, "robot_visible": False,
                    "robot_radius": 0.3, "n_actors": n}
            if kind == "cc":
                sim = SocialNavSim({**base, "circle_radius": 7, "randomize_human_positions": True, "randomize_human_attributes": rand},
                                   scenario="circular_crossing", parallelize_humans=False)
            elif kind == "pt":
                sim = SocialNavSim({**base, "traffic_length": 14, "traffic_height": 3, "randomize_human_attributes": rand},
                                   scenario="parallel_traffic", parallelize_humans=False)
            else:
                sim = SocialNavSim({**base, "circle_radius": 7, "randomize_human_positions": True},
                                   scenario="circular_crossing_with_static_obstacles", parallelize_humans=False)
            D.append(counter[0])
            st, gl = rows(sim)
            S.append(st); G.append(gl)
        key = f"{kind}{n}{'_randattr' if rand else ''}"
        out[key + "_seeds"], out[key + "_states"], out[key + "_goals"], out[key + "_draws"] = seeds, np.stack(S), np.stack(G), np.array(D)
        print("scenario", key, "draws", D[:6])
    np.random.random, np.random.uniform = real_random, real_uniform
    coins = []
    for seed in range(3000, 3064):
        np.random.seed(seed)
        coins.append(["circle_crossing", "parallel_traffic"].index(np.random.choice(["circle_crossing", "parallel_traffic"])))
    out["hybrid_seeds"], out["hybrid_choice"] = np.arange(3000, 3064), np.array(coins)
    path = os.path.join(HERE, "scenarios.npz")
    np.savez_compressed(path, **out)
    print("scenarios ->", os.path.getsize(path), "B")


def push_out_case():
    """RobotAgent.check_collisions (robot_agent.py:35-48): the robot pushed out of the humans and walls it overlaps, through the
    reference's own method on the reference's own objects (7 humans + 3 wall polygons of the dense example, and a wall-free
    circular crossing), from 48 + 16 start positions placed on / near humans and wall edges."""
    out = {}
    rng = np.random.RandomState(5)
    for name, sim in [("walls", custom_sim(dense_example_data(), "hsfm_farina", robot_visible=True)), ("cc", cc_sim("sfm_helbing", 2003, 6, robot_visible=True))]:
        humans = np.array([[h.position[0], h.position[1], h.radius] for h in sim.humans])
        walls = list(sim.walls)
        starts, ends = [], []
        n_cases = 48 if name == "walls" else 16
        for k in range(n_cases):
            if k % 3 == 0 and walls:
                w = walls[rng.randint(len(walls))]
                seg = list(w.segments.values())[rng.randint(len(w.segments))]
                t = rng.uniform(0, 1)
                p = np.array(seg[0]) * (1 - t) + np.array(seg[1]) * t + rng.uniform(-0.35, 0.35, 2)
            elif k % 3 == 1:
                h = sim.humans[rng.randint(len(sim.humans))]
                p = h.position + rng.uniform(-0.5, 0.5, 2)
            else:
                i, j = rng.choice(len(sim.humans), 2, replace=False)
                p = 0.5 * (sim.humans[i].position + sim.humans[j].position) + rng.uniform(-0.3, 0.3, 2)
            sim.robot.position = np.array(p, dtype=np.float64)
            starts.append(sim.robot.position.copy())
            sim.robot.check_collisions(sim.humans, sim.walls)
            ends.append(np.array(sim.robot.position, dtype=np.float64))
        out[name + "_humans"], out[name + "_walls"] = humans, pack_walls(sim.walls)
        out[name + "_radius"] = np.float64(sim.robot.radius)
        out[name + "_start"], out[name + "_end"] = np.array(starts), np.array(ends)
        print("push_out", name, "moved", int((np.abs(np.array(starts) - np.array(ends)).sum(1) > 0).sum(), ), "of", n_cases)
    path = os.path.join(HERE, "push_out.npz")
    np.savez_compressed(path, **out)
    print("push_out ->", os.path.getsize(path), "B")


def numba_cases():
    """Second witness: the reference's Numba operator update_humans_parallel (forces_parallel.py:184)."""
    out = {}
    for tag, mk in [("cc6_robot", lambda m: cc_sim(m, 2003, 6, True, parallel=True)),
                    ("walls7", lambda m: custom_sim(dense_example_data(), m, True, parallel=True))]:
        for model in SFMS:
            sim = mk(model)
            mm = sim.motion_model_manager
            states = mm.states.copy()
            goals = mm.goals.copy()
            obstacles = None if mm.obstacles is None else mm.obstacles.copy()
            key = f"{tag}_{model}"
            out[key + "_states0"] = states.copy()
            out[key + "_goals0"] = goals.copy()
            out[key + "_walls"] = np.zeros((0, 1, 2, 2)) if obstacles is None else obstacles.copy()
            out[key + "_params"] = mm.params.copy()
            out[key + "_safety"] = mm.safety_space.copy()
            out[key + "_flags"] = np.array([mm.sfm_type, int(mm.all_equal_humans), int(mm.consider_robot)], np.int64)
            seq = []
            for _ in range(3):
                states = update_humans_parallel(mm.sfm_type, states, goals, obstacles, mm.params, DT, mm.safety_space,
                                                all_params_equal=mm.all_equal_humans, last_is_robot=mm.consider_robot)
                seq.append(states.copy())
            out[key + "_out"] = np.array(seq)
    path = os.path.join(HERE, "numba_operator.npz")
    np.savez_compressed(path, **out)
    print("numba_operator ->", os.path.getsize(path), "B")


def peek_case():
    """get_next_human_observable_states (motion_model_manager.py:691): peek at dt=0.25 and restore."""
    out = {}
    for model in ["sfm_helbing", "hsfm_farina", "hsfm_new_guo"]:
        sim = cc_sim(model, 1002, 5, robot_visible=True)
        mm = sim.motion_model_manager
        for _ in range(160):
            mm.update_humans(0.0, DT)
        before = np.array([human_row(h) for h in sim.humans])
        out[f"{model}_type"] = np.int64(SFMS.index(model))
        out[f"{model}_states"] = np.array([h.get_safe_state() for h in sim.humans])
        out[f"{model}_desired"] = before[:, 10:12]
        out[f"{model}_goals"] = pack_goals(sim.humans)
        out[f"{model}_robot"] = sim.robot.get_safe_state()
        out[f"{model}_params"] = np.array([h.get_parameters(model) for h in sim.humans])
        out[f"{model}_obs4"] = mm.get_next_human_observable_states(0.25)
        out[f"{model}_obs8"] = mm.get_next_human_observable_states(0.25, theta_and_omega_visible=True)
        after = np.array([human_row(h) for h in sim.humans])
        out[f"{model}_before"] = before
        out[f"{model}_after"] = after
    path = os.path.join(HERE, "peek.npz")
    np.savez_compressed(path, **out)
    print("peek ->", os.path.getsize(path), "B")


def flags_case():
    """collision_detection_and_reaching_goal (social_nav_sim.py:949), compute_reward_and_infos (:986),
    check_actual_collisions_and_goal (social_nav_gym.py:107), run_k_steps collision (social_nav_sim.py:702)."""
    rng = np.random.RandomState(5)
    sim = cc_sim("hsfm_farina", 1002, 5, robot_visible=False)
    sim.time_limit = 50
    sim.collision_penalty = -0.25
    sim.success_reward = 1.0
    sim.discomfort_dist = 0.2
    sim.discomfort_penalty_factor = 0.5
    mm = sim.motion_model_manager
    rows = []
    hum, rob, act, res = [], [], [], []
    info_code = {"Timeout": 1, "Collision": 2, "Reaching goal": 3, "Too close": 4, "": 0}
    for k in range(400):
        if k % 4 == 0:
            for _ in range(8):
                mm.update_humans(0.0, DT)
        h0 = sim.humans[rng.randint(5)]
        mode = k % 5
        if mode == 0:
            sim.robot.position = h0.position + rng.uniform(-1.5, 1.5, 2)
        elif mode == 1:
            ang = rng.uniform(0, 2 * np.pi)
            sim.robot.position = h0.position + (0.6 + rng.uniform(-0.02, 0.3)) * np.array([np.cos(ang), np.sin(ang)])
        elif mode == 2:
            sim.robot.position = np.array(sim.robot.goals[0], np.float64) + rng.uniform(-0.5, 0.5, 2)
        elif mode == 3:
            sim.robot.position = rng.uniform(-7, 7, 2)
        else:
            sim.robot.position = h0.position + rng.uniform(-0.7, 0.7, 2)
        a = rng.uniform(-1.0, 1.0, 2)
        t_now = 49.5 if k % 37 == 0 else rng.uniform(0, 40)
        col, dmin, goal = sim.collision_detection_and_reaching_goal(a, 0.25)
        reward, term, trunc, info = sim.compute_reward_and_infos(col, dmin, goal, t_now, 0.25)
        acol, admin, agoal = gym_mod.SocialNavGym.check_actual_collisions_and_goal(sim)
        kcol = any(np.linalg.norm(h.position - sim.robot.position) < (h.radius + sim.robot.radius) for h in sim.humans)
        hum.append(np.array([h.get_safe_state() for h in sim.humans]))
        rob.append(sim.robot.get_safe_state())
        act.append(a)
        res.append([float(col), dmin, float(goal), reward, float(term), float(trunc), info_code[str(info)],
                    float(acol), admin, float(agoal), float(kcol), t_now])
    path = os.path.join(HERE, "flags.npz")
    np.savez_compressed(path, humans=np.array(hum), robot=np.array(rob), action=np.array(act), result=np.array(res),
                        consts=np.array([50, -0.25, 1.0, 0.2, 0.5, 0.25]))
    r = np.array(res)
    print("flags ->", os.path.getsize(path), "B", "collisions", int(r[:, 0].sum()), "goals", int(r[:, 2].sum()),
          "danger", int((r[:, 6] == 4).sum()), "actual col", int(r[:, 7].sum()))


def laser_hits(sensor, humans, walls):
    """Replay LaserSensor.get_laser_measurements' loop (sensors.py:53-69) with the sensor's own
    intersect functions to record which entity produced each minimum (first strict '<' winner)."""
    angles = np.linspace(sensor.yaw - (sensor.range / 2), sensor.yaw + (sensor.range / 2), sensor.samples)
    hits = []
    for angle in angles:
        m, hit = sensor.max_distance, -1
        d = np.array([math.cos(angle), math.sin(angle)], dtype=np.float64)  # sensors.py:58
        for i, h in enumerate(humans):
            rc = sensor.sphere_ray_intersect(d, h.position, h.radius)
            if rc < m:
                m, hit = rc, i
        k = len(humans)
        for w in walls:
            for seg in w.segments.values():
                rc = sensor.segment_ray_intersect(d, seg)
                if rc < m:
                    m, hit = rc, k
                k += 1
        hits.append(hit)
    return np.array(hits, np.int64)


def laser_case():
    out = {}
    confs = [("dense", lambda: custom_sim(dense_example_data(), "hsfm_farina", True), 360, 2 * np.pi, 10.0, np.pi / 2),
             ("dense_narrow", lambda: custom_sim(dense_example_data(), "hsfm_farina", True), 61, np.pi, 6.0, -2.5),
             ("example", lambda: custom_sim(config_example.data, "sfm_guo"), 360, 2 * np.pi, 10.0, 0.0),
             ("cc25", lambda: cc_sim("hsfm_farina", 2000, 25, True), 360, 2 * np.pi, 10.0, np.pi / 2),
             ("corridor", lambda: custom_sim(config_corridor.data, "sfm_guo"), 180, 2 * np.pi, 10.0, 1.0)]
    for name, mk, samples, rng_, maxd, yaw in confs:
        sim = mk()
        mm = sim.motion_model_manager
        for rep in range(3):
            if rep:
                for _ in range(150):
                    mm.update_humans(0.0, DT)
            if name == "corridor":
                pos = np.array([[0.0, 3.0], [-1.5, 0.2], [0.45, 0.5]][rep])
            elif name.startswith("cc25"):
                pos = np.array([[0.0, -7.0], [0.0, -3.0], [1.0, 2.0]][rep])
            else:
                pos = np.array([[-2.0, -2.0], [-4.4, -3.0], [0.0, 0.0]][rep])
            sensor = LaserSensor(pos, yaw, rng_, samples, maxd, uncertainty=None)
            sensor.uncertainty = None  # deterministic: sensors.py:67 only adds noise when not None
            meas = sensor.get_laser_measurements(sim.humans, mm.walls)
            key = f"{name}_{rep}"
            out[key + "_humans"] = np.array([[h.position[0], h.position[1], h.radius] for h in sim.humans])
            out[key + "_walls"] = pack_walls(mm.walls)
            out[key + "_pose"] = np.array([pos[0], pos[1], yaw, rng_, samples, maxd])
            out[key + "_angles"] = np.array(list(meas.keys()))
            out[key + "_ranges"] = np.array(list(meas.values()))
            out[key + "_hits"] = laser_hits(sensor, sim.humans, mm.walls)
    path = os.path.join(HERE, "laser.npz")
    np.savez_compressed(path, **out)
    nh = sum(int((out[k] >= 0).sum()) for k in out if k.endswith("_hits"))
    print("laser ->", os.path.getsize(path), "B", "rays with a hit:", nh)


def gym_case():
    """SocialNavGym.reset/step (social_nav_gym.py:120,227) with a blind-planner robot, serial human path."""
    gym_mod.PARALLELIZE_HUMANS = False
    out = {}
    for model, visible in [("hsfm_farina", False), ("sfm_helbing", True), ("hsfm_new_guo", True)]:
        cfg = configparser.RawConfigParser()
        cfg.read(os.path.join(ref_shim.REFERENCE_ROOT, "crowd_nav/configs/env.config"))
        cfg.set("humans", "policy", model)
        cfg.set("sim", "train_val_sim", "circle_crossing")
        cfg.set("sim", "test_sim", "circle_crossing")
        cfg.set("robot", "policy", "bp")
        cfg.set("robot", "visible", "true" if visible else "false")
        env = gym_mod.SocialNavGym()
        env.configure(cfg)
        robot = RobotAgent(env)
        robot.configure(cfg, "robot")
        env.set_robot(robot)
        robot.policy.with_theta_and_omega_visible = False
        robot.policy.time_step = 0.25
        ob, _ = env.reset(phase="test", test_case=3)
        mm = env.motion_model_manager
        key = f"{model}_{int(visible)}"
        out[key + "_type"] = np.int64(SFMS.index(model))
        out[key + "_states0"] = np.array([h.get_safe_state() for h in env.humans])
        out[key + "_goals0"] = pack_goals(env.humans)
        out[key + "_robot0"] = robot.get_safe_state()
        out[key + "_params"] = np.array([h.get_parameters(model) for h in env.humans])
        acts, obs, res, rpos = [], [], [], []
        rng = np.random.RandomState(3)
        for k in range(60):
            a = robot.act(ob)
            if k % 3 == 1:
                a = ActionXY(a.vx + rng.uniform(-0.4, 0.4), a.vy + rng.uniform(-0.4, 0.4))
            ob, reward, term, trunc, info = env.step(a)
            acts.append([a.vx, a.vy])
            obs.append([[o.px, o.py, o.vx, o.vy, o.radius] for o in ob])
            code = {"Timeout": 1, "Collision": 2, "Reaching goal": 3, "Too close": 4, "": 0}[str(info[0])]
            res.append([reward, float(term), float(trunc), code])
            rpos.append(robot.position.copy())
        out[key + "_actions"] = np.array(acts)
        out[key + "_obs"] = np.array(obs)
        out[key + "_result"] = np.array(res)
        out[key + "_robot_pos"] = np.array(rpos)
        out[key + "_final"] = np.array([human_row(h) for h in env.humans])
        print("gym", key, "codes", sorted(set(int(r[3]) for r in res)))
    path = os.path.join(HERE, "gym_step.npz")
    np.savez_compressed(path, **out)
    print("gym_step ->", os.path.getsize(path), "B")


if __name__ == "__main__":
    which = sys.argv[1:] or ["traj", "pt", "il", "sim_update", "lookahead", "scenarios", "push_out", "numba", "peek", "flags", "laser", "gym"]
    if "traj" in which:
        traj_cases()
    if "pt" in which:
        pt_cases()
    if "il" in which:
        robot_model_cases()
    if "sim_update" in which:
        sim_update_cases()
    if "lookahead" in which:
        lookahead_case()
    if "scenarios" in which:
        scenario_cases()
    if "push_out" in which:
        push_out_case()
    if "numba" in which:
        numba_cases()
    if "peek" in which:
        peek_case()
    if "flags" in which:
        flags_case()
    if "laser" in which:
        laser_case()
    if "gym" in which:
        gym_case()
