"""ctypes front-end of the C oracle (oracle/snp_oracle.c).  Test infrastructure, not product code.

Array conventions are the reference's own (AoS float64): state rows of 13
(src/agent.py:256), params rows of 20 (src/agent.py:269), goals [N,G,2] and walls [W,S,2,2] NaN padded
(src/motion_model_manager.py:262-276), all with a leading env axis E.
"""
import ctypes
import os
from dataclasses import dataclass

import numpy as np

from . import build as _build

N_STATE = 13
N_PARAMS = 20
# src/motion_model_manager.py:15-17
SFMS = ["sfm_helbing", "sfm_guo", "sfm_moussaid", "hsfm_farina", "hsfm_guo", "hsfm_moussaid",
        "hsfm_new", "hsfm_new_guo", "hsfm_new_moussaid"]


def type_code(title: str) -> int:
    return SFMS.index(title)


def default_params(title: str) -> np.ndarray:
    """The 20-vector Agent.get_parameters builds for a model title (src/agent.py:94-243,268-388)."""
    t = type_code(title)
    p = np.zeros(N_PARAMS)
    p[0] = 0.5
    p[2], p[4] = 2000.0, 0.08
    p[10], p[11] = 120000.0, 240000.0
    soc = t % 3
    if soc in (0, 1):
        p[1], p[3] = 2000.0, 0.08
    if soc == 1:
        p[5], p[6], p[7], p[8] = 120.0, 120.0, 0.6, 0.6
    if soc == 2:
        p[9], p[12], p[13], p[14], p[15] = 360.0, 2.0, 0.35, 2.0, 3.0
    if t >= 3:
        p[16], p[17], p[18], p[19] = 1.0, 500.0, 3.0, 0.1
    return p


class _Cfg(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int) for k in ("type", "n", "g", "n_walls", "n_segs", "consider_robot", "symmetric",
                                             "numba_compat", "walls_per_env", "respawn")] + [("respawn_bounds", ctypes.c_double * 2)]


@dataclass
class OracleConfig:
    type: int
    consider_robot: bool = False
    symmetric: bool = True
    numba_compat: bool = False
    respawn_bounds: tuple = None  # (traffic_length/2, traffic_height/2): parallel-traffic respawn after every update (mmm:407-422)


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.orc_update_humans.argtypes = [ctypes.POINTER(_Cfg), ctypes.c_int, dp, dp, dp, dp, dp, dp, ctypes.c_double,
                                           ctypes.c_int, dp, dp, ctypes.c_int]
        _lib.orc_update_humans.restype = None
        _lib.orc_imitation_steps.argtypes = [ctypes.POINTER(_Cfg), ctypes.c_int, dp, dp, dp, dp, dp, dp, ctypes.c_double, ctypes.c_int, dp, dp,
                                             ctypes.c_int, dp, dp, ctypes.c_int, dp, ctypes.c_int]
        _lib.orc_imitation_steps.restype = None
        _lib.orc_sim_update_steps.argtypes = [ctypes.POINTER(_Cfg), ctypes.c_int, dp, dp, dp, dp, dp, dp, ctypes.c_double, ctypes.c_int, dp, dp,
                                              ctypes.c_int, dp, dp, ctypes.c_int, dp, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int]
        _lib.orc_sim_update_steps.restype = None
        _lib.orc_lookahead.argtypes = [ctypes.c_int] * 4 + [dp, dp, dp, dp, ctypes.c_double, dp, dp, ctypes.c_int]
        _lib.orc_lookahead.restype = None
        _lib.orc_robot_push_out.argtypes = [ctypes.c_int] * 4 + [dp, dp, dp]
        _lib.orc_robot_push_out.restype = None
        _lib.orc_checks.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, dp, dp]
        _lib.orc_checks.restype = None
        _lib.orc_laser.argtypes = [ctypes.c_int] * 5 + [dp, dp, dp, ctypes.c_double, ctypes.c_int, ctypes.c_double, dp,
                                                        ctypes.POINTER(ctypes.c_int64), ctypes.c_int]
        _lib.orc_laser.restype = None
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def update_humans(cfg: OracleConfig, states, goals, walls, params, safety, desired, dt, n_steps=1, robot_vel=None,
                  want_forces=False, n_threads=1):
    """Batched MotionModelManager.update_humans (serial Euler path).  All arrays carry a leading env axis:
    states [E,N(+1),13], goals [E,N,G,2], walls [W,S,2,2] or [E,W,S,2,2] or None, params [E,N,20],
    safety [E,N(+1)], desired [E,N,2].  Returns (states, goals, desired[, forces]) as NEW arrays."""
    lib = _load()
    states = _c(states).copy()
    goals = _c(goals).copy()
    desired = _c(desired).copy()
    params = _c(params)
    safety = _c(safety)
    E, rows, _ = states.shape
    n = goals.shape[1]
    assert rows == n + int(cfg.consider_robot), (rows, n, cfg.consider_robot)
    if walls is None or np.size(walls) == 0 or walls.shape[-4] == 0:
        W, S, per_env, walls_c = 0, 1, 0, np.zeros(4)
    else:
        walls_c = _c(walls)
        per_env = int(walls_c.ndim == 5)
        W, S = walls_c.shape[-4], walls_c.shape[-3]
    rb = cfg.respawn_bounds
    c = _Cfg(cfg.type, n, goals.shape[2], W, S, int(cfg.consider_robot), int(cfg.symmetric), int(cfg.numba_compat), per_env,
             int(rb is not None), (ctypes.c_double * 2)(*(rb if rb is not None else (0.0, 0.0))))
    forces = np.zeros((E, n, 9)) if want_forces else None
    rv = None if robot_vel is None else _c(robot_vel)
    lib.orc_update_humans(ctypes.byref(c), E, _dp(states), _dp(goals), _dp(walls_c), _dp(params), _dp(safety), _dp(desired),
                          float(dt), int(n_steps), _dp(rv), _dp(forces), int(n_threads))
    if want_forces:
        return states, goals, desired, forces
    return states, goals, desired


def imitation_steps(cfg: OracleConfig, states, goals, walls, params, safety, desired, dt, n_steps, robot, robot_goals, robot_desired,
                    robot_params, robot_type, robot_safety=None, n_threads=1):
    """n_steps x (update_robot; update_humans) -- the sub-step loop of SocialNavGym.imitation_learning_step (gym:260-265).
    robot [E,13], robot_goals [E,RG,2], robot_desired [E,2], robot_params [20].  Returns new (states, goals, desired, robot,
    robot_goals, robot_desired)."""
    lib = _load()
    states, goals, desired = _c(states).copy(), _c(goals).copy(), _c(desired).copy()
    robot, robot_goals, robot_desired = _c(robot).copy(), _c(robot_goals).copy(), _c(robot_desired).copy()
    params, safety = _c(params), _c(safety)
    E, rows, _ = states.shape
    n = goals.shape[1]
    assert rows == n + int(cfg.consider_robot)
    if walls is None or np.size(walls) == 0 or walls.shape[-4] == 0:
        W, S, per_env, walls_c = 0, 1, 0, np.zeros(4)
    else:
        walls_c = _c(walls)
        per_env = int(walls_c.ndim == 5)
        W, S = walls_c.shape[-4], walls_c.shape[-3]
    rb = cfg.respawn_bounds
    c = _Cfg(cfg.type, n, goals.shape[2], W, S, int(cfg.consider_robot), int(cfg.symmetric), int(cfg.numba_compat), per_env,
             int(rb is not None), (ctypes.c_double * 2)(*(rb if rb is not None else (0.0, 0.0))))
    rsaf = np.zeros(E) if robot_safety is None else _c(robot_safety)
    lib.orc_imitation_steps(ctypes.byref(c), E, _dp(states), _dp(goals), _dp(walls_c), _dp(params), _dp(safety), _dp(desired), float(dt),
                            int(n_steps), _dp(robot), _dp(robot_goals), robot_goals.shape[1], _dp(robot_desired), _dp(_c(robot_params)),
                            int(robot_type), _dp(rsaf), int(n_threads))
    return states, goals, desired, robot, robot_goals, robot_desired


def sim_update_steps(cfg: OracleConfig, states, goals, walls, params, safety, desired, dt, n_steps, robot, robot_goals, robot_desired,
                     robot_params, robot_type, every, robot_dt, phase=0, robot_safety=None, n_threads=1):
    """n_steps x SocialNavSim.update with a model-driven robot (social_nav_sim.py:476-529): pose advance every update, velocity
    refresh (update_robot, just_velocities=True, dt = robot_dt) every `every` updates -- or a full update_robot(dt) when
    every <= 1 -- and humans that see the robot's previous state.  Same arrays and return value as imitation_steps."""
    lib = _load()
    states, goals, desired = _c(states).copy(), _c(goals).copy(), _c(desired).copy()
    robot, robot_goals, robot_desired = _c(robot).copy(), _c(robot_goals).copy(), _c(robot_desired).copy()
    params, safety = _c(params), _c(safety)
    E, rows, _ = states.shape
    n = goals.shape[1]
    assert rows == n + int(cfg.consider_robot)
    if walls is None or np.size(walls) == 0 or walls.shape[-4] == 0:
        W, S, per_env, walls_c = 0, 1, 0, np.zeros(4)
    else:
        walls_c = _c(walls)
        per_env = int(walls_c.ndim == 5)
        W, S = walls_c.shape[-4], walls_c.shape[-3]
    rb = cfg.respawn_bounds
    c = _Cfg(cfg.type, n, goals.shape[2], W, S, int(cfg.consider_robot), int(cfg.symmetric), int(cfg.numba_compat), per_env,
             int(rb is not None), (ctypes.c_double * 2)(*(rb if rb is not None else (0.0, 0.0))))
    rsaf = np.zeros(E) if robot_safety is None else _c(robot_safety)
    lib.orc_sim_update_steps(ctypes.byref(c), E, _dp(states), _dp(goals), _dp(walls_c), _dp(params), _dp(safety), _dp(desired), ctypes.c_double(dt),
                             int(n_steps), _dp(robot), _dp(robot_goals), robot_goals.shape[1], _dp(robot_desired), _dp(_c(robot_params)),
                             int(robot_type), _dp(rsaf), int(every), ctypes.c_double(robot_dt), int(phase), int(n_threads))
    return states, goals, desired, robot, robot_goals, robot_desired


def lookahead(cur, nxt, robot, actions, dt, visible=False, n_threads=1):
    """compute_rotated_states_and_reward (crowd_nav/policy/cadrl.py:42-83) for E envs: cur [E,N,5|7], nxt [E,N,4|6], robot [E,9],
    actions [A,2].  Returns (rotated [E,A,N,13|15], rewards [E,A])."""
    lib = _load()
    cur, nxt, robot, actions = _c(cur), _c(nxt), _c(robot), _c(actions)
    E, N, _ = cur.shape
    A = actions.shape[0]
    rotated = np.zeros((E, A, N, 15 if visible else 13))
    rewards = np.zeros((E, A))
    lib.orc_lookahead(E, N, A, int(visible), _dp(cur), _dp(nxt), _dp(robot), _dp(actions), float(dt), _dp(rotated), _dp(rewards), int(n_threads))
    return rotated, rewards


def robot_push_out(humans, walls, robot):
    """RobotAgent.check_collisions (robot_agent.py:35-48): humans [E,n,3] = x,y,r; walls [W,S,2,2] NaN padded or None;
    robot [E,3] = x,y,r.  Returns the pushed-out robot positions [E,2]."""
    lib = _load()
    humans, rb = _c(humans), _c(robot).copy()
    E, n, _ = humans.shape
    if walls is None or np.asarray(walls).size == 0:
        W, S, wl = 0, 0, np.zeros(4)
    else:
        wl = _c(np.asarray(walls, np.float64).reshape(walls.shape[0], walls.shape[1], 4))
        W, S = wl.shape[0], wl.shape[1]
    lib.orc_robot_push_out(E, n, W, S, _dp(humans), _dp(wl), _dp(rb))
    return rb[:, :2]


def checks(states, n_humans, robot, action, time_now, consts):
    """states [E,rows,13] (first n_humans rows are humans), robot [E,13], action [E,2], time_now [E],
    consts = (time_limit, collision_penalty, success_reward, discomfort_dist, discomfort_penalty_factor, robot_time_step).
    Returns [E,12], see orc_checks."""
    lib = _load()
    states = _c(states)
    E, rows, _ = states.shape
    out = np.zeros((E, 12))
    lib.orc_checks(E, int(n_humans), rows, _dp(states), _dp(_c(robot)), _dp(_c(action)), _dp(_c(time_now)), _dp(_c(consts)), _dp(out))
    return out


def laser(humans, walls, pose, range_, samples, max_distance, n_threads=1):
    """humans [E,N,3] (x,y,r), walls [W,S,2,2] / [E,W,S,2,2] / None, pose [E,3] (x,y,yaw).
    Returns (ranges [E,samples] float64, hits [E,samples] int64)."""
    lib = _load()
    humans = _c(humans)
    pose = _c(pose)
    E, n, _ = humans.shape
    if walls is None or np.size(walls) == 0 or walls.shape[-4] == 0:
        W, S, per_env, walls_c = 0, 1, 0, np.zeros(4)
    else:
        walls_c = _c(walls)
        per_env = int(walls_c.ndim == 5)
        W, S = walls_c.shape[-4], walls_c.shape[-3]
    ranges = np.zeros((E, samples))
    hits = np.zeros((E, samples), np.int64)
    lib.orc_laser(E, n, W, S, per_env, _dp(humans), _dp(walls_c), _dp(pose), float(range_), int(samples), float(max_distance),
                  _dp(ranges), hits.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), int(n_threads))
    return ranges, hits
