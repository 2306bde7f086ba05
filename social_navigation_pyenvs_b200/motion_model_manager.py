"""MotionModelManager: object-level drop-in for social_gym/src/motion_model_manager.py (SFM / HSFM Euler path only).

Same constructor and method names as the reference class, operating on the reference's own agent objects (anything with
`position, yaw, linear_velocity, body_velocity, angular_velocity, radius, mass, desired_speed, goals, safety_space`), for ONE
environment.  Every `update_humans` call packs the objects into the 13-wide rows of agent.py:256, runs the CUDA step and writes
the rows back exactly like motion_model_manager.py:358-367 does for its Numba path -- but with the serial path's semantics
(the parity oracle).  For thousands of environments use `CrowdEngine` directly: state then stays on the device.

Out of scope here, as in SURVEY.md section 2: ORCA (rvo2), social momentum, RK45 integration, sfm_roboticsupo.
"""
import numpy as np

from .engine import CrowdEngine, SFMS, model_parameters
from .sensors import _walls_array

N_GENERAL_STATES, N_HEADED_STATES, N_NOT_HEADED_STATES = 8, 6, 4


def _row(agent):
    # Agent.get_safe_state (agent.py:256-258)
    return [agent.position[0], agent.position[1], agent.yaw, agent.linear_velocity[0], agent.linear_velocity[1],
            agent.body_velocity[0], agent.body_velocity[1], agent.angular_velocity, agent.radius, agent.mass,
            agent.goals[0][0], agent.goals[0][1], agent.desired_speed]


class MotionModelManager:
    def __init__(self, motion_model_title, consider_robot, runge_kutta, humans, robot, walls, parallelize=False, dtype="float64"):
        if runge_kutta:
            raise NotImplementedError("RK45 integration is out of scope of the B200 engine (the gym always uses Euler, gym:142)")
        self.consider_robot = consider_robot
        self.runge_kutta = False
        self.update_targets = True
        self.humans, self.robot = humans, robot
        self.walls = walls
        self.parallel = parallelize
        self.orca = self.sm = self.sf = False
        self.parallel_traffic_humans_respawn = False
        self.robot_motion_model_title = None
        self._dtype = dtype
        self.set_human_motion_model(motion_model_title)

    # ------------------------------------------------------------------ model selection (mmm:222-283)
    def set_human_motion_model(self, motion_model_title):
        if motion_model_title in ("orca", "social_momentum", "sfm_roboticsupo"):
            raise NotImplementedError(f"Model {motion_model_title} is not implemented for humans in the B200 engine")
        if motion_model_title not in SFMS:
            raise Exception(f"The human motion model '{motion_model_title}' does not exist")
        self.motion_model_title = motion_model_title
        self.type = SFMS.index(motion_model_title) % 3
        self.sfm_type = SFMS.index(motion_model_title)
        self.headed = self.sfm_type >= 3
        self.include_mass = True
        self.params = model_parameters(motion_model_title)
        for h in self.humans:
            if not hasattr(h, "safety_space"):
                h.safety_space = 0
            if not hasattr(h, "desired_force"):
                h.desired_force = np.zeros(2)
        # all_equal_humans keeps only the verdict of the LAST pair compared (mmm:278-283): radius and mass of the last two humans
        # (model parameters are identical for every human, agent.py:79-243)
        n = len(self.humans)
        self.all_equal_humans = True if n < 2 else (self.humans[-2].radius == self.humans[-1].radius and self.humans[-2].mass == self.humans[-1].mass)
        self._engine, self._engine_key = None, None

    def _goal_rows(self):
        g = max(len(h.goals) for h in self.humans)
        out = np.full((1, len(self.humans), g, 2), np.nan)
        for i, h in enumerate(self.humans):
            out[0, i, : len(h.goals)] = np.asarray(h.goals, np.float64)
        return out

    def _pack(self):
        rows = [_row(h) for h in self.humans]
        safety = [h.safety_space for h in self.humans]
        if self.consider_robot:
            rows.append(_row(self.robot))
            safety.append(getattr(self.robot, "safety_space", 0))
        return np.array([rows], np.float64), np.array([safety], np.float64)

    def _sync_engine(self):
        import torch
        rows, safety = self._pack()
        goals = self._goal_rows()
        dt = torch.float64 if self._dtype == "float64" else torch.float32
        eng = self._engine
        walls = _walls_array(self.walls)
        # everything the engine was built from is part of the reuse key: changing consider_robot, all_equal_humans or the walls after
        # construction rebuilds it instead of being silently ignored
        key = (len(self.humans), goals.shape[2], self.motion_model_title, bool(self.consider_robot), bool(self.all_equal_humans), dt,
               None if walls is None else walls.tobytes())
        if eng is None or key != self._engine_key:
            eng = CrowdEngine.from_reference_arrays(self.motion_model_title, rows, goals, walls=walls, safety=safety,
                                                    consider_robot=self.consider_robot, all_params_equal=self.all_equal_humans, dtype=dt)
            self._engine, self._engine_key = eng, key
        else:
            eng.load_rows(rows, safety)
            eng.load_goals(goals)
        eng.set_desired_force(np.array([[h.desired_force for h in self.humans]], np.float64))
        return eng, rows

    def _write_back(self, eng, rows):
        out = eng.rows(rows)[0]
        df = eng.desired_force()[0]
        for i, h in enumerate(self.humans):
            h.position, h.yaw = out[i, 0:2].copy(), float(out[i, 2])            # Agent.set_state (agent.py:260-266)
            h.linear_velocity, h.body_velocity = out[i, 3:5].copy(), out[i, 5:7].copy()
            h.angular_velocity = float(out[i, 7])
            h.desired_force = df[i].copy()
            if not np.array_equal(np.array(h.goals[0], np.float64), out[i, 10:12]):
                if any(np.array_equal(np.array(g, np.float64), out[i, 10:12]) for g in h.goals):  # goal reached: rotate (mmm:364-367)
                    goal = h.goals[0]
                    h.goals.remove(goal)
                    h.goals.append(goal)
                else:                                                                          # respawn: new single goal (mmm:418)
                    h.goals = [[float(out[i, 10]), float(out[i, 11])]]

    # ------------------------------------------------------------------ the hot path (mmm:354-373)
    def update_humans(self, t, dt, post_update=True):
        eng, rows = self._sync_engine()
        # parallel-traffic respawn (mmm:407-422) runs inside the kernel; the caller sets parallel_traffic_humans_respawn and
        # respawn_bounds exactly like social_nav_sim.py:184-186 does
        eng.respawn_bounds = tuple(self.respawn_bounds) if (post_update and self.parallel_traffic_humans_respawn) else None
        eng.update_humans(t, dt, post_update=post_update)
        self._write_back(eng, rows)

    # ------------------------------------------------------------------ state accessors (mmm:285-352)
    def get_human_states(self, include_goal=True, headed=False):
        n = len(self.humans)
        if include_goal:
            state = np.empty([n, N_GENERAL_STATES])
            for i, h in enumerate(self.humans):
                v = h.body_velocity if headed else h.linear_velocity
                state[i] = [h.position[0], h.position[1], h.yaw, v[0], v[1], h.angular_velocity, h.goals[0][0], h.goals[0][1]]
        elif headed:
            state = np.empty([n, N_HEADED_STATES])
            for i, h in enumerate(self.humans):
                state[i] = [h.position[0], h.position[1], h.yaw, h.body_velocity[0], h.body_velocity[1], h.angular_velocity]
        else:
            state = np.empty([n, N_NOT_HEADED_STATES])
            for i, h in enumerate(self.humans):
                state[i] = [h.position[0], h.position[1], h.linear_velocity[0], h.linear_velocity[1]]
        return state

    def set_human_states(self, state, just_visual=False):
        for i, h in enumerate(self.humans):
            h.position[0], h.position[1], h.yaw = state[i, 0], state[i, 1], state[i, 2]
            if just_visual:
                continue
            if self.headed:
                h.body_velocity[0], h.body_velocity[1] = state[i, 3], state[i, 4]
                c, s = np.cos(h.yaw), np.sin(h.yaw)                              # headed_agent_update_linear_velocity (mmm:143-145)
                h.linear_velocity = np.matmul(np.array([[c, -s], [s, c]]), h.body_velocity)
            else:
                h.linear_velocity[0], h.linear_velocity[1] = state[i, 3], state[i, 4]
            h.angular_velocity = state[i, 5]
            goal = [state[i, 6], state[i, 7]]                                    # rewind_goals (mmm:57-64)
            if goal not in [list(g) for g in h.goals]:
                h.goals = [goal]
            else:
                while list(h.goals[0]) != goal:
                    h.goals.append(h.goals.pop(0))

    def get_next_human_observable_states(self, dt, theta_and_omega_visible=False):
        current = self.get_human_states(include_goal=True, headed=self.headed)
        self.update_humans(0, dt, post_update=False)
        nxt = self.get_human_states(include_goal=True, headed=False) if theta_and_omega_visible else self.get_human_states(False, False)
        self.set_human_states(current)
        return nxt

    def set_safety_space(self, safety_space):
        if "sfm" in self.motion_model_title:                                     # mmm:147-164
            for h in self.humans:
                h.safety_space = 0.01 + safety_space
        else:
            raise NotImplementedError(f"Model {self.motion_model_title} is not implemented for humans")
        if self.robot_motion_model_title is not None and "sfm" in self.robot_motion_model_title:
            self.robot.safety_space = 0.01 + safety_space
