"""Harness for the LIVE reference (test infrastructure only -- tests/, bench.py's cpu_baseline / --impl reference).

The reference imports pygame, gymnasium, rvo2, socialforce and matplotlib at module import time (social_gym/__init__.py:1,
src/agent.py:1, src/obstacle.py:1,7, src/motion_model_manager.py:8, social_nav_sim.py:1,14-15,28).  None of them is on the
SFM / HSFM arithmetic path, so `install()` puts inert stand-ins into sys.modules BEFORE the first `import social_gym`, restores
the `np.NaN` alias NumPy 2 removed (motion_model_manager.py:264,271) and puts the reference root on sys.path: /root/reference in
the build container, the copy staged by oracle/build.py::stage_reference() under oracle/_ref/ on the GPU box.
"""
import copy
import os
import sys
import time
import types
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def root():
    """Where the reference lives: $SNP_REFERENCE_ROOT, /root/reference, or the staged copy oracle/_ref (GPU box)."""
    cands = [os.environ.get("SNP_REFERENCE_ROOT"), "/root/reference", os.path.join(HERE, "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "social_gym")):
            return c
    return None


def available() -> bool:
    return root() is not None


class _Sprite:
    def __init__(self, *a, **k):
        pass


class _Group:
    def __init__(self, *a):
        self._items = list(a)

    def add(self, *items):
        self._items.extend(items)

    def empty(self):
        self._items.clear()

    def sprites(self):
        return list(self._items)

    def __len__(self):
        return len(self._items)

    def __iter__(self):
        return iter(self._items)


def install():
    """Install the stubs and put the reference on sys.path.  Idempotent."""
    if "social_gym" in sys.modules:
        return
    if not available():
        raise ImportError("the reference is neither at /root/reference nor staged under oracle/_ref (run oracle/build.py)")
    if not hasattr(np, "NaN"):
        np.NaN = np.nan
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/snp_numba_cache")

    gymn = types.ModuleType("gymnasium")

    class Env:
        pass

    gymn.Env = Env
    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Discrete = lambda n: n
    gymn.spaces = spaces
    envs = types.ModuleType("gymnasium.envs")
    reg = types.ModuleType("gymnasium.envs.registration")
    reg.register = lambda **kw: None
    envs.registration = reg
    gymn.envs = envs
    sys.modules["gymnasium"] = gymn
    sys.modules["gymnasium.spaces"] = spaces
    sys.modules["gymnasium.envs"] = envs
    sys.modules["gymnasium.envs.registration"] = reg

    pg = MagicMock()
    pg.sprite.Sprite = _Sprite
    pg.sprite.Group = _Group
    pg.time.get_ticks = lambda: 0
    sys.modules["pygame"] = pg
    sys.modules["pygame.sprite"] = pg.sprite

    sys.modules["rvo2"] = MagicMock()
    sys.modules["socialforce"] = MagicMock()
    mpl = MagicMock()
    mpl.colors.TABLEAU_COLORS = {"a": "#000000"}
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = mpl.pyplot
    sys.modules["matplotlib.colors"] = mpl.colors

    r = root()
    if r not in sys.path:
        sys.path.insert(0, r)


def sim_from_arrays(model, states, goals, walls=None, robot=None, robot_visible=False, dt=0.0125, parallel=False):
    """A headless reference SocialNavSim (scenario 'custom_config', social_nav_sim.py:58-104) holding exactly the crowd of ONE env
    of this repo's arrays: states [N,13] rows (agent.py:256), goals [N,G,2] NaN padded, walls = list of vertex lists, robot [13].
    Velocities are then set from the rows (the config format has none)."""
    install()
    from social_gym.social_nav_sim import SocialNavSim
    humans = {}
    for i, (s, g) in enumerate(zip(states, goals)):
        gl = [[float(a), float(b)] for a, b in g if a == a]
        humans[i] = {"pos": [float(s[0]), float(s[1])], "yaw": float(s[2]), "goals": gl, "radius": float(s[8]), "mass": float(s[9]),
                     "des_speed": float(s[12])}
    data = {"headless": True, "motion_model": model, "runge_kutta": False, "robot_visible": bool(robot_visible), "grid": False,
            "humans": humans, "walls": [] if walls is None else copy.deepcopy(walls)}
    if robot is not None:
        data["robot"] = {"pos": [float(robot[0]), float(robot[1])], "yaw": float(robot[2]), "radius": float(robot[8]),
                         "goals": [[float(robot[10]), float(robot[11])]]}
    sim = SocialNavSim(data, scenario="custom_config", parallelize_humans=parallel)
    sim.set_time_step(dt)
    for h, s in zip(sim.humans, states):
        h.linear_velocity = np.array(s[3:5], np.float64)
        h.body_velocity = np.array(s[5:7], np.float64)
        h.angular_velocity = float(s[7])
    if robot is not None:
        sim.robot.linear_velocity = np.array(robot[3:5], np.float64)
    return sim


def human_rows(sim):
    """[N,12] px,py,yaw,vx,vy,bvx,bvy,omega,gx,gy,desired_fx,desired_fy of the reference's humans."""
    return np.array([[h.position[0], h.position[1], h.yaw, h.linear_velocity[0], h.linear_velocity[1], h.body_velocity[0], h.body_velocity[1],
                      h.angular_velocity, h.goals[0][0], h.goals[0][1], h.desired_force[0], h.desired_force[1]] for h in sim.humans], np.float64)


def step_like_gym(sim, robot_vel, dt, n_substeps):
    """n_substeps x (robot.step(action, dt); update_humans(t, dt)): the sub-step loop of SocialNavGym.step (social_nav_gym.py:240-245)
    with a holonomic action (robot_agent.py:126-131)."""
    rv = np.array(robot_vel, np.float64)
    mm = sim.motion_model_manager
    for _ in range(n_substeps):
        sim.robot.position = sim.robot.position + rv * dt
        sim.robot.linear_velocity = rv.copy()
        mm.update_humans(0.0, dt)


def _worker(args):
    """One env per process (BASELINE.md section 4 / SURVEY 8d): build the env, warm up, time `seconds` of serial updates."""
    model, states, goals, walls, robot, robot_visible, dt, seconds, parallel = args
    sim = sim_from_arrays(model, states, goals, walls, robot, robot_visible, dt, parallel=parallel)
    step_like_gym(sim, (0.0, 1.0), dt, 10)
    n, t0 = 0, time.perf_counter()
    while True:
        step_like_gym(sim, (0.0, 1.0), dt, 20)
        n += 20
        el = time.perf_counter() - t0
        if el >= seconds:
            return n, el


def time_update_humans(model, states, goals, walls, robot, robot_visible, dt, n_procs, seconds, parallel=False):
    """Aggregate agent-steps/s of the reference's own update_humans on `n_procs` host cores, one env per process: env p of the
    batch goes to process p.  Returns (agent_steps_per_s, updates_done, wall_seconds)."""
    import multiprocessing as mp
    N = states.shape[1] - (1 if robot_visible else 0)
    jobs = []
    for p in range(n_procs):
        e = p % states.shape[0]
        rb = states[e, N] if robot_visible else (None if robot is None else robot[e])
        jobs.append((model, states[e, :N], goals[e], walls, rb, robot_visible, dt, seconds, parallel))
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(n_procs) as pool:
        res = pool.map(_worker, jobs)
    wall = time.perf_counter() - t0
    # every process ran for ~`seconds`; rate = sum over processes of its own updates / its own time
    rate = sum(N * n / el for n, el in res)
    return rate, sum(n for n, _ in res), wall


def _pool_worker(conn, args):
    model, states, goals, walls, robot, robot_visible, dt, parallel = args
    try:
        sim = sim_from_arrays(model, states, goals, walls, robot, robot_visible, dt, parallel=parallel)
        step_like_gym(sim, (0.0, 1.0), dt, 2)
        conn.send(("ready", 0.0))
        while True:
            msg = conn.recv()
            if msg is None:
                break
            t0 = time.perf_counter()
            step_like_gym(sim, (0.0, 1.0), dt, int(msg))
            conn.send(("done", time.perf_counter() - t0))
    except Exception as exc:  # surface the failure instead of hanging the parent
        conn.send(("error", repr(exc)))


class ReferencePool:
    """`n_procs` resident processes, ONE reference env each (env p of the batch): step(k) makes every process run
    k x (robot.step; update_humans) on its env and returns the wall time of the slowest, i.e. the time the host needs for one
    gym step of n_procs envs with all its cores busy.  Fork the pool BEFORE CUDA is initialised in the parent."""

    def __init__(self, n_procs, model, states, goals, walls, robot, robot_visible, dt=0.0125, parallel=False):
        import multiprocessing as mp
        install()
        ctx = mp.get_context("fork")
        self.N = states.shape[1] - (1 if robot_visible else 0)
        self.n_procs, self.procs, self.conns = n_procs, [], []
        for p in range(n_procs):
            e = p % states.shape[0]
            rb = states[e, self.N] if robot_visible else (None if robot is None else robot[e])
            a, b = ctx.Pipe()
            pr = ctx.Process(target=_pool_worker, args=(b, (model, states[e, :self.N], goals[e], walls, rb, robot_visible, dt, parallel)), daemon=True)
            pr.start()
            self.procs.append(pr); self.conns.append(a)
        for c in self.conns:
            kind, val = c.recv()
            if kind != "ready":
                raise RuntimeError(f"reference worker failed: {val}")

    def step(self, n_substeps):
        t0 = time.perf_counter()
        for c in self.conns:
            c.send(int(n_substeps))
        inner = []
        for c in self.conns:
            kind, val = c.recv()
            if kind != "done":
                raise RuntimeError(f"reference worker failed: {val}")
            inner.append(val)
        return time.perf_counter() - t0, inner

    def close(self):
        for c in self.conns:
            try:
                c.send(None)
            except (BrokenPipeError, OSError):
                pass
        for p in self.procs:
            p.join(timeout=5)
