"""Shared helpers for the parity tests: load golden fixtures recorded from the live reference
(tests/golden/make_golden.py) and rebuild the inputs of a single update from a recorded row."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def traj_names():
    return sorted(os.path.basename(f)[5:-4] for f in glob.glob(os.path.join(GOLDEN, "traj_*.npz")))


def load_traj(name):
    d = dict(np.load(os.path.join(GOLDEN, f"traj_{name}.npz")))
    d["name"] = name
    d["consider_robot"], d["all_equal"] = bool(d["flags"][0]), bool(d["flags"][1])
    d["n"] = d["states0"].shape[0]
    d["respawn_bounds"] = tuple(float(x) for x in d["respawn"]) if "respawn" in d else None
    return d


def rel_err(got, ref, scale=1.0):
    """|got-ref| / max(|ref|, scale): the parity metric of SURVEY.md section 8(d)."""
    return np.abs(got - ref) / np.maximum(np.abs(ref), scale)


def rotate_goals_to(goals0, current):
    """Rotate each NaN-padded goal list left until its head equals `current` (what the reference's list
    rotation mmm:66-70 has done by the time `current` was recorded)."""
    G = goals0.copy()
    for i in range(G.shape[0]):
        cnt = int((~np.isnan(G[i, :, 0])).sum())
        for _ in range(cnt):
            if np.array_equal(G[i, 0], current[i]):
                break
            G[i, :cnt] = np.roll(G[i, :cnt], -1, axis=0)
        assert np.array_equal(G[i, 0], current[i]), "recorded goal not in goal list"
    return G


def inputs_at(d, k):
    """Inputs of the update that turns recorded row k into row k+1 (requires steps[k+1] == steps[k]+1).
    Returns states [N(+1),13], goals [N,G,2], desired [N,2], robot_vel [2]."""
    n = d["n"]
    row = d["traj"][k]
    S = d["states0"].copy()
    S[:, :8] = row[:, :8]
    S[:, 10:12] = row[:, 8:10]
    if d.get("respawn_bounds") is not None:  # parallel traffic: a respawn replaces the goal list by the single recorded goal
        G = np.full_like(d["goals0"], np.nan)
        G[:, 0] = row[:, 8:10]
    else:
        G = rotate_goals_to(d["goals0"], row[:, 8:10])
    if d["consider_robot"]:
        rb = d["robot0"].copy()
        rb[0:2] = d["robot_traj"][k]
        S = np.concatenate([S, rb[None]], 0)
    return S, G, row[:, 10:12].copy(), d["robot_vel"].copy()


def observed(states, desired, n):
    """Pack oracle/engine output into the fixture's 12-column row layout."""
    return np.concatenate([states[:n, :8], states[:n, 10:12], desired[:n]], 1)


def consecutive_pairs(d):
    s = d["steps"]
    return [k for k in range(len(s) - 1) if s[k + 1] == s[k] + 1]


def wrap_angle_cols(got, ref, col=2):
    """Headings are angles: -pi and +pi are the same heading (bound_angle, utils.py:7-13, wraps at +-pi, and an fp32 -pi lies
    on the other side of the fp64 one).  Returns a copy of `got` whose heading column is shifted by the multiple of 2*pi
    that brings it closest to `ref`."""
    out = got.copy()
    out[..., col] = ref[..., col] + (got[..., col] - ref[..., col] + np.pi) % (2 * np.pi) - np.pi
    return out


def moussaid_rest_ambiguity(states, n, dt, Ei=360.0, gamma=0.35):
    """Upper bound of what ONE step from rest can differ by when the sign k_ij = sign(theta_ij) of Moussaid's lateral term
    flips (forces.py:100-110): theta_ij is zero up to rounding when both agents are at rest, so k_ij in {-1,0,+1} is decided by
    the last ulp of atan2 -- in the reference itself.  Per human i the lateral force is ambiguous by at most
    sum_j 2*Ei*exp(-d_ij/gamma) (|interaction vector| = 1 at rest), i.e. a velocity change of that times dt/m; the HSFM torque
    law turns the same force ambiguity into at most k_lambda*(pi+1)*m times as much angular velocity."""
    p = states[:, 0:2]
    m = states[:n, 9]
    bound = np.zeros(n)
    for i in range(n):
        d = np.linalg.norm(p[i] - p, axis=1)
        d[i] = np.inf
        bound[i] = (2 * Ei * np.exp(-d / gamma)).sum() * dt / m[i]
    return bound, 0.1 * (np.pi + 1) * m * bound


def unicycle_step(pos, yaw, v, r, dt):
    """RobotAgent.step with unicycle kinematics, action = ActionRot(v, r) (robot_agent.py:116-136), restated with the same NumPy
    expressions: returns (new position, new yaw, new linear velocity).  Pinned against the live reference in
    tests/test_oracle_live_reference.py::test_unicycle_robot_step_matches_the_live_reference."""
    act = np.array([np.cos(yaw + r) * v, np.sin(yaw + r) * v], np.float64)
    pos = np.asarray(pos, np.float64) + act * dt
    yaw = (yaw + r) % (2 * np.pi)
    return pos, yaw, np.array([np.cos(yaw) * v, np.sin(yaw) * v], np.float64)


def constant_velocity_next(cur, dt, visible):
    """propagate_humans_state_with_constant_velocity_model (crowd_nav/policy/cadrl.py:92-105) restated in NumPy: cur [..., N, 5|7] =
    x, y, vx, vy, radius(, theta, omega) -> [..., N, 4|6] = x, y, (yaw,) Vx, Vy(, Omega).  Pinned against the live function in
    tests/test_oracle_live_reference.py."""
    x, y = cur[..., 0] + cur[..., 2] * dt, cur[..., 1] + cur[..., 3] * dt
    if visible:
        return np.stack([x, y, cur[..., 5] + cur[..., 6] * dt, cur[..., 2], cur[..., 3], cur[..., 6]], -1)
    return np.stack([x, y, cur[..., 2], cur[..., 3]], -1)
