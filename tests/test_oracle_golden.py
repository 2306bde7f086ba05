"""Pin the CPU oracle (oracle/snp_oracle.c) against golden vectors recorded from the LIVE reference
(tests/golden/make_golden.py).  CPU only.  These tests are what makes the oracle trustworthy at the sizes
the Python reference cannot reach."""
import os

import numpy as np
import pytest

import oracle
from oracle import OracleConfig
from helpers import (GOLDEN, traj_names, load_traj, rel_err, inputs_at, observed, consecutive_pairs)

# fp64 restatement vs the reference's own fp64: only libm-vs-NumPy ulp noise is allowed.
TOL_STEP = 1e-12
# Moussaid at rest: theta_ij == 0 up to rounding and k_ij = sign(theta_ij) (forces.py:100-101) is decided by the
# last ulp of atan2; NumPy's arctan2 and libm's atan2 differ there, so the REFERENCE is discontinuous at those
# states (its own serial and Numba paths disagree by 3e-3 relative).  Bounded, not hidden:
TOL_MOUSSAID_AT_REST = 1e-6


def _cfg(d, numba=False):
    return OracleConfig(int(d["type"]), d["consider_robot"], d["all_equal"], numba, d.get("respawn_bounds"))


@pytest.mark.parametrize("name", traj_names())
def test_single_step_matches_reference(name):
    d = load_traj(name)
    n, cfg = d["n"], _cfg(d)
    moussaid = int(d["type"]) % 3 == 2
    worst = 0.0
    for k in consecutive_pairs(d):
        S, G, D, rv = inputs_at(d, k)
        saf = d["safety"][: S.shape[0]]
        S2, G2, D2 = oracle.update_humans(cfg, S[None], G[None], d["walls"], d["params"][None], saf[None], D[None],
                                          float(d["dt"]), 1, robot_vel=rv[None] if d["consider_robot"] else None)
        err = rel_err(observed(S2[0], D2[0], n), d["traj"][k + 1]).max()
        at_rest = moussaid and np.all(d["traj"][k][:, 3:5] == 0.0)
        assert err <= (TOL_MOUSSAID_AT_REST if at_rest else TOL_STEP), (name, int(d["steps"][k]), err)
        worst = max(worst, err)
    assert worst < TOL_MOUSSAID_AT_REST


@pytest.mark.parametrize("name", traj_names())
def test_forces_of_first_update(name):
    """desired / obstacle / social / torque / global force of every human after the first update."""
    d = load_traj(name)
    S, G, D, rv = inputs_at(d, 0)
    saf = d["safety"][: S.shape[0]]
    *_, F = oracle.update_humans(_cfg(d), S[None], G[None], d["walls"], d["params"][None], saf[None], D[None],
                                 float(d["dt"]), 1, robot_vel=rv[None] if d["consider_robot"] else None, want_forces=True)
    ref = d["forces1"]
    headed = int(d["type"]) >= 3
    if not headed:
        F, ref = F[..., [0, 1, 2, 3, 4, 5, 7, 8]], ref[..., [0, 1, 2, 3, 4, 5, 7, 8]]  # torque unused for SFM
    # the first update starts from rest: for Moussaid the k_ij sign flip moves the small angular term (see above)
    tol = 1e-3 if int(d["type"]) % 3 == 2 else 1e-12
    assert rel_err(F[0], ref).max() <= tol


@pytest.mark.parametrize("name", traj_names())
def test_multi_step_trajectory(name):
    """Whole recorded trajectory (up to 1600 updates) from the initial state: divergence stays tiny in fp64."""
    d = load_traj(name)
    n, cfg = d["n"], _cfg(d)
    S, G, D, rv = inputs_at(d, 0)
    saf = d["safety"][: S.shape[0]]
    S, G, D = S[None], G[None], D[None]
    cur, worst = 0, 0.0
    for k, s in enumerate(d["steps"]):
        if s > cur:
            S, G, D = oracle.update_humans(cfg, S, G, d["walls"], d["params"][None], saf[None], D, float(d["dt"]), int(s - cur),
                                           robot_vel=rv[None] if d["consider_robot"] else None)
            cur = s
        worst = max(worst, rel_err(observed(S[0], D[0], n), d["traj"][k]).max())
    assert worst < (1e-5 if int(d["type"]) % 3 == 2 else 1e-9), worst


def test_goal_rotation_and_stale_desired_force_are_exercised():
    """The fixtures must actually cover goal switching (mmm:66-70) and the serial path's stale desired force
    inside the goal radius (forces.py:12-15), otherwise the pin is hollow."""
    d = load_traj("corridor_sfm_guo")
    assert len({tuple(g) for g in d["traj"][:, 0, 8:10]}) >= 3          # the goal list rotated at least twice
    d = load_traj("jym_hsfm_new_guo")
    last = d["traj"][-1]
    dist = np.linalg.norm(last[:, 8:10] - last[:, 0:2], axis=1)
    inside = dist <= 0.3
    assert inside.any() and np.abs(last[inside, 10:12]).max() > 0           # inside the radius, force not zeroed


def test_parallel_traffic_respawn_is_exercised():
    """The pt* fixtures must contain respawns (goal y and position x jump) so that mmm:407-422 is really pinned."""
    for name in [n for n in traj_names() if n.startswith("pt")]:
        d = load_traj(name)
        assert d["respawn_bounds"] == (7.0, 1.5)
        jumps = np.diff(d["traj"][:, :, 0], axis=0) > 5.0
        assert jumps.sum() >= 5, name


def _il_keys(z):
    return sorted(k[:-8] for k in z.files if k.endswith("_states0"))


def test_robot_driven_by_motion_model():
    """update_robot + update_humans per sub-step (imitation_learning_step, gym:260-265; mmm:593-653) against 6 recorded runs:
    robot and human models may differ, robot visible or not, walls, a robot goal switch."""
    z = np.load(os.path.join(GOLDEN, "il_robot.npz"))
    keys = _il_keys(z)
    assert len(keys) == 6
    for key in keys:
        vis, equal = (bool(v) for v in z[key + "_flags"])
        S, G, rb = z[key + "_states0"], z[key + "_goals0"][None], z[key + "_robot0"][None].copy()
        n = S.shape[0]
        S = (np.concatenate([S, rb], 0) if vis else S)[None]
        cfg = OracleConfig(int(z[key + "_type"]), vis, equal, False)
        D, rD, rG = np.zeros((1, n, 2)), np.zeros((1, 2)), z[key + "_robot_goals"][None].copy()
        cur = 0
        moussaid = int(z[key + "_type"]) % 3 == 2 or int(z[key + "_robot_type"]) % 3 == 2
        for k, s_ in enumerate(z[key + "_steps"]):
            if s_ > cur:
                S, G, D, rb, rG, rD = oracle.imitation_steps(cfg, S, G, z[key + "_walls"], z[key + "_params"][None], np.zeros((1, S.shape[1])), D,
                                                            0.0125, int(s_ - cur), rb, rG, rD, z[key + "_robot_params"], int(z[key + "_robot_type"]))
                cur = s_
            got_h = np.concatenate([S[0, :n, :8], S[0, :n, 10:12], D[0]], 1)
            got_r = np.concatenate([rb[0, :8], rb[0, 10:12], rD[0]])
            tol = 1e-5 if moussaid else 1e-9
            assert rel_err(got_h, z[key + "_traj"][k]).max() < tol, (key, int(s_))
            assert rel_err(got_r, z[key + "_robot_traj"][k]).max() < tol, (key, int(s_))
    rt = z["cc5_near_goal_hsfm_guo__hsfm_guo_robot_traj"]
    assert (np.abs(np.diff(rt[:, 8:10], axis=0)).sum(1) > 0).sum() == 1   # the robot's goal list rotated once


def test_sim_update_with_model_driven_robot():
    """SocialNavSim.update with a model-driven robot (sim:476-529), recorded by calling sim.update() of the live reference: pose advance every
    update (unwrapped yaw), update_robot(just_velocities=True) every ROBOT_SAMPLING_TIME with that dt -- or a full update_robot when the
    sampling times are equal -- and humans that see the robot's PREVIOUS state.  4 runs: 20 / 4 / 1 updates per robot step, robot visible or
    not, walls, a robot goal switch, a run whose unwrapped yaw reaches -500 rad."""
    z = np.load(os.path.join(GOLDEN, "sim_update.npz"))
    keys = sorted(k[:-len("_robot_type")] for k in z.files if k.endswith("_robot_type"))
    assert len(keys) == 4
    for key in keys:
        vis, equal = (bool(v) for v in z[key + "_flags"])
        S, G, rb = z[key + "_states0"], z[key + "_goals0"][None], z[key + "_robot0"][None].copy()
        n = S.shape[0]
        S = (np.concatenate([S, rb], 0) if vis else S)[None]
        cfg = OracleConfig(int(z[key + "_type"]), vis, equal, False)
        D, rD, rG = np.zeros((1, n, 2)), np.zeros((1, 2)), z[key + "_robot_goals"][None].copy()
        cur = 0
        for k, s_ in enumerate(z[key + "_steps"]):
            if s_ > cur:
                S, G, D, rb, rG, rD = oracle.sim_update_steps(cfg, S, G, z[key + "_walls"], z[key + "_params"][None], np.zeros((1, S.shape[1])), D,
                                                              0.0125, int(s_ - cur), rb, rG, rD, z[key + "_robot_params"], int(z[key + "_robot_type"]),
                                                              int(z[key + "_every"]), float(z[key + "_robot_dt"]), phase=int(cur))
                cur = s_
            got_h = np.concatenate([S[0, :n, :8], S[0, :n, 10:12], D[0]], 1)
            got_r = np.concatenate([rb[0, :8], rb[0, 10:12], rD[0]])
            assert rel_err(got_h, z[key + "_traj"][k]).max() < 1e-9, (key, int(s_))
            assert rel_err(got_r, z[key + "_robot_traj"][k]).max() < 1e-9, (key, int(s_), got_r, z[key + "_robot_traj"][k])
    rt = z["cc5_near_goal_hsfm_guo__hsfm_guo_rt20_robot_traj"]
    assert (np.abs(np.diff(rt[:, 8:10], axis=0)).sum(1) > 0).sum() == 1 and np.abs(rt[:, 2]).max() > 100.0


def test_lookahead_rewards_and_rotated_states():
    """compute_rotated_states_and_reward (cadrl.py:42-83) for 81 actions: rewards bit-exact, rotated states to 1e-13."""
    z = np.load(os.path.join(GOLDEN, "lookahead.npz"))
    keys = sorted(k[:-4] for k in z.files if k.endswith("_cur"))
    assert len(keys) == 18
    seen = set()
    for key in keys:
        vis = key.endswith("_1")
        rot, rew = oracle.lookahead(z[key + "_cur"][None], z[key + "_next"][None], z[key + "_robot"][None], z["actions"], 0.25, visible=vis)
        assert np.array_equal(rew[0], z[key + "_rewards"]), key
        assert np.abs(rot[0] - z[key + "_rotated"]).max() < 1e-13, key
        seen |= set(np.unique(np.sign(rew)))
    assert seen == {-1.0, 0.0, 1.0}


def test_robot_push_out_bit_exact():
    """RobotAgent.check_collisions (robot_agent.py:35-48) from 64 start positions recorded from the live reference: bit for bit."""
    z = np.load(os.path.join(GOLDEN, "push_out.npz"))
    moved = 0
    for name in ("walls", "cc"):
        H, st, en, r = z[name + "_humans"], z[name + "_start"], z[name + "_end"], float(z[name + "_radius"])
        rb = np.concatenate([st, np.full((len(st), 1), r)], 1)
        out = oracle.robot_push_out(np.tile(H, (len(st), 1, 1)), z[name + "_walls"], rb)
        assert np.array_equal(out, en), name
        moved += int((np.abs(st - en).sum(1) > 0).sum())
    assert moved >= 30


def test_numba_operator_semantics():
    """Second witness: numba_compat=1 reproduces forces_parallel.update_humans_parallel (fp:184), including the
    Guo wall force divided by the wall count (fp:161) and '<=' goal switching (fp:226)."""
    z = np.load(os.path.join(GOLDEN, "numba_operator.npz"))
    keys = sorted(k[:-8] for k in z.files if k.endswith("_states0"))
    assert len(keys) == 18
    for key in keys:
        typ, equal, robot = (int(v) for v in z[key + "_flags"])
        S, G = z[key + "_states0"][None], z[key + "_goals0"][None]
        n = G.shape[1]
        D = np.zeros((1, n, 2))
        cfg = OracleConfig(typ, bool(robot), bool(equal), True)
        for step in range(3):
            S, G, D = oracle.update_humans(cfg, S, G, z[key + "_walls"], z[key + "_params"][None], z[key + "_safety"][None], D,
                                           0.0125, 1)
            ref = z[key + "_out"][step]
            tol = 1e-6 if typ % 3 == 2 else 1e-12
            assert rel_err(S[0, :n, :8], ref[:n, :8]).max() <= tol, (key, step)
            assert np.array_equal(S[0, :n, 10:12], ref[:n, 10:12])


def test_peek_next_observable_states():
    """get_next_human_observable_states (mmm:691-709): one update at dt=0.25, state restored afterwards."""
    z = np.load(os.path.join(GOLDEN, "peek.npz"))
    for model in ["sfm_helbing", "hsfm_farina", "hsfm_new_guo"]:
        S = np.concatenate([z[model + "_states"], z[model + "_robot"][None]], 0)[None]
        n = S.shape[1] - 1
        cfg = OracleConfig(int(z[model + "_type"]), True, True, False)
        S2, G2, D2 = oracle.update_humans(cfg, S, z[model + "_goals"][None], None, z[model + "_params"][None],
                                          np.zeros((1, n + 1)), z[model + "_desired"][None], 0.25, 1)
        obs4 = S2[0, :n][:, [0, 1, 3, 4]]
        assert rel_err(obs4, z[model + "_obs4"]).max() < 1e-12
        obs8 = S2[0, :n][:, [0, 1, 2, 3, 4, 7, 10, 11]]
        assert rel_err(obs8, z[model + "_obs8"]).max() < 1e-12
        # the reference restores pose/velocity/goal exactly; only desired_force keeps the peeked value
        assert np.array_equal(z[model + "_before"][:, :10], z[model + "_after"][:, :10])
        assert rel_err(D2[0], z[model + "_after"][:, 10:12]).max() < 1e-12


def test_collision_goal_reward_flags_bit_exact():
    z = np.load(os.path.join(GOLDEN, "flags.npz"))
    H, R, A, ref = z["humans"], z["robot"], z["action"], z["result"]
    out = oracle.checks(H, H.shape[1], R, A, ref[:, 11], z["consts"])
    for col in (0, 2, 4, 5, 6, 7, 9, 10):   # flags and info codes
        assert np.array_equal(out[:, col], ref[:, col]), col
    for col in (1, 3, 8):                    # dmin, reward, actual dmin: same operations -> same bits
        assert np.array_equal(out[:, col], ref[:, col]), col
    assert ref[:, 0].sum() > 50 and (ref[:, 6] == 4).sum() > 20 and ref[:, 2].sum() > 10


def test_laser_ranges_and_hit_indices():
    z = np.load(os.path.join(GOLDEN, "laser.npz"))
    keys = sorted(k[:-5] for k in z.files if k.endswith("_pose"))
    assert len(keys) == 15
    total_hits = 0
    for key in keys:
        x, y, yaw, rng, samples, maxd = z[key + "_pose"]
        ranges, hits = oracle.laser(z[key + "_humans"][None], z[key + "_walls"], np.array([[x, y, yaw]]), rng, int(samples), maxd)
        assert np.array_equal(hits[0], z[key + "_hits"]), key
        assert np.abs(ranges[0] - z[key + "_ranges"]).max() <= 1e-12, key
        total_hits += int((hits >= 0).sum())
    assert total_hits > 1000


def test_gym_step_sequence():
    """SocialNavGym.step (gym:227-250): swept collision/goal test + reward on the pre-step state, then 20 sub-steps of
    robot.step + update_humans; observation = (px,py,vx,vy,r) per human."""
    z = np.load(os.path.join(GOLDEN, "gym_step.npz"))
    consts = np.array([50, -0.25, 1.0, 0.2, 0.5, 0.25])
    for key in ["hsfm_farina_0", "sfm_helbing_1", "hsfm_new_guo_1"]:
        visible = key.endswith("_1")
        S, G, rb = z[key + "_states0"], z[key + "_goals0"][None], z[key + "_robot0"].copy()
        n = S.shape[0]
        cfg = OracleConfig(int(z[key + "_type"]), visible, True, False)
        D = np.zeros((1, n, 2))
        S = (np.concatenate([S, rb[None]], 0) if visible else S)[None]
        t = 0.0
        for k, a in enumerate(z[key + "_actions"]):
            out = oracle.checks(S, n, rb[None], a[None], np.array([t]), consts)
            ref = z[key + "_result"][k]
            assert out[0, 3] == ref[0] and out[0, 4] == ref[1] and out[0, 5] == ref[2] and out[0, 6] == ref[3], (key, k)
            if visible:
                S[0, n] = rb
                S, G, D = oracle.update_humans(cfg, S, G, None, z[key + "_params"][None], np.zeros((1, n + 1)), D, 0.0125, 20,
                                               robot_vel=a[None])
                rb = S[0, n].copy()
            else:
                S, G, D = oracle.update_humans(cfg, S, G, None, z[key + "_params"][None], np.zeros((1, n)), D, 0.0125, 20)
                for _ in range(20):
                    rb[0:2] = rb[0:2] + a * 0.0125
                rb[3:5] = a
            for _ in range(20):
                t += 0.0125
            assert rel_err(rb[0:2], z[key + "_robot_pos"][k]).max() < 1e-13
            obs = S[0, :n][:, [0, 1, 3, 4, 8]]
            assert rel_err(obs, z[key + "_obs"][k]).max() < 1e-9, (key, k)
