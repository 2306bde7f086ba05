"""The policy-side lookahead producer on the GPU: peek (snp_step with dyn_out) + snp_lookahead against
the recorded outputs of the reference's compute_rotated_states_and_reward (crowd_nav/policy/cadrl.py:42-83) and against the
oracle on seeded batches; full-size run checked through size-independent properties."""
import os

import numpy as np
import pytest
import torch

import oracle
from helpers import GOLDEN, rel_err

pytestmark = pytest.mark.gpu


def _robot_row(rs):
    """robot_state [px,py,vx,vy,r,gx,gy,vd,theta] (cadrl.py:243) -> reference state row (agent.py:256)."""
    row = np.zeros(13)
    row[0], row[1], row[2], row[3], row[4] = rs[0], rs[1], rs[8], rs[2], rs[3]
    row[8], row[9], row[10], row[11], row[12] = rs[4], 80.0, rs[5], rs[6], rs[7]
    return row


@pytest.mark.parametrize("model", ["hsfm_farina", "sfm_helbing", "hsfm_new_guo"])
def test_lookahead_vs_reference_golden(model):
    """81 actions x N humans from recorded states: rewards bit-exact, rotated states within 1e-9 (the peek included)."""
    from social_navigation_pyenvs_b200 import CrowdEngine
    z = np.load(os.path.join(GOLDEN, "lookahead.npz"))
    for rep in range(3):
        for vis in (False, True):
            key = f"{model}_{rep}_{int(vis)}"
            eng = CrowdEngine.from_reference_arrays(model, z[f"{model}_{rep}_states"][None], z[f"{model}_{rep}_goals"][None], consider_robot=False,
                                                    all_params_equal=True, robot=_robot_row(z[key + "_robot"])[None])
            eng.set_desired_force(z[f"{model}_{rep}_desired"][None])
            eng.set_action_space(z["actions"])
            before = eng.dyn[:8].clone()
            rot, rew = eng.lookahead(0.25, theta_and_omega_visible=vis)
            assert np.array_equal(rew[0].cpu().numpy(), z[key + "_rewards"]), key
            assert rel_err(rot[0].cpu().numpy(), z[key + "_rotated"]).max() < 1e-9, key
            assert torch.equal(eng.dyn[:8], before)  # the peek leaves pose and velocities alone
            # the observable states the reference's peek returned
            nxt = eng.get_next_human_observable_states(0.25, theta_and_omega_visible=vis)[0]
            ref = z[key + "_next"]
            assert rel_err(nxt[:, :ref.shape[1]], ref).max() < 1e-9, key


def _batch(E, N, seed):
    from social_navigation_pyenvs_b200 import scenarios
    sc = scenarios.ccso_synthetic(E, N, seed)
    rng = np.random.RandomState(seed)
    S = sc["states"].copy()
    S[:, :, 3:5] = rng.uniform(-1, 1, (E, N, 2))          # moving humans so that the swept test has something to sweep
    S[:, :, 5:7] = S[:, :, 3:5]
    R = sc["robot"].copy()
    R[:, 0:2] = S[np.arange(E), rng.randint(N, size=E), 0:2] + rng.uniform(-1.5, 1.5, (E, 2))  # near a human: collisions, discomfort
    R[::7, 0:2] = R[::7, 10:12] - [0.05, 0.1]                                                       # next to the goal
    R[:, 12] = 1.0
    return sc, S, R


def _actions():
    speeds = [(np.exp((i + 1) / 5) - 1) / (np.e - 1) for i in range(5)]
    rot = np.linspace(0, 2 * np.pi, 16, endpoint=False)
    return np.array([[0.0, 0.0]] + [[s * np.cos(r), s * np.sin(r)] for r in rot for s in speeds])


def _oracle_inputs(eng, nxt, vis, sel=slice(None)):
    """The engine's own device state in the array forms compute_rotated_states_and_reward takes (cadrl.py:43-49)."""
    from social_navigation_pyenvs_b200 import _lib as L
    d, st, rb = eng.dyn.double().cpu().numpy()[:, sel], eng.stat.double().cpu().numpy()[:, sel], eng.robot.double().cpu().numpy()[:, sel]
    nx = nxt.double().cpu().numpy()[:, sel]
    ho = nx if eng.headed else d
    cur = np.stack([d[L.DYN_PX], d[L.DYN_PY], d[L.DYN_VX], d[L.DYN_VY], st[L.STAT_R]] + ([d[L.DYN_TH], d[L.DYN_OM]] if vis else []), -1)
    nxo = np.stack([nx[L.DYN_PX], nx[L.DYN_PY]] + ([ho[L.DYN_TH]] if vis else []) + [nx[L.DYN_VX], nx[L.DYN_VY]] + ([ho[L.DYN_OM]] if vis else []), -1)
    rob = np.stack([rb[L.ROBOT_PX], rb[L.ROBOT_PY], rb[L.ROBOT_VX], rb[L.ROBOT_VY], rb[L.ROBOT_R], rb[L.ROBOT_GX], rb[L.ROBOT_GY],
                    rb[L.ROBOT_VD], rb[L.ROBOT_TH]], -1)
    return cur, nxo, rob


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 1e-4)])
@pytest.mark.parametrize("n", [5, 25, 40])
def test_lookahead_vs_oracle_batch(dtype, tol, n):
    from social_navigation_pyenvs_b200 import CrowdEngine
    E = 96 if n <= 25 else 24  # the rejection sampler of the scenario generator is slow for dense crowds
    sc, S, R = _batch(E, n, 4000 + n)
    acts = _actions()
    for model, vis in [("hsfm_farina", False), ("hsfm_farina", True), ("sfm_guo", True)]:
        eng = CrowdEngine.from_reference_arrays(model, S, sc["goals"], consider_robot=False, all_params_equal=True, robot=R, dtype=dtype)
        eng.set_action_space(acts)
        nxt = eng.peek(0.25)
        rot, rew = eng.lookahead_from(nxt, 0.25, theta_and_omega_visible=vis)
        cur, nxo, rob = _oracle_inputs(eng, nxt, vis)
        rot_ref, rew_ref = oracle.lookahead(cur, nxo, rob, acts, 0.25, visible=vis)
        assert np.array_equal(rew.cpu().numpy(), rew_ref), (model, vis)   # flags-like output: bit-exact (double from the engine's state)
        assert rel_err(rot.double().cpu().numpy(), rot_ref).max() < tol, (model, vis)
        kinds = {(rew_ref == -0.25).any(), (rew_ref == 1.0).any(), ((rew_ref < 0) & (rew_ref > -0.25)).any(), (rew_ref == 0).any()}
        assert kinds == {True}
        # per-thread vector stores instead of the bulk asynchronous copy: same bits
        rot2, rew2 = (t.clone() for t in (rot, rew))
        rot.zero_(); rew.zero_()
        rot3, rew3 = eng.lookahead_from(nxt, 0.25, theta_and_omega_visible=vis, bulk_store=False)
        assert torch.equal(rot2, rot3) and torch.equal(rew2, rew3)


def test_onestep_lookahead_equals_the_batched_lookahead_per_action():
    """SocialNavSim.onestep_lookahead (social_nav_sim.py:1031-1049) -- one action at a time: its reward is the batched operator's
    reward for that action (same swept test, cadrl.py:56-72 vs sim:949-1029 at global time 0) and its observation the peek."""
    from social_navigation_pyenvs_b200 import CrowdEngine
    E, n = 48, 12
    sc, S, R = _batch(E, n, 31)
    eng = CrowdEngine.from_reference_arrays("hsfm_new_guo", S, sc["goals"], consider_robot=False, all_params_equal=True, robot=R)
    acts = _actions()
    eng.set_action_space(acts)
    rot, rew = eng.lookahead(0.25)
    rew = rew.cpu().numpy()
    nxt = eng.get_next_human_observable_states(0.25)
    for k in (0, 7, 33, 80):
        ob, r = eng.onestep_lookahead(np.tile(acts[k], (E, 1)), 0.25)
        assert np.array_equal(r, rew[:, k]), k
        assert np.array_equal(ob, nxt)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_lookahead_full_size_properties(dtype):
    """BASELINE size (4096 envs x 81 actions x 25 humans): the rotation is an isometry, shared columns agree, rewards take
    only the reference's values, and every output word is written (sentinel fill)."""
    from social_navigation_pyenvs_b200 import CrowdEngine
    E, N = 4096, 25
    sc, S, R = _batch(E, N, 99)
    eng = CrowdEngine.from_reference_arrays("hsfm_farina", S, sc["goals"], consider_robot=False, all_params_equal=True, robot=R, dtype=dtype)
    eng.set_action_space(_actions())
    rot, rew = eng.lookahead(0.25)
    rot.fill_(float("nan")); rew.fill_(float("nan"))
    rot, rew = eng.lookahead(0.25)
    assert not torch.isnan(rot).any() and not torch.isnan(rew).any()
    tol = 1e-12 if dtype == torch.float64 else 2e-5
    da = torch.hypot(rot[..., 6], rot[..., 7])
    assert ((da - rot[..., 11]).abs() <= tol * (1 + da)).all()                                  # |R x| = |x|
    assert (rot[..., 0] == rot[..., :1, 0]).all() and (rot[..., 2] == 0).all()                  # dg shared by the humans of an action
    assert torch.equal(rot[..., 12], rot[..., 3] + rot[..., 10])                                # radius_sum
    speed = torch.hypot(rot[..., 4], rot[..., 5])
    acts = eng.action_space.to(dtype)
    assert ((speed - acts.norm(dim=1)[None, :, None]).abs() <= tol * 2).all()
    ok = (rew == -0.25) | (rew == 1.0) | (rew == 0.0) | ((rew < 0) & (rew >= -0.2 * 0.5 * 0.25))
    assert ok.all() and (rew == -0.25).any() and (rew == 1.0).any()
    # at this size one CTA walks the whole action space of an env tile by tile (double-buffered bulk copies): oracle on a sample of envs
    sel = np.r_[0:48, 2000:2016, E - 48:E]
    cur, nxo, rob = _oracle_inputs(eng, eng._peek_buf, False, sel)
    rot_ref, rew_ref = oracle.lookahead(cur, nxo, rob, _actions(), 0.25)
    assert np.array_equal(rew[sel].cpu().numpy(), rew_ref)
    assert rel_err(rot[sel].double().cpu().numpy(), rot_ref).max() < (1e-9 if dtype == torch.float64 else 1e-4)
    rot2 = rot.clone()
    rot3, _ = eng.lookahead(0.25, bulk_store=False)
    assert torch.equal(rot2, rot3)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 1e-4)])
def test_lookahead_without_querying_the_env(dtype, tol):
    """query_env = False (crowd_nav/policy/cadrl.py:259): the humans are propagated with the constant-velocity model (:92-105)
    instead of a peek of the motion model; rewards bit-exact, rotated states within tolerance, for headed and non-headed models
    (a non-headed crowd with non-zero omega integrates its yaw here, unlike the peek)."""
    from social_navigation_pyenvs_b200 import CrowdEngine, _lib as L
    from helpers import constant_velocity_next
    E, n = 64, 9
    sc, S, R = _batch(E, n, 515)
    S[:, :, 7] = np.random.RandomState(1).uniform(-0.8, 0.8, (E, n))   # omega
    acts = _actions()
    for model, vis in [("hsfm_farina", True), ("sfm_guo", True), ("sfm_helbing", False)]:
        eng = CrowdEngine.from_reference_arrays(model, S, sc["goals"], consider_robot=False, all_params_equal=True, robot=R, dtype=dtype)
        eng.set_action_space(acts)
        before = eng.dyn.clone()
        rot, rew = eng.lookahead(0.25, theta_and_omega_visible=vis, query_env=False)
        assert torch.equal(eng.dyn, before)
        d, st, rb = eng.dyn.double().cpu().numpy(), eng.stat.double().cpu().numpy(), eng.robot.double().cpu().numpy()
        cur = np.stack([d[L.DYN_PX], d[L.DYN_PY], d[L.DYN_VX], d[L.DYN_VY], st[L.STAT_R]] + ([d[L.DYN_TH], d[L.DYN_OM]] if vis else []), -1)
        rob = np.stack([rb[L.ROBOT_PX], rb[L.ROBOT_PY], rb[L.ROBOT_VX], rb[L.ROBOT_VY], rb[L.ROBOT_R], rb[L.ROBOT_GX], rb[L.ROBOT_GY],
                        rb[L.ROBOT_VD], rb[L.ROBOT_TH]], -1)
        rot_ref, rew_ref = oracle.lookahead(cur, constant_velocity_next(cur, 0.25, vis), rob, acts, 0.25, visible=vis)
        if dtype == torch.float64:
            assert np.array_equal(rew.cpu().numpy(), rew_ref), (model, vis)
        else:  # the propagated positions are single precision here, the reference's double: decisions may differ at the margin only
            assert (rew.cpu().numpy() != rew_ref).mean() < 0.01
        assert rel_err(rot.double().cpu().numpy(), rot_ref).max() < tol, (model, vis)
