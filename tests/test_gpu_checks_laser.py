"""GPU parity of the batched collision / goal / reward reductions, the gym-level fused step and the laser kernel, against
golden vectors recorded from the live reference (tests/golden) and against the CPU oracle.  Flags, info codes, dmin / reward
values and ray hit indices must be BIT-EXACT (fp64 mode); laser ranges within 1e-12."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import OracleConfig
from helpers import GOLDEN, rel_err

pytestmark = pytest.mark.gpu
CONSTS = np.array([50, -0.25, 1.0, 0.2, 0.5, 0.25])


def _engine(model, states, goals, robot, visible, dtype=torch.float64, walls=None):
    from social_navigation_pyenvs_b200 import CrowdEngine
    S = np.concatenate([states, robot[:, None]], 1) if visible else states
    return CrowdEngine.from_reference_arrays(model, S, goals, walls=walls, consider_robot=visible, all_params_equal=True, dtype=dtype,
                                             robot=None if visible else robot)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_flags_bit_exact_vs_reference_golden(dtype):
    """400 recorded (state, action) cases: swept collision + dmin + goal (sim:949-984), reward/terminated/truncated/info
    (sim:986-1029), actual collision/dmin/goal (gym:107-118) -- all envs in one launch."""
    z = np.load(os.path.join(GOLDEN, "flags.npz"))
    H, R, A, ref = z["humans"], z["robot"], z["action"], z["result"]
    if dtype == torch.float32:  # identical inputs: fp32-representable values on both sides, flags computed in double from them
        H, R, A = (x.astype(np.float32).astype(np.float64) for x in (H, R, A))
        ref = oracle.checks(H, H.shape[1], R, A, ref[:, 11], CONSTS)
        ref = np.concatenate([ref[:, :11], z["result"][:, 11:12]], 1)
    goals = np.repeat(H[:, :, None, 10:12], 2, axis=2)
    eng = _engine("hsfm_farina", H, goals, R, visible=False, dtype=dtype)
    eng.time_now.copy_(torch.as_tensor(ref[:, 11]))
    out = eng.run_checks(A, pre=True, post=True)
    assert np.array_equal(out["collision"], ref[:, 0] != 0)
    assert np.array_equal(out["reaching_goal"], ref[:, 2] != 0)
    assert np.array_equal(out["terminated"], ref[:, 4] != 0) and np.array_equal(out["truncated"], ref[:, 5] != 0)
    assert np.array_equal(out["info"], ref[:, 6].astype(int))
    assert np.array_equal(out["actual_collision"], ref[:, 7] != 0) and np.array_equal(out["actual_goal"], ref[:, 9] != 0)
    assert np.array_equal(out["dmin"], ref[:, 1]) and np.array_equal(out["reward"], ref[:, 3]) and np.array_equal(out["actual_dmin"], ref[:, 8])
    assert out["collision"].sum() > 50 and (out["info"] == 4).sum() > 20
    # the same swept test under the reference's own name and return convention (sim:949)
    col, dmin, goal = eng.collision_detection_and_reaching_goal(A, CONSTS[5])
    assert np.array_equal(col, ref[:, 0] != 0) and np.array_equal(dmin, ref[:, 1]) and np.array_equal(goal, ref[:, 2] != 0)


def test_gym_step_sequence_vs_reference_golden():
    """SocialNavGym.step recorded for 60 steps (gym:227-250): terminated / truncated / info identical, observation (px,py,vx,vy)
    and reward within 1e-9, robot position exact -- pre-checks + 20 fused sub-steps + robot motion in ONE launch per step.
    (The discomfort reward is a function of dmin of a state the GPU has integrated for k*20 sub-steps, so it carries that
    state's 1e-15 rounding differences; bit-exactness on IDENTICAL inputs is test_flags_bit_exact_vs_reference_golden.)"""
    z = np.load(os.path.join(GOLDEN, "gym_step.npz"))
    from social_navigation_pyenvs_b200 import SFMS
    for key in ["hsfm_farina_0", "sfm_helbing_1", "hsfm_new_guo_1"]:
        visible = key.endswith("_1")
        eng = _engine(SFMS[int(z[key + "_type"])], z[key + "_states0"][None], z[key + "_goals0"][None], z[key + "_robot0"][None], visible)
        for k, a in enumerate(z[key + "_actions"]):
            eng.step(a[None], 0.0125, n_substeps=20, pre_checks=True)
            r = eng.decode_flags()
            ref = z[key + "_result"][k]
            assert r["terminated"][0] == bool(ref[1]) and r["truncated"][0] == bool(ref[2]) and r["info"][0] == int(ref[3]), (key, k)
            assert abs(r["reward"][0] - ref[0]) <= 1e-9, (key, k)
            obs = eng.get_human_states(include_goal=False, headed=False)[0]
            assert rel_err(obs, z[key + "_obs"][k][:, :4]).max() < 1e-9, (key, k)
            assert rel_err(eng.robot[:2, 0].cpu().numpy(), z[key + "_robot_pos"][k]).max() < 1e-13
        assert abs(float(eng.time_now[0]) - 15.0) < 1e-9


def test_fused_post_checks_and_touch_match_oracle_4096_envs():
    """Full-size batch (4096 x 25 + robot + walls): flags produced inside the fused launch equal the oracle's checks applied to
    the oracle's own pre-state / the kernel's own post-state (identical inputs -> bit-exact)."""
    from social_navigation_pyenvs_b200 import scenarios
    E, N, k, dt = 4096, 25, 20, 0.0125
    sc = scenarios.ccso_synthetic(E, N, seed0=7000)
    rng = np.random.RandomState(1)
    sc["robot"][:, 0:2] = sc["states"][np.arange(E), rng.randint(3, N, E), 0:2] + rng.uniform(-1.2, 1.2, (E, 2))  # robot near a human
    walls = scenarios.pack_walls(scenarios.EXAMPLE_WALLS)
    action = rng.uniform(-1, 1, (E, 2))
    eng = _engine("hsfm_farina", sc["states"], sc["goals"], sc["robot"], visible=True, walls=walls)
    states = np.concatenate([sc["states"], sc["robot"][:, None]], 1)
    pre = oracle.checks(states, N, sc["robot"], action, np.zeros(E), CONSTS)
    eng.step(action, dt, n_substeps=k, pre_checks=True, post_checks=True, track_touch=True)
    r = eng.decode_flags()
    assert np.array_equal(r["collision"], pre[:, 0] != 0) and np.array_equal(r["dmin"], pre[:, 1]) and np.array_equal(r["reward"], pre[:, 3])
    assert np.array_equal(r["info"], pre[:, 6].astype(int))
    post_rows = eng.rows(states)
    rob = sc["robot"].copy()
    rob[:, 0:2] = eng.robot[:2].cpu().numpy().T
    post = oracle.checks(post_rows, N, rob, action, np.zeros(E), CONSTS)
    assert np.array_equal(r["actual_collision"], post[:, 7] != 0) and np.array_equal(r["actual_dmin"], post[:, 8])
    assert r["collision"].sum() > 100 and r["actual_collision"].sum() > 10
    # touched (any sub-step) is implied by a post-step overlap
    assert np.all(r["touched"][post[:, 10] != 0])


def _grazing(humans, walls, origin, angle, hit_pair, range_pair, maxd, eps=2e-3):
    """A single-precision scan may disagree with the reference about WHICH entity a ray hits only where the geometry is marginal:
    the ray passes within `eps` metres of the rim of a circle / the end of a segment it hits in one result and misses in the other,
    or the two results name different entities at (nearly) the same range."""
    if abs(range_pair[0] - range_pair[1]) <= eps:
        return True
    o, d = np.asarray(origin, np.float64), np.array([np.cos(angle), np.sin(angle)])
    n = humans.shape[0]
    segs = walls.reshape(-1, 2, 2)
    for h in set(hit_pair):
        if h < 0:
            continue
        if h < n:  # circle: distance of the centre from the ray's line vs the radius, or a hit next to the range limit
            c = humans[h, :2] - o
            if abs(abs(c[0] * d[1] - c[1] * d[0]) - humans[h, 2]) < eps or abs(np.hypot(*c) - humans[h, 2] - maxd) < eps:
                return True
        else:      # segment: the crossing point sits within eps of one of its ends, or the ray is nearly parallel to it
            a, b = segs[h - n]
            e = b - a
            den = d[0] * e[1] - d[1] * e[0]
            if abs(den) < 1e-6:
                return True
            u = ((a[0] - o[0]) * d[1] - (a[1] - o[1]) * d[0]) / den
            if min(abs(u), abs(u - 1)) * np.hypot(*e) < eps:
                return True
    return False


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_laser_vs_reference_golden(dtype):
    from social_navigation_pyenvs_b200 import sensors
    z = np.load(os.path.join(GOLDEN, "laser.npz"))
    keys = sorted(k[:-5] for k in z.files if k.endswith("_pose"))
    for key in keys:
        x, y, yaw, rng, samples, maxd = z[key + "_pose"]
        ranges, hits = sensors.scan_batch(z[key + "_humans"][None], z[key + "_walls"], np.array([[x, y, yaw]]), rng, int(samples), maxd, dtype=dtype)
        if dtype == "float64":
            assert np.array_equal(hits[0], z[key + "_hits"]), key
            assert np.abs(ranges[0] - z[key + "_ranges"]).max() <= 1e-12, key
        else:  # fp32: ranges to 1e-4 relative wherever both agree on the hit entity; a ray may only name another entity when it GRAZES
            same = hits[0] == z[key + "_hits"]
            assert (np.abs(ranges[0] - z[key + "_ranges"])[same] <= 1e-4 * np.maximum(z[key + "_ranges"][same], 1)).all(), key
            ang = yaw - rng / 2 + np.arange(int(samples)) * (rng / (int(samples) - 1)) if int(samples) > 1 else np.array([yaw])
            for r in np.nonzero(~same)[0]:
                assert _grazing(z[key + "_humans"], z[key + "_walls"], (x, y), z[key + "_angles"][r] if key + "_angles" in z.files else ang[r],
                                (int(hits[0][r]), int(z[key + "_hits"][r])), (float(ranges[0][r]), float(z[key + "_ranges"][r])), maxd), (key, int(r))
            assert same.mean() > 0.9, (key, same.mean())


def test_laser_class_matches_reference_dict():
    from social_navigation_pyenvs_b200.sensors import LaserSensor
    z = np.load(os.path.join(GOLDEN, "laser.npz"))
    x, y, yaw, rng, samples, maxd = z["dense_0_pose"]
    s = LaserSensor(np.array([x, y]), yaw, rng, int(samples), maxd, uncertainty=None)
    m = s.get_laser_measurements(z["dense_0_humans"], z["dense_0_walls"])
    assert np.array_equal(np.array(list(m.keys())), z["dense_0_angles"])
    assert np.abs(np.array(list(m.values())) - z["dense_0_ranges"]).max() <= 1e-12
    with pytest.raises(ValueError):
        LaserSensor(np.zeros(2), 4.0, rng, 10, 10.0)
    with pytest.raises(ValueError):
        LaserSensor(np.zeros(2), 0.0, rng, 10, 11.0)


def test_laser_4096_envs_360_rays_vs_oracle_and_warp_kernel():
    """BASELINE config 4 at full size: 4096 envs x 360 rays over 25 humans + 14 wall segments; hit indices bit-exact vs the
    oracle; the warp-per-ray kernel (used for large crowds) must agree with the thread-per-ray kernel."""
    from social_navigation_pyenvs_b200 import scenarios, sensors
    E, N = 4096, 25
    sc = scenarios.ccso_synthetic(E, N, seed0=2000)
    humans = sc["states"][:, :, [0, 1, 8]]
    walls = scenarios.pack_walls(scenarios.EXAMPLE_WALLS)
    pose = np.concatenate([sc["robot"][:, 0:2], np.full((E, 1), np.pi / 2)], 1)
    ranges, hits = sensors.scan_batch(humans, walls, pose, 2 * np.pi, 360, 10.0)
    r_ref, h_ref = oracle.laser(humans[:256], walls, pose[:256], 2 * np.pi, 360, 10.0)
    # ranges: CUDA's sincos and glibc's differ in the last ulp of the ray direction -> 1e-13 relative on a 10 m range
    assert np.array_equal(hits[:256], h_ref) and np.abs(ranges[:256] - r_ref).max() <= 1e-11
    assert (hits >= 0).mean() > 0.2
    # one crowd of 600 circles -> warp-per-ray path; compare with the oracle
    big = np.concatenate([humans[:24].reshape(1, -1, 3)], 1)
    big[0, :, 0:2] += np.repeat(np.arange(24)[:, None] * 0.37, N, 0)
    r2, h2 = sensors.scan_batch(big, walls, pose[:1], 2 * np.pi, 720, 10.0)
    r2_ref, h2_ref = oracle.laser(big, walls, pose[:1], 2 * np.pi, 720, 10.0)
    assert np.array_equal(h2, h2_ref) and np.abs(r2 - r2_ref).max() <= 1e-11


def test_peek_next_observable_states_vs_reference_golden():
    """get_next_human_observable_states (mmm:691-709): dt=0.25 peek, pose/velocity/goal restored, desired force not."""
    from social_navigation_pyenvs_b200 import CrowdEngine
    z = np.load(os.path.join(GOLDEN, "peek.npz"))
    for model in ["sfm_helbing", "hsfm_farina", "hsfm_new_guo"]:
        S = np.concatenate([z[model + "_states"], z[model + "_robot"][None]], 0)[None]
        eng = CrowdEngine.from_reference_arrays(model, S, z[model + "_goals"][None], consider_robot=True, all_params_equal=True)
        eng.set_desired_force(z[model + "_desired"][None])
        before = eng.get_human_states(include_goal=True, headed=eng.headed)
        obs4 = eng.get_next_human_observable_states(0.25)
        assert rel_err(obs4[0], z[model + "_obs4"]).max() < 1e-9
        assert np.array_equal(eng.get_human_states(include_goal=True, headed=eng.headed), before)
        obs8 = eng.get_next_human_observable_states(0.25, theta_and_omega_visible=True)
        assert rel_err(obs8[0], z[model + "_obs8"]).max() < 1e-9
        assert rel_err(eng.desired_force()[0], z[model + "_after"][:, 10:12], scale=100.0).max() < 1e-9


def test_robot_push_out_vs_reference_golden_and_oracle():
    """RobotAgent.check_collisions (robot_agent.py:35-48) on the device: one env per recorded start position, bit-exact
    against the live reference's result; a seeded 512-env batch with per-env humans against the oracle; fp32 state to 1e-6."""
    from social_navigation_pyenvs_b200 import CrowdEngine, scenarios
    z = np.load(os.path.join(GOLDEN, "push_out.npz"))
    for name in ("walls", "cc"):
        H, st, en, r = z[name + "_humans"], z[name + "_start"], z[name + "_end"], float(z[name + "_radius"])
        E, n = len(st), len(H)
        rows = np.zeros((E, n, 13))
        rows[:, :, 0:2], rows[:, :, 8], rows[:, :, 9] = H[None, :, 0:2], H[None, :, 2], 75.0
        robot = np.zeros((E, 13))
        robot[:, 0:2], robot[:, 8], robot[:, 9] = st, r, 80.0
        walls = z[name + "_walls"]
        for dtype, tol in [(torch.float64, 0.0), (torch.float32, 2e-6)]:
            eng = CrowdEngine.from_reference_arrays("sfm_helbing", rows, np.zeros((E, n, 1, 2)), walls=walls if walls.size else None,
                                                    consider_robot=False, robot=robot, dtype=dtype)
            eng.robot_check_collisions()
            got = eng.robot_rows()[0][:, 0:2]
            assert np.abs(got - en).max() <= tol, (name, dtype)
    E, n = 512, 25
    sc = scenarios.ccso_synthetic(E, n, 77)
    rng = np.random.RandomState(3)
    robot = sc["robot"].copy()
    robot[:, 0:2] = sc["states"][np.arange(E), rng.randint(n, size=E), 0:2] + rng.uniform(-0.6, 0.6, (E, 2))
    walls = scenarios.pack_walls(scenarios.EXAMPLE_WALLS)
    eng = CrowdEngine.from_reference_arrays("hsfm_farina", sc["states"], sc["goals"], walls=walls, consider_robot=False, robot=robot)
    eng.robot_check_collisions()
    ref = oracle.robot_push_out(sc["states"][:, :, [0, 1, 8]], walls, robot[:, [0, 1, 8]])
    got = eng.robot_rows()[0][:, 0:2]
    assert np.array_equal(got, ref) and (np.abs(ref - robot[:, 0:2]).sum(1) > 0).sum() > 100


def test_batched_laser_noise_has_the_reference_distribution():
    """LaserSensor.add_uncertainty (sensors.py:71-74) for batched scans: clip(N(range, sigma), 0, max_distance), drawn on the device
    from a counter-based Philox stream (the reference uses the caller's global np.random stream, so parity is distributional):
    residuals are standard normal (moments + Kolmogorov-Smirnov), uncorrelated between neighbouring rays and envs, reproducible for
    a (seed, scan number) and fresh for the next scan; clipping as in the reference."""
    from scipy import stats
    from social_navigation_pyenvs_b200 import CrowdEngine, scenarios, sensors, _lib as L
    E, N, samples, sigma, maxd, rr = 256, 6, 360, 0.05, 10.0, 0.3
    sc = scenarios.circular_crossing(E, N, seed0=5)
    states = np.concatenate([sc["states"], sc["robot"][:, None]], 1)
    eng = CrowdEngine.from_reference_arrays("sfm_helbing", states, sc["goals"], walls=scenarios.pack_walls(scenarios.EXAMPLE_WALLS), consider_robot=True)
    pose = torch.stack([eng.robot[L.ROBOT_PX], eng.robot[L.ROBOT_PY], torch.full((E,), 1.0, dtype=torch.float64, device="cuda")]).contiguous()
    clean = sensors.EngineScanner(eng, 2 * np.pi, samples, maxd, robot_radius=rr).scan(pose)[0].cpu().numpy().copy()
    noisy_scanner = sensors.EngineScanner(eng, 2 * np.pi, samples, maxd, robot_radius=rr, uncertainty=sigma, seed=1234)
    r0 = noisy_scanner.scan(pose)[0].cpu().numpy().copy()
    r1 = noisy_scanner.scan(pose)[0].cpu().numpy().copy()
    again = sensors.EngineScanner(eng, 2 * np.pi, samples, maxd, robot_radius=rr, uncertainty=sigma, seed=1234).scan(pose)[0].cpu().numpy()
    other = sensors.EngineScanner(eng, 2 * np.pi, samples, maxd, robot_radius=rr, uncertainty=sigma, seed=1235).scan(pose)[0].cpu().numpy()
    assert np.array_equal(r0, again) and not np.array_equal(r0, r1) and not np.array_equal(r0, other)
    inner = (clean + rr > 10 * sigma) & (clean + rr < maxd - 10 * sigma)          # rays whose noise cannot reach a clipping bound
    assert inner.sum() > 20000
    z = ((r0 - clean) / sigma)[inner]
    n = z.size
    assert abs(z.mean()) < 5 / np.sqrt(n) and abs(z.var() - 1) < 0.05 and abs(stats.skew(z)) < 0.08 and abs(stats.kurtosis(z)) < 0.15
    assert stats.kstest(z, "norm").pvalue > 1e-3
    both = inner[:, 1:] & inner[:, :-1]
    za = (r0 - clean) / sigma
    assert abs(np.corrcoef(za[:, 1:][both], za[:, :-1][both])[0, 1]) < 0.02       # neighbouring rays
    z2 = ((r1 - clean) / sigma)[inner]
    assert abs(np.corrcoef(z, z2)[0, 1]) < 0.02                                  # successive scans
    # clipping (sensors.py:73): a miss reads max_distance and can only come back lower; nothing leaves [0, max_distance]
    miss = clean + rr == maxd
    assert miss.any() and (r0[miss] + rr <= maxd).all() and ((r0[miss] + rr == maxd).mean() > 0.4)
    assert (r0 + rr >= 0).all() and (r0 + rr <= maxd).all()


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_scan_host_writes_pinned_buffers_directly(dtype):
    from social_navigation_pyenvs_b200 import CrowdEngine, scenarios, sensors, _lib as L
    E, N, samples = 300, 9, 181
    sc = scenarios.circular_crossing(E, N, seed0=15)
    states = np.concatenate([sc["states"], sc["robot"][:, None]], 1)
    eng = CrowdEngine.from_reference_arrays("sfm_helbing", states, sc["goals"], walls=scenarios.pack_walls(scenarios.EXAMPLE_WALLS), consider_robot=True, dtype=dtype)
    pose = torch.stack([eng.robot[L.ROBOT_PX], eng.robot[L.ROBOT_PY], torch.full((E,), -0.4, dtype=dtype, device="cuda")]).contiguous()
    scanner = sensors.EngineScanner(eng, np.pi, samples, 8.0, robot_radius=0.3)
    ranges, hits = (t.clone() for t in scanner.scan(pose))
    rh = torch.full((E, samples), -1.0, dtype=dtype).pin_memory()
    hh = torch.full((E, samples), -7, dtype=torch.int32).pin_memory()
    scanner.scan_host(pose.cpu().pin_memory(), rh, hh)
    assert torch.equal(rh, ranges.cpu()) and torch.equal(hh, hits.cpu())
    with pytest.raises(ValueError):
        scanner.scan_host(pose.cpu(), rh, hh)   # pageable pose
