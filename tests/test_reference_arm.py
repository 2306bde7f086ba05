"""CPU tests of the reference arm (bench.py --impl reference, cpu_baseline): the live reference staged under oracle/_ref (or at
/root/reference) stepped by a pool of resident processes, one env each, and the JSON contract of the line bench.py prints."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not reference.available(), reason="the live reference is neither at /root/reference nor staged under oracle/_ref")


def test_pool_steps_one_reference_env_per_process_and_matches_the_oracle():
    import oracle
    from oracle import OracleConfig
    from social_navigation_pyenvs_b200 import scenarios
    E, N = 2, 6
    sc = scenarios.circular_crossing(E, N, seed0=31)
    states = np.concatenate([sc["states"], sc["robot"][:, None]], 1)
    pool = reference.ReferencePool(2, "hsfm_farina", states, sc["goals"], scenarios.EXAMPLE_WALLS, None, True, 0.0125)
    try:
        wall, inner = pool.step(20)
        assert len(inner) == 2 and wall > 0 and all(t > 0 for t in inner)
    finally:
        pool.close()
    # the same builder in-process: 12 updates (2 warm-up + 10) agree with the oracle on the same crowd
    sim = reference.sim_from_arrays("hsfm_farina", states[0, :N], sc["goals"][0], scenarios.EXAMPLE_WALLS, states[0, N], True, 0.0125)
    reference.step_like_gym(sim, (0.0, 1.0), 0.0125, 12)
    cfg = OracleConfig(oracle.type_code("hsfm_farina"), True, bool(sim.motion_model_manager.all_equal_humans), False)
    params = np.array([h.get_parameters("hsfm_farina") for h in sim.humans])[None]
    ref, _, _ = oracle.update_humans(cfg, states[:1], sc["goals"][:1], scenarios.pack_walls(scenarios.EXAMPLE_WALLS), params, np.zeros((1, N + 1)),
                                     np.zeros((1, N, 2)), 0.0125, 12, robot_vel=np.array([[0.0, 1.0]]))
    got = reference.human_rows(sim)
    assert np.abs(got[:, :8] - ref[0, :N, :8]).max() < 1e-9


def test_bench_reference_arm_prints_the_contract_line():
    env = dict(os.environ, SNP_BENCH_ENVS="64")   # a small batch keeps the scenario generation of this CPU test short
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "agent-steps/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert set(d["config"]) >= {"workload", "envs_per_gpu", "humans", "motion_model", "substeps_per_step", "dt", "walls", "robot_visible"}
