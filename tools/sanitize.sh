#!/bin/bash
# compute-sanitizer pass over every kernel family (run under gpurun): memcheck + racecheck + synccheck on small inputs.
set -o pipefail
S=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  echo "=== $tool: smoke (fused step, checks, walls, robot)"; $S --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
  echo "=== $tool: adapters / respawn / robot model / laser / large crowd"
  $S --tool $tool --error-exitcode 9 python -m pytest -q -m gpu -x tests/test_gpu_adapters.py "tests/test_gpu_step_parity.py::test_parallel_traffic_with_respawn_full_trajectory[pt7_sfm_helbing]" "tests/test_gpu_checks_laser.py::test_laser_class_matches_reference_dict" "tests/test_gpu_checks_laser.py::test_flags_bit_exact_vs_reference_golden" "tests/test_gpu_sizes_large.py::test_large_crowd_tiled_kernel_vs_oracle" "tests/test_gpu_sizes_large.py::test_per_env_walls_and_per_agent_params" "tests/test_gpu_sizes_large.py::test_warp_packed_and_block_packed_mappings_agree" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|error" | tail -6
  echo "=== $tool: round 2 -- unicycle robot, safety space, zero-copy gym step, laser noise + host scan, constant-velocity lookahead, fused large-crowd run"
  $S --tool $tool --error-exitcode 9 python -m pytest -q -m gpu -x "tests/test_gpu_adapters.py::test_unicycle_robot_kinematics" "tests/test_gpu_adapters.py::test_gym_step_through_the_host_buffer_c_abi_call" "tests/test_gpu_checks_laser.py::test_scan_host_writes_pinned_buffers_directly" "tests/test_gpu_lookahead.py::test_lookahead_without_querying_the_env" "tests/test_gpu_sizes_large.py::test_large_crowd_chunked_sums_and_exact_culling" "tests/test_gpu_math.py" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|error" | tail -6
  echo "=== $tool: lookahead (bulk asynchronous tile stores, peek) / on-device reset"
  $S --tool $tool --error-exitcode 9 python -m pytest -q -m gpu -x "tests/test_gpu_lookahead.py::test_lookahead_vs_reference_golden" "tests/test_gpu_lookahead.py::test_lookahead_vs_oracle_batch[5-dtype0-1e-09]" "tests/test_gpu_lookahead.py::test_lookahead_vs_oracle_batch[25-dtype1-0.0001]" "tests/test_gpu_reset.py::test_device_reset_vs_reference_golden" "tests/test_gpu_reset.py::test_hybrid_scenario_coin_respawn_and_masked_reset" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|error" | tail -6
done
