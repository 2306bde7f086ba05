#!/bin/bash
# Round-end evidence run (under gpurun, ONE GPU): GPU tests, smoke, the bench line of every workload x dtype and the reference arm.
R=${1:-r02}
O=gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -3 > $O/final_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/final_smoke.log 2>&1
: > $O/${R}_bench_lines.jsonl
for dt in f64 f32; do
for w in 4096x25_hsfm_ccso_walls_robot 4096x25_hsfm_ccso_robot 4096x5_sfm_helbing_cc 32768x5_sfm_helbing_cc laser_4096x360 lookahead_4096x81x25 65536_hsfm_single_crowd; do
  extra=--no-cpu-baseline; if [ $w = 4096x25_hsfm_ccso_walls_robot ] && [ $dt = f64 ]; then extra=; fi
  python bench.py --workload $w --dtype $dt $extra 2>$O/bench_err_${w}_${dt}.log | tail -1 >> $O/${R}_bench_lines.jsonl
done
done
python bench.py --impl reference --steps 20 --warmup 3 2>/dev/null | tail -1 >> $O/${R}_bench_lines.jsonl
cat $O/final_gpu_tests.log $O/final_smoke.log; wc -l $O/${R}_bench_lines.jsonl
