#!/bin/bash
# Round profile recipe (run under gpurun on ONE GPU): launch list of the default bench command, then one full capture of each
# dominant kernel.  Outputs land in gpurun_out/ (scratch); summaries are copied into profiles/ by tools/summarise_profiles.py.
R=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 3 -c 1 -o gpurun_out/${R}_k_step_f64 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 3 -c 1 -o gpurun_out/${R}_k_step_f32 python bench.py --steps 3 --warmup 3 --dtype f32 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_laser -s 3 -c 1 -o gpurun_out/${R}_k_laser_f64 python bench.py --workload laser_4096x360 --steps 3 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_large_pairs -s 6 -c 1 -o gpurun_out/${R}_k_large_pairs_f64 python bench.py --workload 65536_hsfm_single_crowd --steps 3 --warmup 3 > /dev/null 2>&1
ls -la gpurun_out | grep ${R}
