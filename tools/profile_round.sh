#!/bin/bash
# Round profile recipe (run under gpurun on ONE GPU): launch list of the default bench command, then one full capture of each
# dominant kernel.  The .ncu-rep files embed the whole module (tens of MB for k_step) and gpurun only brings 64 MiB back, so the
# raw-metrics and per-SASS-instruction pages are exported to CSV on the box and the large reports dropped.
# Outputs land in gpurun_out/ (scratch); summaries are copied into profiles/ by tools/summarise_profiles.py.
R=${1:-r02}
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/${R}_launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/${R}_bench_under_ncu.log 2>&1
cap() {  # name kernel-regex skip bench-args...
  name=$1; k=$2; skip=$3; shift 3
  SNP_BENCH_NO_LARGE=1 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o $O/${R}_${name} python bench.py "$@" > /dev/null 2>&1
  ncu -i $O/${R}_${name}.ncu-rep --page raw --csv > $O/${R}_${name}.raw.csv 2>/dev/null
  ncu -i $O/${R}_${name}.ncu-rep --page source --csv --print-source sass > $O/${R}_${name}.source.csv 2>/dev/null
  if [ $(stat -c %s $O/${R}_${name}.ncu-rep) -gt 8000000 ]; then rm -f $O/${R}_${name}.ncu-rep; fi
}
cap k_step_f64 k_step 3 --steps 3 --warmup 3 --no-cpu-baseline
cap k_step_f32 k_step 3 --steps 3 --warmup 3 --dtype f32 --no-cpu-baseline
cap k_laser_f64 k_laser 3 --workload laser_4096x360 --steps 3 --warmup 3
cap k_large_pairs_f64 k_large_pairs 6 --workload 65536_hsfm_single_crowd --steps 3 --warmup 3
cap k_step_5_f64 k_step 3 --steps 3 --warmup 3 --workload 4096x5_sfm_helbing_cc --no-cpu-baseline
cap k_lookahead_f64 k_lookahead 3 --workload lookahead_4096x81x25 --steps 3 --warmup 3
cap k_lookahead_f32 k_lookahead 3 --workload lookahead_4096x81x25 --dtype f32 --steps 3 --warmup 3
ls -la $O | grep ${R}
