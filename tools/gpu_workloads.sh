#!/bin/bash
# Dev tool (run under gpurun): ms per launch of the fused-step workloads, both dtypes.
for w in 4096x25_hsfm_ccso_walls_robot 4096x5_sfm_helbing_cc 32768x5_sfm_helbing_cc 4096x25_hsfm_ccso_robot; do
for dt in f64 f32; do
  echo -n "$w $dt: "; SNP_BENCH_NO_LARGE=1 python bench.py --workload $w --steps 100 --warmup 5 --dtype $dt --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],5), '%.3e' % d['value'], 'frac', round(d['roofline']['frac'],4), 'e2e %.3e' % d['e2e']['value'])"
done; done
