#!/bin/bash
# Dev tool (run under gpurun): GPU test suite, then the fused-step bench lines of HEAD (and of any tuning variants given as args).
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -40 > $O/gpu_tests.log
tail -5 $O/gpu_tests.log
for dt in f64 f32; do
for v in default "$@"; do
  if [ $v = default ]; then unset SNP_B200_LIB; else export SNP_B200_LIB=$PWD/social_navigation_pyenvs_b200/variants/$v/libsnp_b200.so; fi
  echo -n "$dt $v: "; python bench.py --steps 100 --warmup 5 --dtype $dt --no-cpu-baseline 2>$O/bench_err_${v}_${dt}.log | tail -1 | tee $O/bench_${v}_${dt}.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d.get('e2e',{}).get('value'))"
done
done
unset SNP_B200_LIB
