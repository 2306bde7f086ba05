#!/bin/bash
# Dev tool (run under gpurun): ms per fused launch of tuning variants.  usage: gpu_variants.sh "<dtypes>" "<env counts>" tag1 tag2 ...
O=gpurun_out; mkdir -p $O
dts=$1; envs=$2; shift 2
for dt in $dts; do
for e in $envs; do
for v in default "$@"; do
  if [ $v = default ]; then unset SNP_B200_LIB; else export SNP_B200_LIB=$PWD/social_navigation_pyenvs_b200/variants/$v/libsnp_b200.so; fi
  echo -n "$dt E=$e $v: "; SNP_BENCH_ENVS=$e python bench.py --steps 60 --warmup 5 --dtype $dt --no-cpu-baseline 2>$O/verr_${v}_${dt}.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), '%.3e' % d['value'], d['clocks']['sm_mhz'])"
done
done
done
unset SNP_B200_LIB
