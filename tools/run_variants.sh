#!/bin/bash
# Dev tool: bench the tuning variants built by build.build_variant (ms per fused launch).  usage: run_variants.sh tag1 tag2 ...
for dt in f64 f32; do
for v in default "$@"; do
  if [ $v = default ]; then unset SNP_B200_LIB; else export SNP_B200_LIB=$PWD/social_navigation_pyenvs_b200/variants/$v/libsnp_b200.so; fi
  echo -n "$dt $v: "; python bench.py --steps 100 --warmup 3 --dtype $dt --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks']['sm_mhz'])"
done
done
