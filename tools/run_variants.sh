#!/bin/bash
# Dev tool: bench the occupancy variants built by build.build_variant (ms per fused launch).
for v in default f64_5 f64_6; do
  if [ $v = default ]; then unset SNP_B200_LIB; else export SNP_B200_LIB=$PWD/social_navigation_pyenvs_b200/variants/$v/libsnp_b200.so; fi
  echo -n "f64 $v: "; python bench.py --steps 100 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks'])"
done
for v in default f32_6 f32_5; do
  if [ $v = default ]; then unset SNP_B200_LIB; else export SNP_B200_LIB=$PWD/social_navigation_pyenvs_b200/variants/$v/libsnp_b200.so; fi
  echo -n "f32 $v: "; python bench.py --steps 100 --warmup 3 --dtype f32 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks'])"
done
