#!/bin/bash
# Dev tool (run under gpurun, ONE GPU): one full ncu capture of the fused step kernel; CSV pages exported on the box.
# usage: prof_kstep.sh <name> <dtype> [extra bench args]
name=$1; dt=$2; shift 2
O=gpurun_out; mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:k_step -s 3 -c 1 -f -o $O/${name} python bench.py --steps 3 --warmup 3 --dtype $dt --no-cpu-baseline "$@" > /dev/null 2>&1
ncu -i $O/${name}.ncu-rep --page raw --csv > $O/${name}.raw.csv 2>/dev/null
ncu -i $O/${name}.ncu-rep --page source --csv --print-source sass > $O/${name}.source.csv 2>/dev/null
rm -f $O/${name}.ncu-rep
ls -la $O/${name}.*
