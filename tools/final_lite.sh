#!/bin/bash
# Last check of a round when GPU minutes are short (under gpurun, ONE GPU): GPU tests, smoke, the four bench lines that carry
# large-crowd figures.  No ncu, no sanitizer.
R=${1:-r02}
O=gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -3 > $O/final_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/final_smoke.log 2>&1
: > $O/${R}_large_lines.jsonl
python bench.py 2>$O/bench_err_default.log | tail -1 >> $O/${R}_large_lines.jsonl
python bench.py --dtype f32 --no-cpu-baseline 2>/dev/null | tail -1 >> $O/${R}_large_lines.jsonl
for dt in f64 f32; do
  python bench.py --workload 65536_hsfm_single_crowd --dtype $dt --no-cpu-baseline 2>/dev/null | tail -1 >> $O/${R}_large_lines.jsonl
done
cat $O/final_gpu_tests.log $O/final_smoke.log; wc -l $O/${R}_large_lines.jsonl
