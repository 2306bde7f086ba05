"""Fold freshly measured bench lines into profiles/<round>_bench_lines.jsonl: a new line replaces the old one of the same
(workload, dtype, impl); everything else stays.  usage: merge_lines.py r02 gpurun_out/r02_large_lines.jsonl"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R, src = sys.argv[1], sys.argv[2]
path = os.path.join(ROOT, "profiles", f"{R}_bench_lines.jsonl")
key = lambda d: (d.get("config", {}).get("workload"), d.get("dtype"), d.get("impl", "b200"))
new = {key(json.loads(l)): l.strip() for l in open(src) if l.strip()}
out, seen = [], set()
for l in open(path):
    if not l.strip():
        continue
    k = key(json.loads(l))
    out.append(new.get(k, l.strip()))
    seen.add(k)
out += [l for k, l in new.items() if k not in seen]
open(path, "w").write("\n".join(out) + "\n")
print(f"{len(new)} line(s) merged, {len(out)} in {path}")
