"""Dev tool: issue-port cost model of a captured kernel from its per-SASS-instruction page (gpurun_out/<round>_<name>.source.csv).
cost = sum over instructions of executed warp-instructions x issue cycles (FP64 arithmetic: 2.0 cycles of the SMSP's issue port,
measured by tools/pipe_microbench.cu -- nothing else issues in its shadow; everything else: 1).  Prints the bound this gives next
to the measured kernel time, and the cost per straight-line region (consecutive instructions with the same execution count)."""
import csv
import sys

import json
import os

src, raw = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None
key = sys.argv[3] if len(sys.argv) > 3 else None  # "<workload>:<dtype>": also record the counts in profiles/issue_model.json
FP64 = {"DFMA", "DMUL", "DADD", "DSETP"}
rows = list(csv.reader(open(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
base = int(data[0][ix["Address"]], 16)
L, tot, n64, nother = [], 0.0, 0, 0
for r in data:
    if len(r) < len(hdr):
        continue
    t = r[ix["Source"]].split()
    op = (t[1] if t[0].startswith("@") else t[0]).rstrip(";")
    n = int(float(r[ix["Instructions Executed"]] or 0))
    f = op.split(".")[0] in FP64
    n64 += n if f else 0
    nother += 0 if f else n
    c = n * (2.0 if f else 1.0)
    L.append((int(r[ix["Address"]], 16) - base, op, n, c, f))
    tot += c
smsps = 148 * 4
print(f"warp instructions: fp64 {n64 / 1e6:.1f} M, other {nother / 1e6:.1f} M; issue cost {tot / 1e6:.1f} M cycles = {tot / smsps / 1e3:.1f} K cycles per SMSP")
if raw:
    rr = list(csv.reader(open(raw)))
    m = dict(zip(rr[0], rr[2]))
    cyc = float(m["sm__cycles_elapsed.max"].replace(",", ""))
    print(f"measured {cyc / 1e3:.1f} K cycles elapsed -> the kernel runs at {100 * tot / smsps / cyc:.1f} % of its issue-port bound")
if key:
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "issue_model.json")
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[key] = {"fp64_warp_instructions": n64, "other_warp_instructions": nother, "fp64_issue_cycles": 2.0, "smsps": smsps,
              "issue_cycles_per_smsp": tot / smsps, "source": os.path.basename(src)}
    d["_source"] = ("per-SASS-instruction execution counts of one launch (ncu --set full, source page); FP64 arithmetic costs 2.0 issue cycles "
                    "of its SMSP and nothing issues in its shadow (profiles/r01_pipe_microbench.txt)")
    json.dump(d, open(path, "w"), indent=1)
seg, cur = [], None
for a, op, n, c, f in L:
    if cur and cur[2] == n:
        cur[1] = a; cur[3] += c; cur[4] += 1; cur[5] += 1 if f else 0
    else:
        cur = [a, a, n, c, 1, 1 if f else 0]
        seg.append(cur)
for s in seg:
    if s[3] / tot > 0.004:
        print(f"0x{s[0]:05x}-0x{s[1]:05x} executed {s[2]:9d} x {s[4]:4d} instr ({s[5]:3d} fp64)  {100 * s[3] / tot:5.1f} % of the issue cost")
