"""Dev tool: opcode mix and stall summary from `ncu --page source --print-source sass --csv` output."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); stalls = collections.Counter(); samples = collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    n = int(float(r[ix["Instructions Executed"]] or 0))
    toks = src.split()
    op = toks[0] if not toks[0].startswith("@") else toks[1]
    op = op.rstrip(";")
    base = op.split(".")[0]
    ops[base] += n; tot += n
    samples[base] += int(float(r[ix["# Samples"]] or 0))
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stalls[h] += int(float(r[ix[h]] or 0))
print("total warp instructions", tot)
for k, v in ops.most_common(30):
    print(f"{k:12s} {v:12d} {100*v/tot:5.1f}%   samples {samples[k]}")
print("stalls:", [(k, v) for k, v in stalls.most_common(8)])
