"""Dev tool: pipe / issue cost model of a captured kernel from its per-SASS-instruction page (gpurun_out/<round>_<name>.source.csv).

Round-2 model, from tools/pipe_microbench.cu on B200 (profiles/r02_pipe_microbench.txt), per SM sub-partition (SMSP):
  * FP64 pipe, shared dispatch with the FMA pipe: a DFMA whose three sources are distinct registers occupies it for 3.0 cycles,
    DMUL / DADD / DSETP for 2.06; FMA-pipe instructions (IMAD*, FFMA, FMUL, FADD, ...) do NOT issue in their shadow (8 DFMA + 16
    IMAD take the sum of both), so they are charged on the same port at 1 cycle each;
  * ALU pipe (LOP3, SEL / FSEL, ISETP, VIMNMX, IADD3 / VIADD, SHF, LEA, MOV, PRMT ...): 16 lanes, 2 cycles per warp instruction, and it
    DOES overlap with FP64 (8 DFMA + 16 LOP3: 155 cycles against 87 + 128 separately);
  * one instruction issued per cycle.
bound = max(FP64/FMA port, ALU pipe, issue slots).  Prints the bound next to the measured kernel time, and the cost per
straight-line region (consecutive instructions with the same execution count).
    python tools/issue_cost.py <source.csv> [<raw.csv> [<workload>:<dtype>]]   (the key records the counts in profiles/issue_model.json)
"""
import csv
import json
import os
import sys

src, raw = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None
key = sys.argv[3] if len(sys.argv) > 3 else None
FP64_3 = {"DFMA"}
FP64_2 = {"DMUL", "DADD", "DSETP", "DMNMX"}
FMA_PIPE = {"IMAD", "FFMA", "FMUL", "FADD", "FFMA2", "FMUL2", "FADD2", "HFMA2", "HMUL2", "HADD2", "IDP", "IMAD.MOV", "IMAD.SHL", "IMAD.IADD", "IMAD.WIDE", "IMAD.X", "IMAD.U32", "IMAD.HI"}
ALU_PIPE = {"LOP3", "SEL", "FSEL", "ISETP", "VIMNMX", "VIMNMX3", "VIADDMNMX", "IADD3", "VIADD", "SHF", "LEA", "MOV", "PRMT", "FSETP", "FMNMX", "IABS", "FLO",
            "BREV", "POPC", "PLOP3", "IMNMX", "SGXT", "BMSK", "FCHK", "CS2R", "P2R", "R2P", "LOP", "SHL", "SHR"}
rows = list(csv.reader(open(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
base = int(data[0][ix["Address"]], 16)
L = []
n = {"dfma": 0, "fp64_other": 0, "fma_pipe": 0, "alu": 0, "rest": 0}
for r in data:
    if len(r) < len(hdr):
        continue
    t = r[ix["Source"]].split()
    op = (t[1] if t[0].startswith("@") else t[0]).rstrip(";")
    b = op.split(".")[0]
    cnt = int(float(r[ix["Instructions Executed"]] or 0))
    if b in FP64_3:
        cls = "dfma"
    elif b in FP64_2:
        cls = "fp64_other"
    elif b in FMA_PIPE:
        cls = "fma_pipe"
    elif b in ALU_PIPE:
        cls = "alu"
    else:
        cls = "rest"
    n[cls] += cnt
    port = cnt * (3.0 if cls == "dfma" else 2.06 if cls == "fp64_other" else 1.0 if cls == "fma_pipe" else 0.0)
    L.append((int(r[ix["Address"]], 16) - base, op, cnt, port, cls))
smsps = 148 * 4
port = 3.0 * n["dfma"] + 2.06 * n["fp64_other"] + n["fma_pipe"]
alu = 2.0 * n["alu"]
issue = float(sum(n.values()))
bound = max(port, alu, issue)
n64, nother = n["dfma"] + n["fp64_other"], n["fma_pipe"] + n["alu"] + n["rest"]
print(f"warp instructions: DFMA {n['dfma'] / 1e6:.1f} M, other FP64 {n['fp64_other'] / 1e6:.1f} M, FMA pipe {n['fma_pipe'] / 1e6:.1f} M, ALU pipe {n['alu'] / 1e6:.1f} M, "
      f"rest {n['rest'] / 1e6:.1f} M (total {issue / 1e6:.1f} M, non-FP64 {nother / 1e6:.1f} M)")
print(f"cycles: FP64/FMA port {port / 1e6:.1f} M, ALU pipe {alu / 1e6:.1f} M, issue {issue / 1e6:.1f} M -> bound {bound / smsps / 1e3:.1f} K cycles per SMSP")
if raw:
    rr = list(csv.reader(open(raw)))
    m = dict(zip(rr[0], rr[2]))
    cyc = float(m["sm__cycles_elapsed.max"].replace(",", ""))
    print(f"measured {cyc / 1e3:.1f} K cycles elapsed -> the kernel runs at {100 * bound / smsps / cyc:.1f} % of its pipe / issue bound")
if key:
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "issue_model.json")
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[key] = {"fp64_warp_instructions": n64, "dfma_warp_instructions": n["dfma"], "other_warp_instructions": nother, "fma_pipe_warp_instructions": n["fma_pipe"],
              "alu_pipe_warp_instructions": n["alu"], "fp64_issue_cycles": "3.0 per DFMA (distinct register sources), 2.06 per DMUL / DADD / DSETP",
              "smsps": smsps, "port_cycles_per_smsp": port / smsps, "alu_cycles_per_smsp": alu / smsps, "issue_slots_per_smsp": issue / smsps,
              "issue_cycles_per_smsp": bound / smsps, "source": os.path.basename(src)}
    d["_source"] = ("per-SASS-instruction execution counts of one launch (ncu --set full, source page); pipe costs measured by tools/pipe_microbench.cu "
                    "(profiles/r02_pipe_microbench.txt): DFMA 3.0 cycles of the SMSP's FP64 pipe with three distinct register sources, DMUL / DADD / DSETP 2.06, "
                    "FMA-pipe instructions share that dispatch port (1 cycle, no overlap), ALU-pipe instructions overlap it at 2 cycles each; bound = max(port, ALU, issue)")
    json.dump(d, open(path, "w"), indent=1)
seg, cur = [], None
for a, op, cnt, c, cls in L:
    if cur and cur[2] == cnt:
        cur[1] = a; cur[3] += c; cur[4] += 1; cur[5] += 1 if cls in ("dfma", "fp64_other") else 0
    else:
        cur = [a, a, cnt, c, 1, 1 if cls in ("dfma", "fp64_other") else 0]
        seg.append(cur)
for s in seg:
    if port and s[3] / port > 0.004:
        print(f"0x{s[0]:05x}-0x{s[1]:05x} executed {s[2]:9d} x {s[4]:4d} instr ({s[5]:3d} fp64)  {100 * s[3] / port:5.1f} % of the FP64/FMA port cost")
