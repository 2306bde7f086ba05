"""Writes profiles/<round>_summary.md from the committed evidence files (bench lines, scale lines, pipe model).
    python tools/make_summary.py r02
"""
import json
import os
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r02"
P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles") + "/"
lines = [json.loads(l) for l in open(P + f"{R}_bench_lines.jsonl")]
scale = [json.loads(l) for l in open(P + f"{R}_scale_lines.jsonl")] if os.path.exists(P + f"{R}_scale_lines.jsonl") else []
im = json.load(open(P + "issue_model.json"))
out = [f"# Round {int(R[1:])} evidence summary (B200, sm_100a, 1965 MHz under load, no throttling)\n",
       f"All numbers come from files in this directory: `{R}_bench_lines.jsonl` (one box, `tools/final_run.sh`), `{R}_scale_lines.jsonl` (one 8-GPU box), the ncu captures of",
       f"`tools/profile_round.sh` (`ncu --set full --clock-control none`, one launch per kernel; not bench values), `{R}_pipe_microbench.txt` (`tools/pipe_microbench.cu`).\n",
       "## bench.py lines\n", "| workload | dtype | value | ms / step | end to end | roofline frac | pipe-bound frac |", "|---|---|---|---|---|---|---|"]
for d in lines:
    r = d.get("roofline") or {}
    iss = (r.get("issue") or {}).get("frac")
    tag = " (reference arm: live reference, one env per process)" if d.get("impl") == "reference" else ""
    out.append(f"| {d['config']['workload']}{tag} | {d['dtype']} | {d['value']:.3e} {d['unit']} | {d['ms_per_step']:.4f} | {d['e2e']['value']:.3e} | "
               f"{'' if r.get('frac') is None else '%.3f' % r['frac']} | {'' if iss is None else '%.2f' % iss} |")
cbs = [d for d in lines if d.get("cpu_baseline") and d.get("impl") != "reference"]
if cbs:
    cb = cbs[0]["cpu_baseline"]
    port = (cb.get("c_port") or {}).get("value")
    out.append(f"\ncpu_baseline of the default workload: **{cb['value']:.3e} {cb['unit']}** on {cb['cores']} host cores, kind `{cb['kind']}` ({cb['sample']})"
               + (f"; C port beside it: {port:.3e}.\n" if port else ".\n"))
if scale:
    out += ["## Scaling (default workload env-sharded, weak; one 65536-human crowd agent-sharded, strong)\n",
            "| GPUs | agent-steps/s | ms / step | end to end | large crowd ms / sub-step (culled) | all ordered pairs | bit-equal to 1 GPU | exchange |", "|---|---|---|---|---|---|---|---|"]
    for d in scale:
        lc = d["large_crowd"]
        out.append(f"| {d['n_gpus']} | {d['value']:.3e} | {d['ms_per_step']:.4f} | {d['e2e']['value']:.3e} | {lc['ms_per_substep']:.4f} | "
                   f"{lc['ms_per_substep_all_pairs']:.3f} | {lc['bit_equal_to_single_gpu']} | {lc['exchange'][:60]} |")
out += ["\n## Pipe / issue model (`issue_model.json`, `tools/issue_cost.py`)\n",
        "Measured pipe costs per SM sub-partition: DFMA with three distinct register sources 3.0 cycles (2.64 with a reused / constant operand), DMUL / DADD 2.06; FMA-pipe",
        "instructions (IMAD, FFMA) do not issue in an FP64 instruction's shadow, ALU-pipe instructions (LOP3, SEL, ISETP, VIMNMX, IADD3) do, at 2 cycles each.",
        "bound = max(FP64/FMA port, ALU pipe, issue slots).\n",
        "| workload:dtype | FP64 warp instr | of which DFMA | other warp instr | bound cycles / SMSP |", "|---|---|---|---|---|"]
for k, v in im.items():
    if k.startswith("_"):
        continue
    out.append(f"| {k} | {v['fp64_warp_instructions'] / 1e6:.1f} M | {v.get('dfma_warp_instructions', 0) / 1e6:.1f} M | {v['other_warp_instructions'] / 1e6:.1f} M | "
               f"{v['issue_cycles_per_smsp'] / 1e3:.1f} K |")
out += ["\nk_step fp64 runs at 63 % of its bound (fp32 69 %, k_large_pairs 95 %, k_laser 76 %): the remainder is dependency latency -- one full wave of",
        "4 / 8 / 12 / 16 warps per SM takes 0.101 / 0.111 / 0.128 / 0.147 ms (`SNP_BENCH_ENVS` = 592 / 1184 / 1776 / 2368, measured mid-round), i.e. 0.09 ms for a lone",
        "warp per scheduler plus 0.0036 ms per additional warp (DESIGN.md 4.1).\n",
        "## k_step fp64, round 1 -> round 2 (ncu, one launch of the default workload)\n",
        "| | round 1 | round 2 |", "|---|---|---|",
        "| warp instructions | 182.4 M | 144.8 M |", "| FP64 (DFMA + DMUL + DADD + DSETP) | 67.5 M | 59.0 M |", "| other | 115.0 M | 85.8 M |",
        "| UMOV | 7.7 M | 2.2 M |", "| duration under ncu | 273.9 us | 229.0 us |", "| bench ms per launch | 0.2666 | 0.2212 |",
        "| roofline frac (FP64 FMA peak measured in the run) | 0.297 | 0.360 |"]
open(P + f"{R}_summary.md", "w").write("\n".join(out) + "\n")
print(f"wrote profiles/{R}_summary.md ({len(out)} lines)")
