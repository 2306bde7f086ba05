"""Turn the scratch ncu outputs of tools/profile_round.sh (gpurun_out/<round>_*) into the tracked summaries under profiles/."""
import csv
import collections
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def ncu_csv(rep, page, extra=()):
    """Page of a capture: the CSV tools/profile_round.sh exported on the GPU box, else read from the .ncu-rep here."""
    exported = rep[:-8] + (".raw.csv" if page == "raw" else ".source.csv")
    if os.path.exists(exported):
        return list(csv.reader(open(exported)))
    return list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout.splitlines()))


def launches():
    path = os.path.join(SRC, f"{R}_launches_bench.csv")
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    def phase(name):
        if "_peak" in name:
            return "after the timed region: roofline denominators (snp_measure_pipe_peak)"
        if "k_unpack" in name or "FillFunctor<double>" in name or "FillFunctor<int>" in name:
            return "setup"
        if "FillFunctor<unsigned char>" in name:
            return "timed loop, outside the event pairs: 256 MiB L2 flush between steps"
        return "step"
    step_total = sum(sum(v) for k, v in agg.items() if phase(k) == "step")
    with open(os.path.join(OUT, f"{R}_launches_bench.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "mean_us", "total_us", "phase", "share_of_step_gpu_time"])
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            w.writerow([k[:110], len(v), round(sum(v) / len(v) / 1e3, 2), round(sum(v) / 1e3, 1), phase(k),
                        round(sum(v) / step_total, 4) if phase(k) == "step" else ""])


def metrics():
    reps = sorted({f[:-8] + ".ncu-rep" for f in os.listdir(SRC) if f.startswith(R + "_k_") and f.endswith(".raw.csv")} |
                  {f for f in os.listdir(SRC) if f.startswith(R + "_k_") and f.endswith(".ncu-rep")})
    cols, names = [], []
    SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    units = {}
    for f in reps:
        rows = ncu_csv(os.path.join(SRC, f), "raw")
        col, unit = dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))
        for m in WANT:  # ncu picks a unit per capture: bring times to us and sizes to Mbyte so that a row reads across kernels
            if m in col and unit.get(m) in SCALE and col[m]:
                col[m] = "%.6g" % (float(col[m].replace(",", "")) * SCALE[unit[m]])
                unit[m] = "us" if unit[m] in ("ns", "us", "ms", "s") else "Mbyte"
        cols.append(col)
        units.update({m: unit.get(m, "") for m in WANT})
        names.append(f[len(R) + 1:-8] + " :: " + cols[-1].get("Kernel Name", "")[:60])
    with open(os.path.join(OUT, f"{R}_ncu_metrics.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + names)
        for m in WANT:
            w.writerow([m, units.get(m, "")] + [c.get(m, "") for c in cols])
    for f in reps:  # opcode mix + stall reasons of each capture
        rows = ncu_csv(os.path.join(SRC, f), "source", ["--print-source", "sass"])
        hdr = rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        ops, stalls, tot = collections.Counter(), collections.Counter(), 0
        for r in rows[2:]:
            if len(r) < len(hdr):
                continue
            n = int(float(r[ix["Instructions Executed"]] or 0))
            t = r[ix["Source"]].split()
            op = (t[1] if t[0].startswith("@") else t[0]).rstrip(";").split(".")[0]
            ops[op] += n
            tot += n
            for h in hdr:
                if h.startswith("stall_") and "Not Issued" not in h:
                    stalls[h] += int(float(r[ix[h]] or 0))
        with open(os.path.join(OUT, f"{R}_{f[len(R) + 1:-8]}_sass_mix.csv"), "w") as g:
            w = csv.writer(g)
            w.writerow(["opcode", "warp_instructions", "share"])
            for k, v in ops.most_common(24):
                w.writerow([k, v, round(v / tot, 4)])
            w.writerow([])
            w.writerow(["stall_reason", "samples"])
            for k, v in stalls.most_common(8):
                w.writerow([k, v])


if __name__ == "__main__":
    launches()
    metrics()
    print(open(os.path.join(OUT, f"{R}_launches_bench.csv")).read())
    print(open(os.path.join(OUT, f"{R}_ncu_metrics.csv")).read())
