"""Dev tool: registers / spills / smem per kernel from `nvcc -Xptxas -v` (run on one .cu of csrc/)."""
import re, subprocess, sys, os
CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "social_navigation_pyenvs_b200", "csrc")
src = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--expt-relaxed-constexpr", "-I", CSRC,
       "-I", os.path.join(CSRC, "..", "..", "include"), "-Xptxas=-v", "-c", os.path.join(CSRC, src), "-o", "/dev/null"]
err = subprocess.run(cmd, capture_output=True, text=True).stderr
name = None
for line in err.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"snp::\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(snp::.*$", "", name)
        spill = None
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        spill = m.groups()
    m = re.search(r"Used (\d+) registers", line)
    if m and name and flt in name:
        print(f"{name:70s} regs={m.group(1):>3s} stack/spill={spill}")
