"""Build the CPU oracle shared library (test infrastructure): gcc on oracle/snp_oracle.c -> oracle/_build/, and stage the
live reference for the GPU box.

The reference is pure Python.  In the build container it is imported from /root/reference (tests/golden/make_golden.py records
the golden fixtures from it, tests/test_oracle_live_reference.py re-derives the pin on fresh seeds).  /root/reference does not
exist on the GPU box, so `stage_reference()` lays the two packages the path imports (social_gym, crowd_nav: .py files and the
.config files they read, nothing else) out under oracle/_ref/ -- git-ignored, never committed, but shipped by gpurun like the
built libraries.  There the same live cross-checks run next to the CUDA path (tests/test_gpu_live_reference.py) and bench.py
times the reference's own serial update (`cpu_baseline.kind = "reference"`, `--impl reference`).  oracle/reference.py is the
only importer; the product package never touches it.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "snp_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libsnp_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC", "-Wall",
           "-o", LIB, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


REFERENCE_SRC = "/root/reference"
REF_DIR = os.path.join(HERE, "_ref")


def stage_reference(force: bool = False):
    """Copy the reference's Python packages (unmodified) to oracle/_ref/.  Returns the staged root, or None when /root/reference is
    absent (GPU box: the staged copy that travelled with the snapshot is used as is)."""
    if not os.path.isdir(os.path.join(REFERENCE_SRC, "social_gym")):
        return REF_DIR if os.path.isdir(os.path.join(REF_DIR, "social_gym")) else None
    marker = os.path.join(REF_DIR, ".staged_from")
    if not force and os.path.exists(marker):
        return REF_DIR
    keep = (".py", ".config")
    for pkg in ("social_gym", "crowd_nav"):
        dst_pkg = os.path.join(REF_DIR, pkg)
        shutil.rmtree(dst_pkg, ignore_errors=True)
        for root, dirs, files in os.walk(os.path.join(REFERENCE_SRC, pkg)):
            dirs[:] = [d for d in dirs if d not in ("__pycache__", "fonts", "data", "output")]
            rel = os.path.relpath(root, REFERENCE_SRC)
            for f in files:
                if f.endswith(keep):
                    os.makedirs(os.path.join(REF_DIR, rel), exist_ok=True)
                    shutil.copyfile(os.path.join(root, f), os.path.join(REF_DIR, rel, f))
    with open(marker, "w") as f:
        f.write(REFERENCE_SRC + "\n")
    return REF_DIR


if __name__ == "__main__":
    print(build(force=True))
    print(stage_reference(force=True))
