"""Build libsnp_b200.so (hand-written sm_100a CUDA kernels + the C ABI of include/snp_b200.h) in-tree with nvcc.

    python -m social_navigation_pyenvs_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU.  Objects go to csrc/_build/, the library next to the package's
__init__.py so that it travels with the source tree to the GPU box.
"""
import concurrent.futures
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(PKG, "libsnp_b200.so")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-O2", "--expt-relaxed-constexpr", "-I", INCLUDE, "-I", CSRC]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".inl"))]
    files.append(os.path.join(INCLUDE, "snp_b200.h"))
    return max(os.path.getmtime(f) for f in files)


# The reset kernel replays the reference's scenario generators operation by operation (csrc/snp_reset_core.h): no FMA contraction
# there, so that a + b * c rounds twice as in NumPy (explicit fma() calls are kept).
PER_FILE_FLAGS = {"snp_reset.cu": ["-fmad=false"]}


def _compile(src, verbose, extra=(), obj_dir=OBJ):
    obj = os.path.join(obj_dir, src[:-3] + ".o")
    cmd = ["nvcc", *NVCC_FLAGS, *PER_FILE_FLAGS.get(src, []), *extra, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    srcs = _sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), srcs))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    objs = [o for o, _ in results]
    cmd = ["nvcc", "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


def build_variant(tag: str, defines) -> str:
    """Tuning experiments: a second library built with extra -D flags into variants/<tag>/ (select with SNP_B200_LIB)."""
    out_dir = os.path.join(PKG, "variants", tag)
    os.makedirs(out_dir, exist_ok=True)
    extra = [f"-D{d}" for d in defines]
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        objs = [o for o, _ in ex.map(lambda s: _compile(s, False, extra, out_dir), _sources())]
    lib = os.path.join(out_dir, "libsnp_b200.so")
    subprocess.run(["nvcc", "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"], check=True)
    for o in objs:  # only the library travels to the GPU box
        os.remove(o)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
