"""Build libspk.so (the sm_100a CUDA library) and the CPU oracle in-tree.

`python -m subphaser_b200.build` compiles every `csrc/*.cu` with
`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo` (cross-compiles without a GPU) and links
`subphaser_b200/libspk.so`.  Object files are cached by source mtime under `subphaser_b200/csrc/build/`.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libspk.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-diag-suppress", "550",
]
# files whose fp64 arithmetic must reproduce numpy / Python operation by operation
NO_FMA = {"spk_matrix.cu", "spk_pmatrix.cu", "spk_stats.cu", "spk_cluster.cu"}


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libspk.so cannot be built")
    return nvcc


def _newer(src, dst, extra=()):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(p) > t for p in (src,) + tuple(extra))


def build_lib(verbose=False, force=False):
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    headers = tuple(
        os.path.join(d, f)
        for d in (CSRC, os.path.join(ROOT, "include"))
        for f in os.listdir(d)
        if f.endswith((".cuh", ".h"))
    )
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    jobs = []
    objs = []
    for f in sources:
        src = os.path.join(CSRC, f)
        obj = os.path.join(OBJDIR, f[:-3] + ".o")
        objs.append(obj)
        if force or _newer(src, obj, headers):
            cmd = [nvcc] + NVCC_FLAGS + (["-fmad=false"] if f in NO_FMA else [])
            if verbose:
                cmd += ["-Xptxas", "-v"]
            cmd += ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for " + cmd[-3])
    if jobs or force or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link of libspk.so failed")
    return LIB


def build_oracle(force=False):
    """Compile the CPU oracle (test infrastructure) — building the checker is not using it."""
    odir = os.path.join(ROOT, "oracle")
    mk = os.path.join(odir, "Makefile")
    if os.path.exists(mk):
        r = subprocess.run(["make", "-C", odir] + (["-B"] if force else []), capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("oracle build failed")


if __name__ == "__main__":
    build_lib(verbose="-v" in sys.argv, force="-f" in sys.argv)
    build_oracle()
    print(LIB)
