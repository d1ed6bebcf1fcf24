"""Builds libxpoly_b200.so (the C-ABI shared library) in-tree with nvcc for sm_100a.

No torch, no JIT cache: explicit nvcc commands (one object per translation unit, compiled in
parallel, then one link) so the built .so travels with the repo snapshot.
`python -m xpoly_b200.build` or __graft_entry__.build().
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "libxpoly_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",  # FP64 parity: mul and add round separately (reference lpsol.h:1487-1488)
    "-Xcompiler", "-fPIC,-O2,-fno-fast-math,-ffp-contract=off",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h")) + \
        glob.glob(os.path.join(HERE, "host", "*.hpp"))


def _obj(src):
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(f) > t for f in deps)


def needs_build():
    return _stale(OUT, sources() + headers() + [os.path.abspath(__file__)])


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdrs = headers() + [os.path.abspath(__file__)]
    inc = ["-I", os.path.join(HERE, "..", "include")]

    def compile_one(src):
        obj = _obj(src)
        if not force and not _stale(obj, [src] + hdrs):
            return
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + inc + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(compile_one, sources()))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + \
        [_obj(s) for s in sources()] + ["-ldl"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
