"""Builds libsednet_b200.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python sed-net_b200/build.py [--force]

One object per .cu, compiled in parallel, then linked with -shared.  The built library is git-ignored but
travels to the GPU box with the repository snapshot.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libsednet_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(HERE, "..", "include", "sednet_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, force, hdr_m):
    s = os.path.join(CSRC, src)
    o = os.path.join(OBJ, src[:-3] + ".o")
    if not force and os.path.exists(o) and os.path.getmtime(o) > max(os.path.getmtime(s), hdr_m):
        return o, ""
    r = subprocess.run([NVCC, *FLAGS, "-c", s, "-o", o], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    with open(o + ".log", "w") as f:
        f.write(r.stderr)
    return o, r.stderr


def build_library(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdr_m = _deps_mtime()
    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        res = list(ex.map(lambda s: _compile(s, force, hdr_m), _sources()))
    objs = [o for o, _ in res]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        for _, log in res:
            for line in log.splitlines():
                if "spill" in line and "0 bytes spill stores, 0 bytes spill loads" not in line:
                    print(line)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
