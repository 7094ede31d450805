"""Build libcpimh.so (hand-written sm_100a CUDA + the C ABI of include/cpimh.h) in-tree with nvcc.

    python -m unbiased_pimh_b200.build [--force]

One object per translation unit, compiled in parallel; the shared library lands next to this file
(unbiased_pimh_b200/libcpimh.so) so it travels to the GPU box with the repository snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libcpimh.so")
UNITS = (["model_ar%d.cu" % d for d in range(16, 4, -1)] + ["model_lorenz16.cu", "model_lorenz8.cu", "model_sk.cu", "model_sv.cu", "model_lorenz4.cu",
          "model_ar4.cu", "model_ar3.cu", "model_ar2.cu", "model_ar1.cu", "model_gss.cu"]      # slowest first; get_ar(d) for every d <= 16
         + ["api.cu", "api_chains.cu", "api_ccpf.cu", "primitives.cu", "lorenz_wide.cu", "wide.cu"])
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "cpimh.h"))
    return hs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(unit, verbose):
    src = os.path.join(CSRC, unit)
    obj = os.path.join(OBJ, unit.replace(".cu", ".o"))
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (unit, r.stdout, r.stderr))
    return unit, r.stderr


def build_library(force=False, verbose=False, jobs=None):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    todo = [u for u in UNITS if force or _stale(os.path.join(OBJ, u.replace(".cu", ".o")), [os.path.join(CSRC, u)] + hdrs)]
    logs = {}
    if todo:
        with ThreadPoolExecutor(max_workers=jobs or min(len(todo), os.cpu_count() or 4)) as ex:
            for unit, log in ex.map(lambda u: _compile(u, verbose), todo):
                logs[unit] = log
    objs = [os.path.join(OBJ, u.replace(".cu", ".o")) for u in UNITS]
    if todo or _stale(LIB, objs):
        cmd = ["nvcc", "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB, logs


if __name__ == "__main__":
    lib, logs = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    for u, l in logs.items():
        if l.strip():
            print("==", u)
            print(l)
    print(lib)
