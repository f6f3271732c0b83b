"""Builds librrrmc_b200.so (hand-written sm_100a CUDA behind the C ABI of include/rrrmc_b200.h) in-tree.

Each translation unit is compiled to an object (in parallel, only when it or a header changed) and the objects are
linked into the shared library; no relocatable device code is needed (kernels never call across units)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
SO = os.path.join(LIBDIR, "librrrmc_b200.so")
SOURCES = ["api.cu", "ea_multispin.cu", "ea_poisson.cu", "ea_tma.cu", "ea_normal.cu", "tempering.cu", "chain.cu", "chain_trace.cu", "chain_ea.cu", "chain_warp.cu", "sk_dense.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--fmad=false", "-Xptxas", "-v"]


def _includes(path, seen=None):
    """Quoted #include closure of a source file (so that editing one header rebuilds only its users)."""
    import re
    seen = set() if seen is None else seen
    for m in re.finditer(r'^\s*#\s*include\s+"([^"]+)"', open(path).read(), re.M):
        h = os.path.normpath(os.path.join(os.path.dirname(path), m.group(1)))
        if h not in seen and os.path.exists(h):
            seen.add(h)
            _includes(h, seen)
    return seen


def _headers(src=None):
    if src is not None:
        return sorted(_includes(src)) + [__file__]
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return hs + [os.path.join(HERE, "..", "include", "rrrmc_b200.h"), __file__]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    return _stale(SO, [os.path.join(CSRC, f) for f in SOURCES] + _headers())


def _compile(nvcc, src, obj):
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("RRRMC_NVCC_EXTRA", "").split() + ["-c", "-o", obj, src]   # (diagnostic builds: -DFLOW_DIAG)
    res = subprocess.run(cmd, capture_output=True, text=True)
    return res.returncode, " ".join(cmd) + "\n" + res.stdout + res.stderr


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    jobs = []
    for f in SOURCES:
        src, obj = os.path.join(CSRC, f), os.path.join(OBJDIR, f[:-3] + ".o")
        if force or _stale(obj, [src] + _headers(src)):
            jobs.append((src, obj))
    logs, failed = [], False
    with ThreadPoolExecutor(max_workers=max(1, len(jobs))) as ex:
        for rc, log in ex.map(lambda j: _compile(nvcc, *j), jobs):
            logs.append(log)
            failed |= rc != 0
    if not failed:
        cmd = [nvcc, "-shared", "-o", SO] + [os.path.join(OBJDIR, f[:-3] + ".o") for f in SOURCES]
        res = subprocess.run(cmd, capture_output=True, text=True)
        logs.append(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        failed = res.returncode != 0
    log = "\n".join(logs)
    mode = "a" if jobs and len(jobs) < len(SOURCES) and os.path.exists(os.path.join(LIBDIR, "build.log")) else "w"
    with open(os.path.join(LIBDIR, "build.log"), mode) as f:
        f.write(log)
    if failed:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building librrrmc_b200.so")
    if verbose:
        print(log)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
