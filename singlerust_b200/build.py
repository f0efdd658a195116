"""Builds libsrb200.so (sm_100a) in-tree with nvcc. `python -m singlerust_b200.build [--force]`."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsrb200.so")
SOURCES = ["api.cu", "upload.cu", "major_stats.cu", "minor_moments.cu", "select.cu", "pca.cu", "gram_tc.cu", "gram_tc2.cu", "eig.cu", "comm.cu",
           "synth.cu", "stream.cu", "convert.cu", "subset.cu"]
HOST_SOURCES = ["host_pack.cpp"]  # plain C++ (g++): host-side marshalling of the upload path, no CUDA
CXX = os.environ.get("CXX", "/usr/bin/g++")
CXXFLAGS = ["-O3", "-std=c++17", "-fPIC", "-pthread", "-Wall"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-ccbin", "/usr/bin/g++", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "srb200.h"))
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + hdrs):
            jobs.append((s, o))

    for src in HOST_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cpp", ".o"))
        if force or _stale(o, [s] + hdrs):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        if s.endswith(".cpp"):
            r = subprocess.run([CXX] + CXXFLAGS + ["-c", s, "-o", o], capture_output=True, text=True)
        else:
            r = subprocess.run([NVCC] + FLAGS + ["-c", s, "-o", o], capture_output=True, text=True)
        return s, r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, r in ex.map(compile_one, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {s}")
            with open(os.path.join(objdir, os.path.basename(s) + ".ptxas.log"), "w") as f:
                f.write(r.stderr)
    objs = [os.path.join(objdir, src.replace(".cu", ".o")) for src in SOURCES]
    objs += [os.path.join(objdir, src.replace(".cpp", ".o")) for src in HOST_SOURCES]
    if force or jobs or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-ccbin", "/usr/bin/g++", "-lcusolver", "-lcublas", "-lcuda", "-ldl", "-lpthread",
                                                      "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
