"""The C++ host side (include/single_rust_b200.hpp) — the stand-in for the Rust host code north_star asks for (no Rust
toolchain in this image). tests/cpp/host_mirror_test.cpp is written like the reference's own tests; it is compiled with
g++ against libsrb200.so (+ the C oracle as the checker) and run here."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
OUT_DIR = os.path.join(ROOT, "tests", "cpp", "build")
EXE = os.path.join(OUT_DIR, "host_mirror_test")


def build_exe():
    from oracle import oracle as O
    from singlerust_b200 import _ffi
    O.build()
    lib_dir, orc_dir = os.path.dirname(_ffi.lib_path()), os.path.join(ROOT, "oracle")
    deps = [SRC, os.path.join(ROOT, "include", "single_rust_b200.hpp"), os.path.join(ROOT, "include", "srb200.h")]
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(d) for d in deps):
        return EXE
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
           "-L", lib_dir, "-lsrb200", "-L", orc_dir, "-loracle", f"-Wl,-rpath,{lib_dir}", f"-Wl,-rpath,{orc_dir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return EXE


def test_cpp_host_compiles_and_fails_loudly_without_a_gpu():
    """Host logic (quantiles, FlexValue arms, chunk iterator) and: no device => single_rust::Error, never a CPU path."""
    import torch
    exe = build_exe()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the no-device behaviour is checked on the CPU box")
    r = subprocess.run([exe, "--cpu-check"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout


@pytest.mark.gpu
def test_cpp_host_parity_on_gpu():
    exe = build_exe()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert "0 failed" in r.stdout
