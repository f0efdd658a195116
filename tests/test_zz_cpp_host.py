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
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run([exe, "--cpu-check", d], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout


@pytest.mark.parametrize("index_dtype", ["uint64", "uint32"])
def test_cpp_reads_the_chunk_store_python_writes(index_dtype):
    """One on-disk format for both host sides: BackedAnnData.write_store (Python) -> StoreChunkSource (C++)."""
    import tempfile

    import numpy as np

    from singlerust_b200.anndata import BackedAnnData
    from tests._util import random_csr
    exe = build_exe()
    a = random_csr(np.random.default_rng(3), 500, 90, 0.1, empty_rows=(0, 499))
    with tempfile.TemporaryDirectory() as d:
        BackedAnnData.write_store(d, a, np.dtype(index_dtype))
        r = subprocess.run([exe, "--store-sums", d, "128"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        rows = [line.split() for line in r.stdout.strip().splitlines()]
        back = BackedAnnData.open_store(d)
        want = [(ch.shape[0], ch.nnz, float(ch.data.astype(np.float64).sum()), int(ch.indices.astype(np.int64).sum()))
                for ch, _s, _e in back.iter_chunks(128)]
    assert len(rows) == len(want) == 4
    for got, w in zip(rows, want):
        assert (int(got[0]), int(got[1]), int(got[3])) == (w[0], w[1], w[3])
        assert float(got[2]) == w[2]


@pytest.mark.gpu
def test_cpp_host_parity_on_gpu():
    exe = build_exe()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert "0 failed" in r.stdout
