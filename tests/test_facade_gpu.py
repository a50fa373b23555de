"""GPU tier: a C++ caller compiled against the reference-shaped facade (src/library/msufsort.h) and
linked to libmsufsort.so + libb200sa.so reproduces the oracle's digests — the drop-in boundary."""
import os
import subprocess

import numpy as np
import pytest

from cases import gen
from conftest import ROOT, has_gpu

LIBDIR = os.path.join(ROOT, "msufsort_b200", "lib")
EXE = os.path.join(ROOT, "tests", "cpp", "facade_test")


def build_exe():
    src = os.path.join(ROOT, "tests", "cpp", "facade_test.cpp")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < os.path.getmtime(src):
        subprocess.run(["g++", "-std=c++17", "-O2", f"-I{ROOT}/src", f"-I{ROOT}/include", src, "-o", EXE,
                        f"-L{LIBDIR}", "-lmsufsort", "-lb200sa", f"-Wl,-rpath,{LIBDIR}"], check=True)


def test_facade_caller_compiles_against_drop_in_header():
    """CPU tier part: the reference-shaped caller compiles and links"""
    build_exe()
    assert os.path.exists(EXE)


@pytest.mark.skipif(has_gpu(), reason="no-GPU behaviour")
def test_facade_throws_without_gpu(tmp_path):
    build_exe()
    f = tmp_path / "in.bin"
    f.write_bytes(b"mississippi")
    out = subprocess.run([EXE, str(f)], capture_output=True, text=True)
    assert out.returncode == 1 and "EXCEPTION" in out.stdout and "no CPU fallback" in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [None, "0,0,0"], ids=["one-gpu", "sharded-3-contexts"])
@pytest.mark.parametrize("family,n", [("markov3", 300001), ("zeros", 5000), ("rand", 1), ("acgt_rep", 1 << 20)])
def test_facade_matches_oracle(oracle, tmp_path, family, n, devices):
    """the reference-shaped C++ caller; with MSUFSORT_DEVICES the facade shards every text over a group of contexts"""
    build_exe()
    x = gen(family, n)
    f = tmp_path / "in.bin"
    f.write_bytes(x.tobytes())
    env = dict(os.environ)
    if devices:
        env["MSUFSORT_DEVICES"] = devices
    out = subprocess.run([EXE, str(f)], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = dict(l.split(" ", 1) for l in out.stdout.strip().splitlines())
    sa = oracle.sa(x)
    bwt, s = oracle.bwt_from_sa(x, sa)
    assert lines["SA"] == f"{oracle.fnv(sa):016x} {n + 1}"
    assert lines["BWT"] == f"{oracle.fnv(bwt):016x} {s}"
    assert lines["UNBWT"] == "roundtrip-ok"
    assert lines["CLASS"] == "same"
