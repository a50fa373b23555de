"""C++ extension header (src/library/msufsort/msufsort_b200.h: LCP + batched transforms over the C ABI): a caller is
compiled against it; CPU tier links the emulator build of the ABI (test infrastructure), GPU tier the product library."""
import os
import subprocess

import numpy as np
import pytest

from cases import gen
from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "cpp", "ext_test.cpp")


GROUP_SRC = os.path.join(ROOT, "tests", "cpp", "ext_group_test.cpp")


def _build(exe, libdir, lib, src=SRC):
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "src", "library", "msufsort", "msufsort_b200.h"))):
        subprocess.run(["g++", "-std=c++17", "-O2", f"-I{ROOT}/src", f"-I{ROOT}/include", src, "-o", exe, f"-L{libdir}", f"-l{lib}",
                        f"-Wl,-rpath,{libdir}"], check=True)
    return exe


def check_group_caller(exe, tmp_path, n, block, devices):
    """tests/cpp/ext_group_test.cpp: a batch spread over `devices` (maniscalco::b200::gpu_group) equals the one-context batch"""
    f = tmp_path / "in.bin"
    f.write_bytes(gen("markov3", n).tobytes())
    out = subprocess.run([exe, str(f), str(block), devices], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[0] == f"GROUP {len(devices.split(','))} BLOCKS {(n + block - 1) // block + 1}"
    assert lines[1:] == ["SA same", "BWT same", "ROUNDTRIP ok"], out.stdout


def _check(exe, oracle, tmp_path, n, block):
    x = gen("markov3", n)
    f = tmp_path / "in.bin"
    f.write_bytes(x.tobytes())
    out = subprocess.run([exe, str(f), str(block)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    sa = oracle.sa(x)
    assert lines[0] == f"LCP {oracle.fnv(sa):016x} {oracle.fnv(oracle.lcp(x, sa, kasai=True)):016x}"
    blocks = [x[a:a + block] for a in range(0, n, block)] + [x[:0]]
    for b, blk in enumerate(blocks):
        if blk.size:
            bsa = oracle.sa(blk)
            bw, s = oracle.bwt_from_sa(blk, bsa)
            want = f"BLOCK {b} {blk.size} {oracle.fnv(bsa):016x} {oracle.fnv(bw):016x} {s}"
        else:
            want = f"BLOCK {b} 0 {oracle.fnv(np.zeros(1, np.int32)):016x} {oracle.fnv(blk):016x} 0"
        assert lines[1 + b] == want, b
    assert lines[-1] == "ROUNDTRIP ok"


def test_cpp_extension_header_emu(oracle, tmp_path):
    emudir = os.path.join(ROOT, "tests", "emu")
    if not os.path.exists(os.path.join(emudir, "libb200sa_emu.so")):
        subprocess.run(["make", "-s", "emu"], cwd=ROOT, check=True)
    exe = _build(os.path.join(ROOT, "tests", "cpp", "ext_test_emu"), emudir, "b200sa_emu")
    _check(exe, oracle, tmp_path, 30011, 7000)


def test_cpp_extension_header_group_emu(tmp_path):
    emudir = os.path.join(ROOT, "tests", "emu")
    exe = _build(os.path.join(ROOT, "tests", "cpp", "ext_group_test_emu"), emudir, "b200sa_emu", GROUP_SRC)
    check_group_caller(exe, tmp_path, 60011, 3000, "0,0,0")


@pytest.mark.gpu
def test_cpp_extension_header_gpu(oracle, tmp_path):
    libdir = os.path.join(ROOT, "msufsort_b200", "lib")
    exe = _build(os.path.join(ROOT, "tests", "cpp", "ext_test"), libdir, "b200sa")
    _check(exe, oracle, tmp_path, 1 << 20, 100000)
