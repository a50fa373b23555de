"""A plain C caller of b200sa_group_* (tests/cpp/group_test.c): one text sharded over several contexts behind the reference's
three calls, bound exactly as a cgo / JNI / ctypes stub would bind them.  CPU tier: the emulator build of the ABI;
GPU tier: the product library, three contexts on cuda:0 (and one per GPU when the box has several)."""
import os
import subprocess

import pytest

from cases import gen
from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "cpp", "group_test.c")


def _build(exe, libdir, lib):
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-std=c11", "-O2", f"-I{ROOT}/include", SRC, "-o", exe, f"-L{libdir}", f"-l{lib}", f"-Wl,-rpath,{libdir}"], check=True)
    return exe


def _check(exe, oracle, tmp_path, family, n, devices):
    x = gen(family, n)
    f = tmp_path / "in.bin"
    f.write_bytes(x.tobytes())
    out = subprocess.run([exe, str(f), devices], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = dict(l.split(" ", 1) for l in out.stdout.strip().splitlines())
    sa = oracle.sa(x)
    bwt, s = oracle.bwt_from_sa(x, sa)
    assert lines["GROUP"] == str(len(devices.split(",")))
    assert lines["SA"] == f"{oracle.fnv(sa):016x} {n + 1}"
    assert lines["BWT"] == f"{oracle.fnv(bwt):016x} {s}"
    assert lines["UNBWT"] == "roundtrip-ok"
    assert lines["CORRUPT"] in ("rc=1 unchanged=1", "rc=0 unchanged=0")   # rejected (B200SA_EINVAL) unless the flipped bit happens to give a valid BWT


def test_c_group_caller_emu(oracle, tmp_path):
    emudir = os.path.join(ROOT, "tests", "emu")
    exe = _build(os.path.join(ROOT, "tests", "cpp", "group_test_emu"), emudir, "b200sa_emu")
    _check(exe, oracle, tmp_path, "markov3", 50021, "0,0,0")
    _check(exe, oracle, tmp_path, "fib", 20000, "0,0")


@pytest.mark.gpu
def test_c_group_caller_gpu(oracle, tmp_path):
    import torch
    libdir = os.path.join(ROOT, "msufsort_b200", "lib")
    exe = _build(os.path.join(ROOT, "tests", "cpp", "group_test"), libdir, "b200sa")
    _check(exe, oracle, tmp_path, "markov3", (1 << 22) + 3, "0,0,0")
    _check(exe, oracle, tmp_path, "acgt_rep", 1 << 21, "0,0,0,0,0")
    ng = torch.cuda.device_count()
    if ng >= 2:
        _check(exe, oracle, tmp_path, "markov3", (1 << 23) + 1, ",".join(str(g) for g in range(min(ng, 8))))
