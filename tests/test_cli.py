"""Command line tool: the reference demo's modes b / s / l / t on the B200 engine.
CPU tier: the same sources linked against the emulator build of the C ABI (test infrastructure) so the tool's
own logic — argument handling, validators, the batched self-test grid — is exercised without a GPU.
GPU tier: the shipped binary (msufsort_b200/lib/msufsort)."""
import os
import subprocess

import pytest

from cases import gen
from conftest import ROOT

LIBDIR = os.path.join(ROOT, "msufsort_b200", "lib")
EMUDIR = os.path.join(ROOT, "tests", "emu")
EMU_EXE = os.path.join(ROOT, "tests", "cpp", "msufsort_emu")


def build_emu_cli():
    srcs = [os.path.join(ROOT, "src", "executable", "msufsort", "main.cpp"), os.path.join(ROOT, "src", "library", "msufsort", "msufsort.cpp")]
    lib = os.path.join(EMUDIR, "libb200sa_emu.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-s", "emu"], cwd=ROOT, check=True)
    newest = max(os.path.getmtime(p) for p in srcs + [lib])
    if not os.path.exists(EMU_EXE) or os.path.getmtime(EMU_EXE) < newest:
        subprocess.run(["g++", "-std=c++17", "-O2", f"-I{ROOT}/src", f"-I{ROOT}/include", *srcs, "-o", EMU_EXE,
                        f"-L{EMUDIR}", "-lb200sa_emu", f"-Wl,-rpath,{EMUDIR}"], check=True)
    return EMU_EXE


def run(exe, *args):
    return subprocess.run([exe, *args], capture_output=True, text=True, timeout=600)


def _modes(exe, tmp_path, n):
    f = tmp_path / "in.bin"
    f.write_bytes(gen("markov3", n).tobytes())
    out = run(exe, "s", str(f), "4")
    assert out.returncode == 0 and "suffix array verified" in out.stdout, out.stdout + out.stderr
    out = run(exe, "b", str(f))
    assert out.returncode == 0 and "round trip verified" in out.stdout, out.stdout + out.stderr
    out = run(exe, "L", str(f))
    assert out.returncode == 0 and "lcp array verified" in out.stdout, out.stdout + out.stderr
    z = tmp_path / "zeros.bin"
    z.write_bytes(bytes(3000))
    out = run(exe, "l", str(z))
    assert out.returncode == 0 and "lcp array verified" in out.stdout, out.stdout + out.stderr


def test_cli_emu_modes(tmp_path):
    exe = build_emu_cli()
    _modes(exe, tmp_path, 20000)
    assert "usage" in run(exe).stdout
    assert "usage" in run(exe, "x", "nothing").stdout
    assert run(exe, "s", str(tmp_path / "missing.bin")).returncode == 2


def test_cli_emu_self_test_grid():
    out = run(build_emu_cli(), "t", "5", "60")
    assert out.returncode == 0 and "600 inputs, 0 errors" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_cli_gpu(tmp_path):
    exe = os.path.join(LIBDIR, "msufsort")
    assert os.path.exists(exe), "run `make cli`"
    _modes(exe, tmp_path, 1 << 20)
    out = run(exe, "t", "12", "300")
    assert out.returncode == 0 and "7200 inputs, 0 errors" in out.stdout, out.stdout + out.stderr
