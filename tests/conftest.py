"""Shared fixtures.

Tiers:
  * ``-m "not gpu"``: oracle vs golden vectors / the reference build, host logic, ABI surface, and
    the kernel LOGIC under the CPU SIMT emulator (tests/emu, test infrastructure only).
  * ``-m gpu``: the parity tests proper — the nvcc-built library on a real B200 through the C ABI.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: larger CPU-side cases")


def _redirect_emu_library():
    """B200SA_EMU_SANITIZER=asan|ubsan (tools/emu_asan.sh): every test that loads tests/emu/libb200sa_emu.so gets the sanitizer
    build of the same sources (tests/emu/asan/, tests/emu/ubsan/; `make emu-asan emu-ubsan`) instead.  Test infrastructure only."""
    variant = os.environ.get("B200SA_EMU_SANITIZER", "")
    if variant not in ("asan", "ubsan"):
        return
    from msufsort_b200 import api
    plain = os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so")
    asan = os.path.join(ROOT, "tests", "emu", variant, "libb200sa_emu.so")
    init = api.Library.__init__

    def patched(self, path_or_cdll):
        if isinstance(path_or_cdll, str) and os.path.abspath(path_or_cdll) == plain:
            path_or_cdll = asan
        init(self, path_or_cdll)
    api.Library.__init__ = patched


_redirect_emu_library()


def _make(target):
    subprocess.run(["make", "-s", target], cwd=ROOT, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)


class Oracle:
    """ctypes view of oracle/liboracle.so and, when present, oracle/_ref/libmsufsort_ref.so."""

    def __init__(self):
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            _make("oracle")
        self.lib = C.CDLL(path)
        P, I64, I32 = C.c_void_p, C.c_int64, C.c_int32
        self.lib.oracle_make_suffix_array.argtypes = [P, I64, P]
        self.lib.oracle_make_suffix_array_bruteforce.argtypes = [P, I64, P]
        self.lib.oracle_forward_bwt.argtypes = [P, I64]
        self.lib.oracle_forward_bwt.restype = I32
        self.lib.oracle_bwt_from_sa.argtypes = [P, I64, P, P]
        self.lib.oracle_bwt_from_sa.restype = I32
        self.lib.oracle_reverse_bwt.argtypes = [P, I64, I32]
        self.lib.oracle_check_suffix_array.argtypes = [P, I64, P]
        self.lib.oracle_check_suffix_array.restype = I64
        self.lib.oracle_make_lcp_array.argtypes = [P, I64, P, P]
        self.lib.oracle_lcp_kasai.argtypes = [P, I64, P, P]
        self.lib.oracle_fnv1a64.argtypes = [P, I64]
        self.lib.oracle_fnv1a64.restype = C.c_uint64
        ref_path = os.path.join(ROOT, "oracle", "_ref", "libmsufsort_ref.so")
        self.ref = None
        if os.path.exists(ref_path):
            self.ref = C.CDLL(ref_path)
            self.ref.ref_make_suffix_array.argtypes = [P, I64, P, I32]
            self.ref.ref_forward_bwt.argtypes = [P, I64, I32]
            self.ref.ref_forward_bwt.restype = I32
            self.ref.ref_reverse_bwt.argtypes = [P, I64, I32, I32]
            self.ref.ref_lcp.argtypes = [P, I64, P, P, I32]

    # --- restatement
    def sa(self, text: np.ndarray) -> np.ndarray:
        text = np.ascontiguousarray(text, dtype=np.uint8)
        out = np.empty(text.size + 1, dtype=np.int32)
        assert self.lib.oracle_make_suffix_array(text.ctypes.data, text.size, out.ctypes.data) == 0
        return out

    def sa_bruteforce(self, text: np.ndarray) -> np.ndarray:
        text = np.ascontiguousarray(text, dtype=np.uint8)
        out = np.empty(text.size + 1, dtype=np.int32)
        assert self.lib.oracle_make_suffix_array_bruteforce(text.ctypes.data, text.size, out.ctypes.data) == 0
        return out

    def bwt(self, text: np.ndarray):
        buf = np.array(text, dtype=np.uint8, copy=True)
        s = self.lib.oracle_forward_bwt(buf.ctypes.data, buf.size)
        assert s >= 0
        return buf, int(s)

    def bwt_from_sa(self, text: np.ndarray, sa: np.ndarray):
        text = np.ascontiguousarray(text, dtype=np.uint8)
        sa = np.ascontiguousarray(sa, dtype=np.int32)
        out = np.empty(text.size, dtype=np.uint8)
        s = self.lib.oracle_bwt_from_sa(text.ctypes.data, text.size, sa.ctypes.data, out.ctypes.data)
        return out, int(s)

    def unbwt(self, bwt: np.ndarray, sentinel: int) -> np.ndarray:
        buf = np.array(bwt, dtype=np.uint8, copy=True)
        assert self.lib.oracle_reverse_bwt(buf.ctypes.data, buf.size, sentinel) == 0
        return buf

    def check_sa(self, text: np.ndarray, sa: np.ndarray) -> int:
        text = np.ascontiguousarray(text, dtype=np.uint8)
        sa = np.ascontiguousarray(sa, dtype=np.int32)
        return int(self.lib.oracle_check_suffix_array(text.ctypes.data, text.size, sa.ctypes.data))

    def lcp(self, text: np.ndarray, sa: np.ndarray, kasai: bool = False) -> np.ndarray:
        """n+1 entries aligned with the SA (lcp[0] = lcp[1] = 0); the reference demo's recursion or Kasai"""
        text = np.ascontiguousarray(text, dtype=np.uint8)
        sa = np.ascontiguousarray(sa, dtype=np.int32)
        out = np.empty(text.size + 1, dtype=np.int32)
        fn = self.lib.oracle_lcp_kasai if kasai else self.lib.oracle_make_lcp_array
        assert fn(text.ctypes.data, text.size, sa.ctypes.data, out.ctypes.data) == 0
        return out

    def fnv(self, arr: np.ndarray) -> int:
        a = np.ascontiguousarray(arr)
        return int(self.lib.oracle_fnv1a64(a.ctypes.data, a.nbytes))

    # --- the unmodified reference (oracle/_ref)
    def ref_sa(self, text: np.ndarray, threads: int = 1) -> np.ndarray:
        text = np.ascontiguousarray(text, dtype=np.uint8)
        out = np.empty(text.size + 1, dtype=np.int32)
        assert self.ref.ref_make_suffix_array(text.ctypes.data, text.size, out.ctypes.data, threads) == 0
        return out

    def ref_bwt(self, text: np.ndarray, threads: int = 1):
        buf = np.array(text, dtype=np.uint8, copy=True)
        s = self.ref.ref_forward_bwt(buf.ctypes.data, buf.size, threads)
        return buf, int(s)

    def ref_lcp(self, text: np.ndarray, sa: np.ndarray, threads: int = 1) -> np.ndarray:
        """the reference demo's lcp_multithreaded, mapped to this repository's convention (n+1 entries)"""
        text = np.ascontiguousarray(text, dtype=np.uint8)
        sa = np.ascontiguousarray(sa, dtype=np.int32)
        out = np.zeros(text.size + 1, dtype=np.int32)
        if text.size >= 2:
            tail = np.empty(text.size - 1, dtype=np.int32)
            assert self.ref.ref_lcp(text.ctypes.data, text.size, sa.ctypes.data, tail.ctypes.data, threads) == 0
            out[2:] = tail
        return out

    def ref_unbwt(self, bwt: np.ndarray, sentinel: int, threads: int = 1) -> np.ndarray:
        buf = np.array(bwt, dtype=np.uint8, copy=True)
        assert self.ref.ref_reverse_bwt(buf.ctypes.data, buf.size, sentinel, threads) == 0
        return buf


@pytest.fixture(scope="session")
def oracle():
    return Oracle()


@pytest.fixture(scope="session")
def ref(oracle):
    if oracle.ref is None:
        pytest.skip("oracle/_ref/libmsufsort_ref.so not built (reference tree absent)")
    return oracle


@pytest.fixture(scope="session")
def emu_engine():
    """Engine over the CPU SIMT emulator build of the kernels (tests only, never the product)."""
    from msufsort_b200.api import Engine, Library
    path = os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so")
    if not os.path.exists(path):
        _make("emu")
    eng = Engine(0, library=Library(path))
    yield eng
    eng.close()


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_engine():
    """Engine over the nvcc-built product library on cuda:0."""
    if not has_gpu():
        pytest.skip("no CUDA device")
    from msufsort_b200.api import Engine
    eng = Engine(0)
    yield eng
    eng.close()
