"""Synthetic inputs of SURVEY.md §8(d) (ctypes over msufsort_b200/lib/libb200sa_textgen.so)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "lib", "libb200sa_textgen.so")
_lib = None

SEED_RAND = 0xB2000001
SEED_MARKOV = 0xB2000002
SEED_ACGT = 0xB2000003
SEED_PERIODIC = 0xB2000005


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise RuntimeError(f"{_PATH} is missing — run `make textgen`")
        _lib = C.CDLL(_PATH)
        P, I64, U64 = C.c_void_p, C.c_int64, C.c_uint64
        _lib.textgen_rand.argtypes = [P, I64, U64]
        _lib.textgen_alphabet.argtypes = [P, I64, U64, C.c_int, C.c_int]
        _lib.textgen_markov3.argtypes = [P, I64, U64]
        _lib.textgen_acgt_rep.argtypes = [P, I64, U64, I64]
        _lib.textgen_periodic.argtypes = [P, I64, U64, I64]
        _lib.textgen_fib.argtypes = [P, I64]
        _lib.textgen_zero_tail.argtypes = [P, I64, U64, C.c_int, I64]
        _lib.textgen_reference_selftest.argtypes = [P, I64, C.c_uint, C.c_int]
        _lib.textgen_fnv1a64.argtypes = [P, I64]
        _lib.textgen_fnv1a64.restype = U64
        for f in ("textgen_rand", "textgen_alphabet", "textgen_markov3", "textgen_acgt_rep", "textgen_periodic",
                  "textgen_fib", "textgen_zero_tail", "textgen_reference_selftest"):
            getattr(_lib, f).restype = None
    return _lib


def _out(n: int, out=None) -> np.ndarray:
    if out is None:
        return np.empty(n, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size >= n and out.flags.c_contiguous
    return out


def rand(n: int, seed: int = SEED_RAND, out=None) -> np.ndarray:
    a = _out(n, out); _load().textgen_rand(a.ctypes.data, n, seed); return a[:n]


def alphabet(n: int, sigma: int, seed: int = SEED_RAND, base: int = 0, out=None) -> np.ndarray:
    a = _out(n, out); _load().textgen_alphabet(a.ctypes.data, n, seed, sigma, base); return a[:n]


def markov3(n: int, seed: int = SEED_MARKOV, out=None) -> np.ndarray:
    a = _out(n, out); _load().textgen_markov3(a.ctypes.data, n, seed); return a[:n]


def acgt_rep(n: int, seed: int = SEED_ACGT, repeats: int = -1, out=None) -> np.ndarray:
    if repeats < 0:
        repeats = max(1, n >> 18)
    a = _out(n, out); _load().textgen_acgt_rep(a.ctypes.data, n, seed, repeats); return a[:n]


def periodic(n: int, p: int = 7, seed: int = SEED_PERIODIC, out=None) -> np.ndarray:
    a = _out(n, out); _load().textgen_periodic(a.ctypes.data, n, seed, p); return a[:n]


def fib(n: int, out=None) -> np.ndarray:
    a = _out(n, out); _load().textgen_fib(a.ctypes.data, n); return a[:n]


def zeros(n: int) -> np.ndarray:
    return np.zeros(n, dtype=np.uint8)


def tiled(pattern: bytes, n: int) -> np.ndarray:
    p = np.frombuffer(pattern, dtype=np.uint8)
    return np.tile(p, n // len(p) + 1)[:n].copy()


def zero_tail(n: int, sigma: int = 3, tail: int = 5, seed: int = SEED_RAND, out=None) -> np.ndarray:
    a = _out(n, out); _load().textgen_zero_tail(a.ctypes.data, n, seed, sigma, tail); return a[:n]


def reference_selftest(n: int, sigma: int, seed: int) -> np.ndarray:
    a = _out(n); _load().textgen_reference_selftest(a.ctypes.data, n, seed & 0xFFFFFFFF, sigma); return a


def fnv1a64(arr: np.ndarray) -> int:
    a = np.ascontiguousarray(arr)
    return int(_load().textgen_fnv1a64(a.ctypes.data, a.nbytes))


GENERATORS = {
    "rand": lambda n: rand(n),
    "markov3": lambda n: markov3(n),
    "acgt_rep": lambda n: acgt_rep(n),
    "periodic7": lambda n: periodic(n, 7),
    "periodic1009": lambda n: periodic(n, 1009),
    "fib": lambda n: fib(n),
    "zeros": lambda n: zeros(n),
    "abcabca": lambda n: tiled(b"abcabca", n),
    "sigma2": lambda n: alphabet(n, 2),
    "sigma3": lambda n: alphabet(n, 3),
    "sigma4": lambda n: alphabet(n, 4),
    "zero_tail": lambda n: zero_tail(n),
}
