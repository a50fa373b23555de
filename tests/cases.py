"""Input families shared by the emulator tier and the GPU tier."""
import numpy as np
from msufsort_b200 import textgen as t

# edge sizes named in SURVEY.md §8(d) "extra" row
EDGE_SIZES = [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 255, 256, 257, 4095, 4096, 4097]

FAMILIES = ["rand", "markov3", "acgt_rep", "periodic7", "periodic1009", "fib", "zeros", "abcabca",
            "sigma2", "sigma3", "sigma4", "zero_tail"]


def gen(name: str, n: int) -> np.ndarray:
    return np.ascontiguousarray(t.GENERATORS[name](n))


def small_alphabet_exhaustive(max_len: int = 8, sigma: int = 2):
    """every string over `sigma` symbols (bytes 0..sigma-1) up to max_len: byte-0-vs-sentinel edge"""
    for n in range(1, max_len + 1):
        for code in range(sigma ** n):
            s = np.empty(n, dtype=np.uint8)
            c = code
            for i in range(n):
                s[i] = c % sigma
                c //= sigma
            yield s
