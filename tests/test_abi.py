"""CPU tier: the C-ABI library loads and exports every symbol include/b200sa.h declares; without a
GPU every compute entry point fails loudly (no CPU fallback); host-side helpers behave."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, has_gpu

LIB = os.path.join(ROOT, "msufsort_b200", "lib", "libb200sa.so")
HEADER = os.path.join(ROOT, "include", "b200sa.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"B200SA_API\s+[\w\s\*]+?\b(b200sa_\w+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ("b200sa_suffix_array", "b200sa_bwt", "b200sa_unbwt", "b200sa_suffix_array_dev", "b200sa_bwt_dev",
                 "b200sa_unbwt_dev", "b200sa_create", "b200sa_destroy", "b200sa_last_error"):
        assert must in syms


def test_product_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "run `make lib` (python -c 'import __graft_entry__ as g; g.build()')"
    lib = C.CDLL(LIB)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/b200sa.h but not exported"


def test_python_binding_covers_every_declared_symbol():
    from msufsort_b200.api import ABI
    assert sorted(n for n, _, _ in ABI) == declared_symbols()


def test_product_library_contains_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_a_device():
    from msufsort_b200 import B200SAError
    from msufsort_b200.api import Engine
    with pytest.raises(B200SAError) as e:
        Engine(0)
    assert e.value.code == 2  # B200SA_ENODEVICE
    assert "no CPU fallback" in str(e.value)
    import msufsort_b200
    with pytest.raises(B200SAError):
        msufsort_b200.make_suffix_array(b"banana")


def test_facade_library_exports_reference_symbols():
    path = os.path.join(ROOT, "msufsort_b200", "lib", "libmsufsort.so")
    assert os.path.exists(path)
    out = subprocess.run(["nm", "-DC", path], capture_output=True, text=True).stdout
    for sig in ("maniscalco::msufsort::make_suffix_array(unsigned char const*, unsigned char const*)",
                "maniscalco::msufsort::forward_burrows_wheeler_transform(unsigned char*, unsigned char*)",
                "maniscalco::msufsort::reverse_burrows_wheeler_transform(unsigned char*, unsigned char*, int, int)",
                "maniscalco::msufsort::msufsort(int)"):
        assert sig in out, sig


def test_alphabet_plan_fits_64_bits():
    # mirror of plan_alphabet() in b200sa.cu: k symbols of `bits` bits + clamped length field
    for sigma in range(1, 257):
        bits = max(1, (sigma - 1).bit_length())
        k = max(c for c in range(1, 59) if c * bits + c.bit_length() <= 64)
        assert k >= 7 and k * bits + k.bit_length() <= 64


def test_textgen_is_deterministic():
    from msufsort_b200 import textgen as t
    a, b = t.markov3(10000), t.markov3(10000)
    assert np.array_equal(a, b)
    assert bytes(t.fib(13)) == b"abaababaabaab"
    assert set(bytes(t.acgt_rep(5000))) <= set(b"ACGT")
    assert t.rand(3).tolist() == t.rand(100)[:3].tolist()
