"""msufsort_b200 — B200-native suffix array / BWT / inverse BWT engine.

Host-side mirror of the reference's public interface (``maniscalco::make_suffix_array``,
``forward_burrows_wheeler_transform``, ``reverse_burrows_wheeler_transform``;
/root/reference/src/library/msufsort/msufsort.h:403-426) over the C ABI in ``include/b200sa.h``.
The compute path is hand-written CUDA for sm_100a in ``msufsort_b200/lib/libb200sa.so``; there is
no CPU fallback — importing works anywhere, computing raises without the library or a GPU.
"""
from .api import (  # noqa: F401
    B200SAError,
    Library,
    Engine,
    Pipeline,
    load_library,
    make_suffix_array,
    forward_burrows_wheeler_transform,
    reverse_burrows_wheeler_transform,
    make_lcp_array,
    PHASES,
)
from . import textgen  # noqa: F401

__all__ = [
    "B200SAError", "Library", "Engine", "Pipeline", "load_library", "make_suffix_array",
    "forward_burrows_wheeler_transform", "reverse_burrows_wheeler_transform", "make_lcp_array", "textgen", "PHASES",
]
