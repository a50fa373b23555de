"""ctypes binding of include/b200sa.h and the reference-shaped Python entry points.

Reference interface mirrored (names, argument meaning, in-place behaviour, return values):
  maniscalco::make_suffix_array(begin, end, numThreads)                  msufsort.h:432-445
  maniscalco::forward_burrows_wheeler_transform(begin, end, numThreads)  msufsort.h:449-462
  maniscalco::reverse_burrows_wheeler_transform(begin, end, sentinelIndex, numThreads)  :466-476
``numThreads`` is accepted and ignored (the work runs on the GPU); results do not depend on it in
the reference either.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(_HERE, "lib", "libb200sa.so")

PHASES = {
    "alphabet": 0, "pack": 1, "sort_hist": 2, "sort_pass": 3, "build": 4, "rerank": 5,
    "bwt": 6, "unbwt_build": 7, "unbwt_walk": 8, "check": 9, "segsort": 10, "isa": 11, "lcp": 12, "peer_send": 13, "peer_apply": 14,
}
_PH_COUNT = 16


class B200SAError(RuntimeError):
    """Raised for every non-zero status of the C ABI (the C++ facade throws std::runtime_error)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"b200sa error {code}: {message}")
        self.code = code


class _Profile(C.Structure):
    _fields_ = [
        ("ms", C.c_double * _PH_COUNT),
        ("launches", C.c_uint64 * _PH_COUNT),
        ("alg_bytes", C.c_uint64 * _PH_COUNT),
        ("rounds", C.c_uint64),
        ("sort_passes", C.c_uint64),
        ("sorted_tuples", C.c_uint64),
        ("active_tuples", C.c_uint64),
        ("memsets", C.c_uint64),
    ]


# every symbol include/b200sa.h declares: (name, restype, argtypes)
_P = C.c_void_p
ABI = [
    ("b200sa_create", C.c_int, [C.POINTER(_P), C.c_int]),
    ("b200sa_destroy", None, [_P]),
    ("b200sa_release_workspace", C.c_int, [_P]),
    ("b200sa_last_error", C.c_char_p, []),
    ("b200sa_version", C.c_int, []),
    ("b200sa_device_count", C.c_int, []),
    ("b200sa_suffix_array", C.c_int, [_P, _P, C.c_int64, _P]),
    ("b200sa_bwt", C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int32)]),
    ("b200sa_unbwt", C.c_int, [_P, _P, C.c_int64, C.c_int32]),
    ("b200sa_suffix_array_bwt", C.c_int, [_P, _P, C.c_int64, _P, _P, C.POINTER(C.c_int32)]),
    ("b200sa_suffix_array_dev", C.c_int, [_P, _P, C.c_int64, _P, _P]),
    ("b200sa_bwt_dev", C.c_int, [_P, _P, C.c_int64, _P, _P, C.POINTER(C.c_int32), _P]),
    ("b200sa_unbwt_dev", C.c_int, [_P, _P, C.c_int64, C.c_int32, _P, _P]),
    ("b200sa_check_suffix_array_dev", C.c_int, [_P, _P, C.c_int64, _P, C.POINTER(C.c_int64), _P]),
    ("b200sa_suffix_array_u32_dev", C.c_int, [_P, _P, C.c_int64, _P, _P]),
    ("b200sa_bwt_u32_dev", C.c_int, [_P, _P, C.c_int64, _P, _P, C.POINTER(C.c_int64), _P]),
    ("b200sa_check_suffix_array_u32_dev", C.c_int, [_P, _P, C.c_int64, _P, C.POINTER(C.c_int64), _P]),
    ("b200sa_suffix_array_bwt_u32", C.c_int, [_P, _P, C.c_int64, _P, _P, C.POINTER(C.c_int64)]),
    ("b200sa_unbwt_u32_dev", C.c_int, [_P, _P, C.c_int64, C.c_int64, _P, _P]),
    ("b200sa_unbwt_u32", C.c_int, [_P, _P, C.c_int64, C.c_int64]),
    ("b200sa_lcp_u32_dev", C.c_int, [_P, _P, C.c_int64, _P, _P, _P]),
    ("b200sa_lcp_dev", C.c_int, [_P, _P, C.c_int64, _P, _P, _P]),
    ("b200sa_lcp", C.c_int, [_P, _P, C.c_int64, _P, _P, _P]),
    ("b200sa_suffix_array_batch", C.c_int, [_P, _P, _P, C.c_int64, _P]),
    ("b200sa_bwt_batch", C.c_int, [_P, _P, _P, C.c_int64, _P]),
    ("b200sa_unbwt_batch", C.c_int, [_P, _P, _P, C.c_int64, _P]),
    ("b200sa_batch_dev", C.c_int, [_P, _P, _P, C.c_int64, _P, _P, _P, _P]),
    ("b200sa_unbwt_batch_dev", C.c_int, [_P, _P, _P, C.c_int64, _P, _P, _P]),
    ("b200sa_pipeline_create", C.c_int, [C.POINTER(_P), C.c_int, C.c_int]),
    ("b200sa_pipeline_create_devices", C.c_int, [C.POINTER(_P), C.POINTER(C.c_int), C.c_int, C.c_int]),
    ("b200sa_pipeline_destroy", None, [_P]),
    ("b200sa_pipeline_submit_bwt", C.c_int, [_P, _P, _P, C.c_int64, _P, C.POINTER(C.c_int64)]),
    ("b200sa_pipeline_submit_unbwt", C.c_int, [_P, _P, _P, C.c_int64, _P, C.POINTER(C.c_int64)]),
    ("b200sa_pipeline_submit_suffix_array", C.c_int, [_P, _P, _P, C.c_int64, _P, C.POINTER(C.c_int64)]),
    ("b200sa_pipeline_wait", C.c_int, [_P, C.c_int64]),
    ("b200sa_pipeline_drain", C.c_int, [_P]),
    ("b200sa_check_suffix_array", C.c_int, [_P, _P, C.c_int64, _P, C.POINTER(C.c_int64)]),
    ("b200sa_shard_begin", C.c_int, [_P, _P, C.c_int64, _P, C.c_int, C.c_int, C.POINTER(C.c_int64), _P]),
    ("b200sa_shard_round0", C.c_int, [_P, C.c_int64, C.POINTER(C.c_int64), _P]),
    ("b200sa_shard_round", C.c_int, [_P, C.POINTER(C.c_int64), _P]),
    ("b200sa_shard_updates", C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_int64)]),
    ("b200sa_shard_copy_updates", C.c_int, [_P, _P, _P, C.c_int64, _P]),
    ("b200sa_shard_apply_updates", C.c_int, [_P, _P, _P, C.c_int64, _P]),
    ("b200sa_shard_partition", C.c_int, [_P, _P, _P, C.c_int64, C.c_int, _P, _P, C.POINTER(C.c_uint32), _P]),
    ("b200sa_shard_requests", C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int64), _P]),
    ("b200sa_shard_gather_ranks", C.c_int, [_P, _P, C.c_int64, _P, _P]),
    ("b200sa_shard_peer_export", C.c_int, [_P, C.c_int64, _P]),
    ("b200sa_shard_peer_attach", C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int64, _P]),
    ("b200sa_shard_peer_layout", C.c_int, [_P, _P, C.c_int]),
    ("b200sa_shard_peer_scatter", C.c_int, [_P, _P]),
    ("b200sa_shard_peer_apply", C.c_int, [_P, _P]),
    ("b200sa_shard_peer_detach", C.c_int, [_P]),
    ("b200sa_shard_bwt", C.c_int, [_P, C.c_int64, C.c_int64, _P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32), _P]),
    ("b200sa_unbwt_shard_build", C.c_int, [_P, _P, C.c_int64, C.c_int32, C.POINTER(C.c_int64), _P]),
    ("b200sa_unbwt_shard_measure", C.c_int, [_P, C.c_int64, C.c_int64, _P]),
    ("b200sa_unbwt_shard_segments", C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, _P, _P, _P]),
    ("b200sa_unbwt_shard_finish", C.c_int, [_P, C.c_int64, C.c_int64, _P, _P]),
    ("b200sa_group_create", C.c_int, [C.POINTER(_P), C.POINTER(C.c_int), C.c_int]),
    ("b200sa_group_destroy", None, [_P]),
    ("b200sa_group_size", C.c_int, [_P]),
    ("b200sa_group_context", _P, [_P, C.c_int]),
    ("b200sa_group_suffix_array", C.c_int, [_P, _P, C.c_int64, _P]),
    ("b200sa_group_bwt", C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int32)]),
    ("b200sa_group_suffix_array_bwt", C.c_int, [_P, _P, C.c_int64, _P, _P, C.POINTER(C.c_int32)]),
    ("b200sa_group_unbwt", C.c_int, [_P, _P, C.c_int64, C.c_int32]),
    ("b200sa_group_suffix_array_batch", C.c_int, [_P, _P, _P, C.c_int64, _P]),
    ("b200sa_group_bwt_batch", C.c_int, [_P, _P, _P, C.c_int64, _P]),
    ("b200sa_group_unbwt_batch", C.c_int, [_P, _P, _P, C.c_int64, _P]),
    ("b200sa_suffix_array_gpus", C.c_int, [_P, C.c_int64, _P, C.c_int]),
    ("b200sa_bwt_gpus", C.c_int, [_P, C.c_int64, C.POINTER(C.c_int32), C.c_int]),
    ("b200sa_unbwt_gpus", C.c_int, [_P, C.c_int64, C.c_int32, C.c_int]),
    ("b200sa_comm_create_local", C.c_int, [C.POINTER(_P), C.c_int]),
    ("b200sa_comm_create_shm", C.c_int, [C.POINTER(_P), C.c_char_p, C.c_int, C.c_int]),
    ("b200sa_comm_destroy", None, [_P]),
    ("b200sa_comm_set_timeout_ms", C.c_int, [_P, C.c_int]),
    ("b200sa_comm_barrier", C.c_int, [_P]),
    ("b200sa_comm_allreduce_sum", C.c_int, [_P, C.c_int64, C.POINTER(C.c_int64)]),
    ("b200sa_shard_sort", C.c_int, [_P, _P, _P, C.c_int64, _P, _P, C.POINTER(C.c_int64), _P]),
    ("b200sa_shard_unbwt", C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, _P, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _P]),
    ("b200sa_set_profiling", C.c_int, [_P, C.c_int]),
    ("b200sa_profile_reset", C.c_int, [_P]),
    ("b200sa_profile_get", C.c_int, [_P, C.POINTER(_Profile)]),
    ("b200sa_launch_count", C.c_uint64, [_P]),
    ("b200sa_debug_rerank_descriptor", C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]),
    ("b200sa_radix_sort_pairs_dev", C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int), _P]),
]


class Library:
    """A loaded copy of the C ABI.  The product always uses :func:`load_library` (the nvcc-built
    ``msufsort_b200/lib/libb200sa.so``); tests may wrap another CDLL that exports the same ABI."""

    def __init__(self, path_or_cdll):
        self.cdll = C.CDLL(path_or_cdll) if isinstance(path_or_cdll, str) else path_or_cdll
        self.path = path_or_cdll if isinstance(path_or_cdll, str) else getattr(path_or_cdll, "_name", "?")
        for name, restype, argtypes in ABI:
            fn = getattr(self.cdll, name)  # AttributeError if the symbol is missing
            fn.restype = restype
            fn.argtypes = argtypes

    def last_error(self) -> str:
        msg = self.cdll.b200sa_last_error()
        return msg.decode("utf-8", "replace") if msg else ""

    def check(self, rc: int) -> None:
        if rc != 0:
            raise B200SAError(rc, self.last_error())


_lib_lock = threading.Lock()
_lib: Optional[Library] = None


def load_library() -> Library:
    """Loads the CUDA library.  Fails loudly when it has not been built: there is no fallback."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(PRODUCT_LIB):
                raise B200SAError(2, f"{PRODUCT_LIB} is missing — run `make lib` (nvcc, sm_100a); "
                                     "msufsort_b200 has no CPU fallback")
            _lib = Library(PRODUCT_LIB)
        return _lib


CUDA_STREAM_LEGACY = 1  # cudaStreamLegacy: the explicit handle of the legacy default stream


def torch_stream_handle() -> int:
    """cudaStream_t of torch's current stream, suitable for the `stream` argument of the *_dev entry points.
    torch reports the default stream as handle 0, which this ABI reads as "use the context's own stream" (a
    non-blocking stream that is NOT ordered with torch's work); the explicit legacy handle keeps the kernels on
    the stream torch is using, so torch fills, NCCL collectives and torch.cuda.Event timings are ordered with them."""
    import torch
    return int(torch.cuda.current_stream().cuda_stream) or CUDA_STREAM_LEGACY


def _ptr(x) -> Optional[int]:
    """Address of a numpy array / bytearray / torch tensor / raw int."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    if isinstance(x, (bytearray, memoryview)):
        return C.addressof((C.c_char * len(x)).from_buffer(x))
    raise TypeError(f"cannot take the address of {type(x)!r}")


class Engine:
    """One context of the C ABI = one GPU, one stream, one reusable device workspace — the
    analogue of one ``maniscalco::msufsort`` object (msufsort.h:50-75)."""

    def __init__(self, device: int = 0, library: Optional[Library] = None):
        self.lib = library if library is not None else load_library()
        self._ctx = _P()
        self.lib.check(self.lib.cdll.b200sa_create(C.byref(self._ctx), device))
        self.device = device

    def _st(self, stream: Optional[int]):
        """stream argument for the C ABI: an explicit handle is passed through; None means "the stream torch is
        currently using" when torch has a CUDA context (so torch fills, copies, collectives and events stay ordered
        with the kernels), else the context's own stream (NULL)"""
        if stream is not None:
            return stream or None
        import sys
        torch = sys.modules.get("torch")
        if torch is not None and os.path.basename(str(self.lib.path)) == "libb200sa.so":
            try:
                if torch.cuda.is_available() and torch.cuda.is_initialized():
                    return torch_stream_handle()
            except Exception:
                pass
        return None

    def close(self) -> None:
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self.lib.cdll.b200sa_destroy(self._ctx)
            self._ctx = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- host-buffer entry points (what a user of the reference calls) ----------------------
    def make_suffix_array(self, data) -> np.ndarray:
        """SA of ``data`` (bytes-like / uint8 array): int32 array of len(data)+1, SA[0]=len(data)."""
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        if buf.dtype.itemsize != 1:
            raise TypeError("input must be a 1-byte element buffer")
        buf = np.ascontiguousarray(buf)
        n = buf.size
        sa = np.empty(n + 1, dtype=np.int32)
        self.lib.check(self.lib.cdll.b200sa_suffix_array(self._ctx, _ptr(buf) if n else None, n, _ptr(sa)))
        return sa

    def forward_burrows_wheeler_transform(self, data) -> int:
        """In-place BWT of a writable buffer; returns the sentinel index."""
        buf = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
        if not buf.flags.writeable or not buf.flags.c_contiguous or buf.dtype.itemsize != 1:
            raise TypeError("forward_burrows_wheeler_transform needs a writable contiguous byte buffer")
        s = C.c_int32(0)
        self.lib.check(self.lib.cdll.b200sa_bwt(self._ctx, _ptr(buf) if buf.size else None, buf.size, C.byref(s)))
        return int(s.value)

    def reverse_burrows_wheeler_transform(self, data, sentinel_index: int) -> None:
        """In-place inverse BWT."""
        buf = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
        if not buf.flags.writeable or not buf.flags.c_contiguous or buf.dtype.itemsize != 1:
            raise TypeError("reverse_burrows_wheeler_transform needs a writable contiguous byte buffer")
        self.lib.check(self.lib.cdll.b200sa_unbwt(self._ctx, _ptr(buf) if buf.size else None, buf.size, int(sentinel_index)))

    def check_suffix_array(self, data, sa) -> int:
        """Number of offending rows of ``sa`` as the suffix array of ``data`` (0 = correct): the O(n) validator."""
        buf = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        sa = np.ascontiguousarray(sa, dtype=np.int32)
        if sa.size != buf.size + 1:
            raise ValueError("suffix array must have len(data)+1 entries")
        bad = C.c_int64(-1)
        self.lib.check(self.lib.cdll.b200sa_check_suffix_array(self._ctx, _ptr(buf) if buf.size else None, buf.size, _ptr(sa), C.byref(bad)))
        return int(bad.value)

    def suffix_array_and_bwt(self, data):
        """Superset call: one sort, returns (sa, bwt, sentinel_index)."""
        buf = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        n = buf.size
        sa = np.empty(n + 1, dtype=np.int32)
        bwt = np.empty(n, dtype=np.uint8)
        s = C.c_int32(0)
        self.lib.check(self.lib.cdll.b200sa_suffix_array_bwt(self._ctx, _ptr(buf) if n else None, n, _ptr(sa),
                                                             _ptr(bwt) if n else None, C.byref(s)))
        return sa, bwt, int(s.value)

    def make_lcp_array(self, data, sa: Optional[np.ndarray] = None, return_sa: bool = False):
        """LCP array (int32, len(data)+1 entries aligned with the suffix array: lcp[0] = lcp[1] = 0,
        lcp[r] = lcp(SA[r-1], SA[r])) — the demo's ``make_lcp_array`` (main.cpp:141-159; its output[i] is lcp[i+2]).
        ``sa`` is the suffix array of ``data``; when omitted it is computed first."""
        buf = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        if buf.dtype.itemsize != 1:
            raise TypeError("input must be a 1-byte element buffer")
        n = buf.size
        lcp = np.empty(n + 1, dtype=np.int32)
        sa_in = None
        if sa is not None:
            sa_in = np.ascontiguousarray(sa, dtype=np.int32)
            if sa_in.size != n + 1:
                raise ValueError("suffix array must have len(data)+1 entries")
        sa_out = np.empty(n + 1, dtype=np.int32) if (return_sa and sa_in is None) else None
        self.lib.check(self.lib.cdll.b200sa_lcp(self._ctx, _ptr(buf) if n else None, n, _ptr(sa_in), _ptr(sa_out), _ptr(lcp)))
        if return_sa:
            return lcp, (sa_in if sa_in is not None else sa_out)
        return lcp

    # ---- batches of independent blocks (one launch sequence for all of them) ------------------
    @staticmethod
    def _pack_blocks(blocks):
        arrs = [np.ascontiguousarray(np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b).view(np.uint8).ravel()
                for b in blocks]
        offsets = np.zeros(len(arrs) + 1, dtype=np.int64)
        if arrs:
            np.cumsum([a.size for a in arrs], out=offsets[1:])
        packed = np.concatenate(arrs) if arrs else np.empty(0, dtype=np.uint8)
        return np.ascontiguousarray(packed), offsets

    def suffix_array_batch(self, blocks) -> list:
        """Suffix arrays of independent blocks: a list of int32 arrays, block b's has len(block)+1 entries."""
        packed, offsets = self._pack_blocks(blocks)
        count = len(offsets) - 1
        sa = np.empty(int(offsets[-1]) + count, dtype=np.int32)
        self.lib.check(self.lib.cdll.b200sa_suffix_array_batch(self._ctx, _ptr(packed) if packed.size else None, _ptr(offsets), count, _ptr(sa)))
        return [sa[int(offsets[b]) + b: int(offsets[b + 1]) + b + 1] for b in range(count)]

    def bwt_batch(self, blocks):
        """Forward BWT of independent blocks: (list of uint8 arrays, list of sentinel indices)."""
        packed, offsets = self._pack_blocks(blocks)
        count = len(offsets) - 1
        sent = np.zeros(max(count, 1), dtype=np.int32)
        packed = packed.copy()
        self.lib.check(self.lib.cdll.b200sa_bwt_batch(self._ctx, _ptr(packed) if packed.size else None, _ptr(offsets), count, _ptr(sent)))
        return [packed[int(offsets[b]): int(offsets[b + 1])] for b in range(count)], [int(v) for v in sent[:count]]

    def unbwt_batch(self, blocks, sentinel_indices) -> list:
        """Inverse BWT of independent blocks."""
        packed, offsets = self._pack_blocks(blocks)
        count = len(offsets) - 1
        sent = np.ascontiguousarray(np.asarray(list(sentinel_indices) + [0], dtype=np.int32))
        if sent.size != count + 1:
            raise ValueError("one sentinel index per block")
        packed = packed.copy()
        self.lib.check(self.lib.cdll.b200sa_unbwt_batch(self._ctx, _ptr(packed) if packed.size else None, _ptr(offsets), count, _ptr(sent)))
        return [packed[int(offsets[b]): int(offsets[b + 1])] for b in range(count)]

    def batch_dev(self, d_blocks, offsets: np.ndarray, d_bwt=None, d_sa=None, stream: Optional[int] = None) -> np.ndarray:
        """Device-resident batch: returns the sentinel indices (host int32 array)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        count = offsets.size - 1
        sent = np.zeros(max(count, 1), dtype=np.int32)
        self.lib.check(self.lib.cdll.b200sa_batch_dev(self._ctx, _ptr(d_blocks), _ptr(offsets), count, _ptr(d_bwt), _ptr(d_sa), _ptr(sent),
                                                      self._st(stream)))
        return sent[:count]

    def unbwt_batch_dev(self, d_bwt, offsets: np.ndarray, sentinel_indices, d_out, stream: Optional[int] = None) -> None:
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        sent = np.ascontiguousarray(np.asarray(sentinel_indices, dtype=np.int32))
        self.lib.check(self.lib.cdll.b200sa_unbwt_batch_dev(self._ctx, _ptr(d_bwt), _ptr(offsets), offsets.size - 1, _ptr(sent), _ptr(d_out),
                                                            self._st(stream)))

    # ---- raw-pointer variants of the host entry points (pinned buffers in bench.py) ----------
    def suffix_array_ptr(self, text_ptr: int, n: int, sa_ptr: int) -> None:
        self.lib.check(self.lib.cdll.b200sa_suffix_array(self._ctx, text_ptr, n, sa_ptr))

    def bwt_ptr(self, text_ptr: int, n: int) -> int:
        s = C.c_int32(0)
        self.lib.check(self.lib.cdll.b200sa_bwt(self._ctx, text_ptr, n, C.byref(s)))
        return int(s.value)

    def unbwt_ptr(self, bwt_ptr: int, n: int, sentinel_index: int) -> None:
        self.lib.check(self.lib.cdll.b200sa_unbwt(self._ctx, bwt_ptr, n, int(sentinel_index)))

    # ---- device-resident entry points ---------------------------------------------------------
    def suffix_array_dev(self, d_text, n: int, d_sa, stream: Optional[int] = None) -> None:
        self.lib.check(self.lib.cdll.b200sa_suffix_array_dev(self._ctx, _ptr(d_text), n, _ptr(d_sa), self._st(stream)))

    def bwt_dev(self, d_text, n: int, d_bwt, d_sa=None, stream: Optional[int] = None) -> int:
        s = C.c_int32(0)
        self.lib.check(self.lib.cdll.b200sa_bwt_dev(self._ctx, _ptr(d_text), n, _ptr(d_bwt), _ptr(d_sa), C.byref(s), self._st(stream)))
        return int(s.value)

    def unbwt_dev(self, d_bwt, n: int, sentinel_index: int, d_out, stream: Optional[int] = None) -> None:
        self.lib.check(self.lib.cdll.b200sa_unbwt_dev(self._ctx, _ptr(d_bwt), n, int(sentinel_index), _ptr(d_out), self._st(stream)))

    # ---- wide-index superset (uint32 suffix arrays, n up to 2^32 - 8194) -----------------------------
    def suffix_array_u32_dev(self, d_text, n: int, d_sa, stream: Optional[int] = None) -> None:
        self.lib.check(self.lib.cdll.b200sa_suffix_array_u32_dev(self._ctx, _ptr(d_text), n, _ptr(d_sa), self._st(stream)))

    def bwt_u32_dev(self, d_text, n: int, d_bwt, d_sa=None, stream: Optional[int] = None) -> int:
        s = C.c_int64(0)
        self.lib.check(self.lib.cdll.b200sa_bwt_u32_dev(self._ctx, _ptr(d_text), n, _ptr(d_bwt), _ptr(d_sa), C.byref(s), self._st(stream)))
        return int(s.value)

    def check_suffix_array_u32_dev(self, d_text, n: int, d_sa, stream: Optional[int] = None) -> int:
        bad = C.c_int64(-1)
        self.lib.check(self.lib.cdll.b200sa_check_suffix_array_u32_dev(self._ctx, _ptr(d_text), n, _ptr(d_sa), C.byref(bad), self._st(stream)))
        return int(bad.value)

    def suffix_array_and_bwt_u32(self, data):
        """(uint32 SA, BWT bytes, sentinel index) through the wide host entry point."""
        buf = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        n = buf.size
        sa = np.empty(n + 1, dtype=np.uint32)
        bwt = np.empty(n, dtype=np.uint8)
        s = C.c_int64(0)
        self.lib.check(self.lib.cdll.b200sa_suffix_array_bwt_u32(self._ctx, _ptr(buf) if n else None, n, _ptr(sa), _ptr(bwt) if n else None, C.byref(s)))
        return sa, bwt, int(s.value)

    def unbwt_u32_dev(self, d_bwt, n: int, sentinel_index: int, d_out, stream: Optional[int] = None) -> None:
        self.lib.check(self.lib.cdll.b200sa_unbwt_u32_dev(self._ctx, _ptr(d_bwt), n, int(sentinel_index), _ptr(d_out), self._st(stream)))

    def reverse_burrows_wheeler_transform_u32(self, data, sentinel_index: int) -> None:
        """In-place inverse BWT through the wide host entry point (int64 sentinel index)."""
        buf = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
        self.lib.check(self.lib.cdll.b200sa_unbwt_u32(self._ctx, _ptr(buf) if buf.size else None, buf.size, int(sentinel_index)))

    def lcp_u32_dev(self, d_text, n: int, d_sa, d_lcp, stream: Optional[int] = None) -> None:
        self.lib.check(self.lib.cdll.b200sa_lcp_u32_dev(self._ctx, _ptr(d_text), n, _ptr(d_sa), _ptr(d_lcp), self._st(stream)))

    def lcp_dev(self, d_text, n: int, d_sa, d_lcp, stream: Optional[int] = None) -> None:
        self.lib.check(self.lib.cdll.b200sa_lcp_dev(self._ctx, _ptr(d_text), n, _ptr(d_sa), _ptr(d_lcp), self._st(stream)))

    def check_suffix_array_dev(self, d_text, n: int, d_sa, stream: Optional[int] = None) -> int:
        bad = C.c_int64(-1)
        self.lib.check(self.lib.cdll.b200sa_check_suffix_array_dev(self._ctx, _ptr(d_text), n, _ptr(d_sa), C.byref(bad), self._st(stream)))
        return int(bad.value)

    def radix_sort_pairs_dev(self, d_keys, d_keys_alt, d_vals, d_vals_alt, m: int, begin_bit: int, end_bit: int, stream: Optional[int] = None) -> int:
        side = C.c_int(0)
        self.lib.check(self.lib.cdll.b200sa_radix_sort_pairs_dev(self._ctx, _ptr(d_keys), _ptr(d_keys_alt), _ptr(d_vals), _ptr(d_vals_alt),
                                                                 m, begin_bit, end_bit, C.byref(side), self._st(stream)))
        return int(side.value)

    # ---- one text over several GPUs, round loop in C++ (include/b200sa.h "driven from C++") -------------------
    def shard_sort(self, comm: "Comm", d_text, n: int, d_sa, d_bwt=None, stream: Optional[int] = None) -> dict:
        """collective over the ranks of ``comm``; returns this rank's row / byte ranges and counters"""
        info = (C.c_int64 * 8)()
        self.lib.check(self.lib.cdll.b200sa_shard_sort(self._ctx, comm._c, _ptr(d_text), n, _ptr(d_sa), _ptr(d_bwt), info, self._st(stream)))
        keys = ("row_begin", "row_end", "out_begin", "out_end", "sentinel", "rounds", "sent_bytes", "n_local")
        return dict(zip(keys, (int(v) for v in info)))

    def shard_unbwt(self, comm: "Comm", d_bwt, n: int, sentinel_index: int, d_out, gather_all: bool = True, stream: Optional[int] = None):
        """collective inverse BWT; returns the text range [begin, end) this rank owns"""
        b, e = C.c_int64(0), C.c_int64(0)
        self.lib.check(self.lib.cdll.b200sa_shard_unbwt(self._ctx, comm._c, _ptr(d_bwt), n, int(sentinel_index), _ptr(d_out), 1 if gather_all else 0,
                                                        C.byref(b), C.byref(e), self._st(stream)))
        return int(b.value), int(e.value)

    # ---- sharded building blocks (see msufsort_b200/sharded.py) ----------------------------------
    def shard_begin(self, d_text, n: int, d_sa, part: int, nparts: int, stream: Optional[int] = None) -> int:
        out = C.c_int64(0)
        self.lib.check(self.lib.cdll.b200sa_shard_begin(self._ctx, _ptr(d_text), n, _ptr(d_sa), part, nparts, C.byref(out), self._st(stream)))
        return int(out.value)

    def shard_round0(self, slot_base: int, stream: Optional[int] = None) -> int:
        out = C.c_int64(0)
        self.lib.check(self.lib.cdll.b200sa_shard_round0(self._ctx, slot_base, C.byref(out), self._st(stream)))
        return int(out.value)

    def shard_round(self, stream: Optional[int] = None) -> int:
        out = C.c_int64(0)
        self.lib.check(self.lib.cdll.b200sa_shard_round(self._ctx, C.byref(out), self._st(stream)))
        return int(out.value)

    def shard_updates(self):
        """(idx_ptr, rank_ptr, count) of the ISA updates produced by the last step"""
        pi, pr, cnt = _P(), _P(), C.c_int64(0)
        self.lib.check(self.lib.cdll.b200sa_shard_updates(self._ctx, C.byref(pi), C.byref(pr), C.byref(cnt)))
        return (pi.value or 0), (pr.value or 0), int(cnt.value)

    def shard_copy_updates(self, d_idx_dst, d_rank_dst, capacity: int, stream: Optional[int] = None) -> None:
        self.lib.check(self.lib.cdll.b200sa_shard_copy_updates(self._ctx, _ptr(d_idx_dst), _ptr(d_rank_dst), capacity, self._st(stream)))

    def shard_apply_updates(self, d_idx, d_rank, count: int, stream: Optional[int] = None) -> None:
        self.lib.check(self.lib.cdll.b200sa_shard_apply_updates(self._ctx, _ptr(d_idx), _ptr(d_rank), count, self._st(stream)))

    def shard_partition(self, d_keys, d_vals, count: int, shift: int, d_keys_out, d_vals_out, stream: Optional[int] = None) -> list:
        counts = (C.c_uint32 * 256)()
        self.lib.check(self.lib.cdll.b200sa_shard_partition(self._ctx, _ptr(d_keys), _ptr(d_vals), count, shift, _ptr(d_keys_out),
                                                            _ptr(d_vals_out), counts, self._st(stream)))
        return list(counts)

    def shard_requests(self, d_pos_out, capacity: int, stream: Optional[int] = None) -> int:
        cnt = C.c_int64(0)
        self.lib.check(self.lib.cdll.b200sa_shard_requests(self._ctx, _ptr(d_pos_out), capacity, C.byref(cnt), self._st(stream)))
        return int(cnt.value)

    def shard_gather_ranks(self, d_pos, count: int, d_out, stream: Optional[int] = None) -> None:
        self.lib.check(self.lib.cdll.b200sa_shard_gather_ranks(self._ctx, _ptr(d_pos), count, _ptr(d_out), self._st(stream)))

    def shard_peer_export(self, n: int) -> bytes:
        buf = (C.c_uint8 * 128)()
        self.lib.check(self.lib.cdll.b200sa_shard_peer_export(self._ctx, n, buf))
        return bytes(buf)

    def shard_peer_attach(self, part: int, nparts: int, shift: int, n: int, handles: bytes) -> None:
        assert len(handles) == 128 * nparts
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        self.lib.check(self.lib.cdll.b200sa_shard_peer_attach(self._ctx, part, nparts, shift, n, buf))

    def shard_peer_layout(self, counts) -> None:
        arr = (C.c_int64 * len(counts))(*[int(c) for c in counts])
        self.lib.check(self.lib.cdll.b200sa_shard_peer_layout(self._ctx, arr, len(counts)))

    def shard_peer_apply(self, stream: Optional[int] = None) -> None:
        self.lib.check(self.lib.cdll.b200sa_shard_peer_apply(self._ctx, self._st(stream)))

    def shard_peer_scatter(self, stream: Optional[int] = None) -> None:
        self.lib.check(self.lib.cdll.b200sa_shard_peer_scatter(self._ctx, self._st(stream)))

    def shard_peer_detach(self) -> None:
        self.lib.check(self.lib.cdll.b200sa_shard_peer_detach(self._ctx))

    def shard_bwt(self, row_begin: int, row_end: int, d_bwt, stream: Optional[int] = None):
        ob, oe, s = C.c_int64(0), C.c_int64(0), C.c_int32(0)
        self.lib.check(self.lib.cdll.b200sa_shard_bwt(self._ctx, row_begin, row_end, _ptr(d_bwt), C.byref(ob), C.byref(oe), C.byref(s), self._st(stream)))
        return int(ob.value), int(oe.value), int(s.value)

    def unbwt_shard_build(self, d_bwt, n: int, sentinel_index: int, stream: Optional[int] = None) -> int:
        w = C.c_int64(0)
        self.lib.check(self.lib.cdll.b200sa_unbwt_shard_build(self._ctx, _ptr(d_bwt), n, int(sentinel_index), C.byref(w), self._st(stream)))
        return int(w.value)

    def unbwt_shard_measure(self, w_begin: int, w_end: int, stream: Optional[int] = None) -> None:
        self.lib.check(self.lib.cdll.b200sa_unbwt_shard_measure(self._ctx, w_begin, w_end, self._st(stream)))

    def unbwt_shard_segments(self, direction: int, w_begin: int, w_end: int, d_len, d_next, stream: Optional[int] = None) -> None:
        self.lib.check(self.lib.cdll.b200sa_unbwt_shard_segments(self._ctx, direction, w_begin, w_end, _ptr(d_len), _ptr(d_next), self._st(stream)))

    def unbwt_shard_finish(self, w_begin: int, w_end: int, d_text_out, stream: Optional[int] = None) -> None:
        self.lib.check(self.lib.cdll.b200sa_unbwt_shard_finish(self._ctx, w_begin, w_end, _ptr(d_text_out), self._st(stream)))

    # ---- instrumentation ----------------------------------------------------------------------
    def set_profiling(self, enabled: bool) -> None:
        self.lib.check(self.lib.cdll.b200sa_set_profiling(self._ctx, 1 if enabled else 0))

    def profile_reset(self) -> None:
        self.lib.check(self.lib.cdll.b200sa_profile_reset(self._ctx))

    def profile(self) -> dict:
        p = _Profile()
        self.lib.check(self.lib.cdll.b200sa_profile_get(self._ctx, C.byref(p)))
        out = {"rounds": p.rounds, "sort_passes": p.sort_passes, "sorted_tuples": p.sorted_tuples,
               "active_tuples": p.active_tuples, "memsets": p.memsets, "phases": {}}
        for name, i in PHASES.items():
            out["phases"][name] = {"ms": p.ms[i], "launches": p.launches[i], "alg_bytes": p.alg_bytes[i]}
        return out

    def launch_count(self) -> int:
        return int(self.lib.cdll.b200sa_launch_count(self._ctx))

    def release_workspace(self) -> None:
        self.lib.check(self.lib.cdll.b200sa_release_workspace(self._ctx))


class Comm:
    """Control plane of a sharded run (include/b200sa.h b200sa_comm_*): barriers and sums through memory all ranks map."""

    def __init__(self, handle, library: Library):
        self._c = handle
        self.lib = library

    @classmethod
    def local(cls, nranks: int, library: Optional[Library] = None) -> list:
        """``nranks`` handles for the threads of this process"""
        lib = library if library is not None else load_library()
        arr = (_P * nranks)()
        lib.check(lib.cdll.b200sa_comm_create_local(arr, nranks))
        return [cls(_P(arr[r]), lib) for r in range(nranks)]

    @classmethod
    def shared_memory(cls, name: str, rank: int, nranks: int, library: Optional[Library] = None) -> "Comm":
        """the processes of one node; ``name`` = a fresh POSIX shared-memory name starting with '/'"""
        lib = library if library is not None else load_library()
        h = _P()
        lib.check(lib.cdll.b200sa_comm_create_shm(C.byref(h), name.encode(), rank, nranks))
        return cls(h, lib)

    def set_timeout_ms(self, timeout_ms: int) -> None:
        self.lib.check(self.lib.cdll.b200sa_comm_set_timeout_ms(self._c, int(timeout_ms)))

    def barrier(self) -> None:
        self.lib.check(self.lib.cdll.b200sa_comm_barrier(self._c))

    def allreduce_sum(self, value: int) -> int:
        out = C.c_int64(0)
        self.lib.check(self.lib.cdll.b200sa_comm_allreduce_sum(self._c, int(value), C.byref(out)))
        return int(out.value)

    def close(self) -> None:
        if getattr(self, "_c", None):
            self.lib.cdll.b200sa_comm_destroy(self._c)
            self._c = None


class Group:
    """Several GPUs behind the reference's three calls (include/b200sa.h b200sa_group_*): one host thread and one context
    per listed device, one text sharded over them.  ``devices`` may list a device more than once."""

    def __init__(self, devices, library: Optional[Library] = None):
        self.lib = library if library is not None else load_library()
        devs = (C.c_int * len(devices))(*devices)
        self._g = _P()
        self.lib.check(self.lib.cdll.b200sa_group_create(C.byref(self._g), devs, len(devices)))
        self.size = len(devices)

    def make_suffix_array(self, data) -> np.ndarray:
        buf = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        sa = np.empty(buf.size + 1, dtype=np.int32)
        self.lib.check(self.lib.cdll.b200sa_group_suffix_array(self._g, _ptr(buf) if buf.size else None, buf.size, _ptr(sa)))
        return sa

    def forward_burrows_wheeler_transform(self, buf: np.ndarray) -> int:
        s = C.c_int32(0)
        self.lib.check(self.lib.cdll.b200sa_group_bwt(self._g, _ptr(buf) if buf.size else None, buf.size, C.byref(s)))
        return int(s.value)

    def suffix_array_and_bwt(self, data):
        buf = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        sa = np.empty(buf.size + 1, dtype=np.int32)
        bwt = np.empty(buf.size, dtype=np.uint8)
        s = C.c_int32(0)
        self.lib.check(self.lib.cdll.b200sa_group_suffix_array_bwt(self._g, _ptr(buf) if buf.size else None, buf.size, _ptr(sa),
                                                                   _ptr(bwt) if buf.size else None, C.byref(s)))
        return sa, bwt, int(s.value)

    def reverse_burrows_wheeler_transform(self, buf: np.ndarray, sentinel_index: int) -> None:
        self.lib.check(self.lib.cdll.b200sa_group_unbwt(self._g, _ptr(buf) if buf.size else None, buf.size, int(sentinel_index)))

    # ---- batches of independent blocks, one run of blocks per GPU (include/b200sa.h b200sa_group_*_batch) --------------------
    def suffix_array_batch(self, blocks) -> list:
        packed, offsets = Engine._pack_blocks(blocks)
        count = len(offsets) - 1
        sa = np.empty(int(offsets[-1]) + count, dtype=np.int32)
        self.lib.check(self.lib.cdll.b200sa_group_suffix_array_batch(self._g, _ptr(packed) if packed.size else None, _ptr(offsets), count, _ptr(sa)))
        return [sa[int(offsets[b]) + b: int(offsets[b + 1]) + b + 1] for b in range(count)]

    def bwt_batch(self, blocks):
        packed, offsets = Engine._pack_blocks(blocks)
        count = len(offsets) - 1
        sent = np.zeros(max(count, 1), dtype=np.int32)
        packed = packed.copy()
        self.lib.check(self.lib.cdll.b200sa_group_bwt_batch(self._g, _ptr(packed) if packed.size else None, _ptr(offsets), count, _ptr(sent)))
        return [packed[int(offsets[b]): int(offsets[b + 1])] for b in range(count)], [int(v) for v in sent[:count]]

    def unbwt_batch(self, blocks, sentinel_indices) -> list:
        packed, offsets = Engine._pack_blocks(blocks)
        count = len(offsets) - 1
        sent = np.ascontiguousarray(np.asarray(list(sentinel_indices) + [0], dtype=np.int32))
        if sent.size != count + 1:
            raise ValueError("one sentinel index per block")
        packed = packed.copy()
        self.lib.check(self.lib.cdll.b200sa_group_unbwt_batch(self._g, _ptr(packed) if packed.size else None, _ptr(offsets), count, _ptr(sent)))
        return [packed[int(offsets[b]): int(offsets[b + 1])] for b in range(count)]

    def launch_count(self) -> int:
        return sum(int(self.lib.cdll.b200sa_launch_count(self.lib.cdll.b200sa_group_context(self._g, r))) for r in range(self.size))

    def close(self) -> None:
        if getattr(self, "_g", None):
            self.lib.cdll.b200sa_group_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Pipeline:
    """Streaming pipeline over batches: ``depth`` contexts behind one queue, so that the transfers of one batch
    overlap the sort of another (b200sa_pipeline_*).  Buffers passed to submit must stay alive and untouched
    until ``wait``; numpy arrays are kept referenced here until then."""

    def __init__(self, device: int = 0, depth: int = 2, library: Optional[Library] = None, devices=None):
        """devices: a list of CUDA devices — ``depth`` contexts on EACH of them serve the one queue (b200sa_pipeline_create_devices)"""
        self.lib = library if library is not None else load_library()
        self._p = _P()
        if devices is not None:
            devs = (C.c_int * len(devices))(*devices)
            self.lib.check(self.lib.cdll.b200sa_pipeline_create_devices(C.byref(self._p), devs, len(devices), depth))
        else:
            self.lib.check(self.lib.cdll.b200sa_pipeline_create(C.byref(self._p), device, depth))
        self._keep = {}

    def submit_bwt(self, packed: np.ndarray, offsets: np.ndarray, sentinels_out: np.ndarray) -> int:
        """packed: writable uint8 array (transformed in place); offsets: int64[count+1]; sentinels_out: int32[count]"""
        t = C.c_int64(0)
        self.lib.check(self.lib.cdll.b200sa_pipeline_submit_bwt(self._p, _ptr(packed), _ptr(offsets), offsets.size - 1, _ptr(sentinels_out), C.byref(t)))
        self._keep[t.value] = (packed, offsets, sentinels_out)
        return int(t.value)

    def submit_unbwt(self, packed: np.ndarray, offsets: np.ndarray, sentinels: np.ndarray) -> int:
        t = C.c_int64(0)
        self.lib.check(self.lib.cdll.b200sa_pipeline_submit_unbwt(self._p, _ptr(packed), _ptr(offsets), offsets.size - 1, _ptr(sentinels), C.byref(t)))
        self._keep[t.value] = (packed, offsets, sentinels)
        return int(t.value)

    def submit_suffix_array(self, packed: np.ndarray, offsets: np.ndarray, sa_out: np.ndarray) -> int:
        t = C.c_int64(0)
        self.lib.check(self.lib.cdll.b200sa_pipeline_submit_suffix_array(self._p, _ptr(packed), _ptr(offsets), offsets.size - 1, _ptr(sa_out), C.byref(t)))
        self._keep[t.value] = (packed, offsets, sa_out)
        return int(t.value)

    def submit_bwt_ptr(self, blocks_ptr: int, offsets: np.ndarray, sentinels_ptr: int) -> int:
        t = C.c_int64(0)
        self.lib.check(self.lib.cdll.b200sa_pipeline_submit_bwt(self._p, blocks_ptr, _ptr(offsets), offsets.size - 1, sentinels_ptr, C.byref(t)))
        self._keep[t.value] = (offsets,)
        return int(t.value)

    def wait(self, ticket: int) -> None:
        rc = self.lib.cdll.b200sa_pipeline_wait(self._p, ticket)
        self._keep.pop(ticket, None)
        self.lib.check(rc)

    def drain(self) -> None:
        rc = self.lib.cdll.b200sa_pipeline_drain(self._p)
        self._keep.clear()
        self.lib.check(rc)

    def close(self) -> None:
        if getattr(self, "_p", None) is not None and self._p:
            self.lib.cdll.b200sa_pipeline_destroy(self._p)
            self._p = _P()
            self._keep.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# ---- module-level functions with the reference's names ---------------------------------------
_default_engine: Optional[Engine] = None


def _engine() -> Engine:
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(int(os.environ.get("MSUFSORT_DEVICE", "0")))
    return _default_engine


def make_suffix_array(data, num_threads: int = 1) -> np.ndarray:
    return _engine().make_suffix_array(data)


def forward_burrows_wheeler_transform(data, num_threads: int = 1) -> int:
    return _engine().forward_burrows_wheeler_transform(data)


def reverse_burrows_wheeler_transform(data, sentinel_index: int, num_threads: int = 1) -> None:
    _engine().reverse_burrows_wheeler_transform(data, sentinel_index)


def make_lcp_array(data, suffix_array=None, num_threads: int = 1) -> np.ndarray:
    return _engine().make_lcp_array(data, suffix_array)
