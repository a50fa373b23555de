"""GPU tier, batches of independent blocks over the contexts of a group (b200sa_group_*_batch).
These entry points were added after the round's last GPU run: the file sorts last so that the GPU tier reaches it after
everything that had already been verified on a B200.  The emulator tier of the same code is tests/test_group.py."""
import os

import numpy as np
import pytest

from cases import gen
from msufsort_b200.api import Group
from conftest import ROOT
from test_cpp_ext import GROUP_SRC, _build, check_group_caller
from test_group import _batch_blocks, _check_batch


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_gpu_group_batch_contexts_on_one_device(oracle, world):
    rng = np.random.default_rng(100 + world)
    g = Group([0] * world)
    try:
        before = g.launch_count()
        _check_batch(g, oracle, _batch_blocks(rng, 6 << 20, 96))
        _check_batch(g, oracle, [gen("markov3", 1 << 20)] + [gen("rand", 1000)] * 7)
        assert g.launch_count() > before
    finally:
        g.close()


@pytest.mark.gpu
def test_gpu_group_batch_one_context_per_gpu(oracle):
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("one context per GPU needs at least 2 GPUs (test_gpu_group_batch_contexts_on_one_device covers the path)")
    g = Group(list(range(min(ng, 8))))
    try:
        _check_batch(g, oracle, _batch_blocks(np.random.default_rng(7), 32 << 20, 512))
    finally:
        g.close()


@pytest.mark.gpu
def test_gpu_cpp_group_caller(tmp_path):
    """the C++ extension header's gpu_group (tests/cpp/ext_group_test.cpp) against the product library: three contexts on cuda:0"""
    libdir = os.path.join(ROOT, "msufsort_b200", "lib")
    exe = _build(os.path.join(ROOT, "tests", "cpp", "ext_group_test"), libdir, "b200sa", GROUP_SRC)
    check_group_caller(exe, tmp_path, (4 << 20) + 13, 50000, "0,0,0")


@pytest.mark.gpu
def test_gpu_pipeline_over_listed_devices(oracle):
    """b200sa_pipeline_create_devices: two contexts on every listed device behind one queue (here cuda:0 listed twice, and every
    GPU of the box when there are several)"""
    import torch
    from msufsort_b200.api import load_library
    from test_batch import _pipeline_roundtrip
    lib = load_library()
    _pipeline_roundtrip(lib, oracle, nbatches=6, blocks_per_batch=8, block_len=200000, depth=2, devices=[0, 0])
    ng = torch.cuda.device_count()
    if ng >= 2:
        _pipeline_roundtrip(lib, oracle, nbatches=2 * ng, blocks_per_batch=8, block_len=200000, depth=2, devices=list(range(min(ng, 8))))
