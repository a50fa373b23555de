"""CPU tier: the N > 1 path — world_size-2 and -3 runs over gloo, kernels under the emulator build.
Checks that the sharded driver (msufsort_b200/sharded.py) reproduces the oracle's SA and BWT."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, family, n, q, isa="owner"):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    # drive the bucketed ISA update even at these small sizes
    os.environ["B200SA_ISA_DIRECT_BYTES"] = "0"
    os.environ["B200SA_ISA_MIN_UPDATES"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cases import gen
        from msufsort_b200.api import Engine, Library
        from msufsort_b200.sharded import ShardedSorter
        eng = Engine(0, library=Library(os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so")))
        x = gen(family, n)
        d_text = torch.from_numpy(x.copy())
        sorter = ShardedSorter(eng, isa=isa)
        res = sorter.suffix_array_bwt(d_text)
        sa = sorter.gather_sa(res).numpy()
        bwt = sorter.gather_bwt(res).numpy()
        back = sorter.inverse_bwt(sorter.gather_bwt(res), res.sentinel)
        assert bool((back.cpu() == d_text.cpu()).all()), "sharded inverse BWT did not restore the text"
        if rank == 0:
            q.put((sa, bwt, res.sentinel, sorter.owned_counts(res), res.rounds))
        dist.barrier()
        eng.close()
    finally:
        dist.destroy_process_group()


CASES = [(w, isa, f, n)
         for w, isa in ((2, "owner"), (2, "replicated"), (3, "owner"))
         for f, n in (("markov3", 40000), ("acgt_rep", 30011), ("abcabca", 9000), ("zeros", 3000), ("fib", 10000), ("rand", 3))
         if not (w == 3 and f in ("acgt_rep", "abcabca"))]


@pytest.mark.parametrize("world,isa,family,n", CASES)
def test_sharded_matches_oracle(oracle, world, family, n, isa):
    from cases import gen
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, family, n, q, isa)) for r in range(world)]
    for p in procs:
        p.start()
    sa, bwt, sentinel, counts, rounds = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    x = gen(family, n)
    want = oracle.sa(x)
    assert sum(counts) == n
    assert np.array_equal(sa, want), (family, n, world)
    wb, ws = oracle.bwt_from_sa(x, want)
    assert sentinel == ws and np.array_equal(bwt, wb)


# ---- ISA in peer memory (isa="peer"): G contexts of ONE process driven in lock step ------------------------------
# CUDA IPC needs real GPUs; under the emulator the "handle" carries the pointer, so the contexts of one process can
# map each other's ISA arrays.  This exercises the engine side of the protocol (export / attach / peer reads in the
# doubling rounds / peer scatter / BWT through the shard of GPU 0); the torch.distributed side runs in the GPU tier.
def _peer_lockstep(lib, oracle, x, world):
    """G contexts of one process run the peer-ISA protocol in lock step; returns nothing, asserts SA / BWT / sentinel."""
    from msufsort_b200.api import Engine
    n = x.size
    engs = [Engine(0, library=lib) for _ in range(world)]
    try:
        per = (n + world - 1) // world
        shift = max(0, (per - 1).bit_length())
        handles = b"".join(e.shard_peer_export(n) for e in engs)
        for g, e in enumerate(engs):
            e.shard_peer_attach(g, world, shift, n, handles)
        sas = [np.zeros(n + 1, dtype=np.int32) for _ in range(world)]
        counts = [e.shard_begin(x, n, sas[g], g, world) for g, e in enumerate(engs)]
        assert sum(counts) == n
        bases = [sum(counts[:g]) for g in range(world)]
        for e in engs:
            e.shard_peer_layout(counts)
        m = [e.shard_round0(bases[g]) for g, e in enumerate(engs)]
        rounds = 1
        while True:
            for e in engs:            # write phase: everyone has finished reading
                e.shard_peer_scatter()
            for e in engs:            # all sends have landed
                e.shard_peer_apply()
            if sum(m) == 0:
                break
            m = [e.shard_round() for e in engs]   # read phase
            rounds += 1
            assert rounds < 64
        want = oracle.sa(x)
        sa = np.zeros(n + 1, dtype=np.int32)
        bwt = np.zeros(n, dtype=np.uint8)
        sentinel = None
        for g, e in enumerate(engs):
            rb = 0 if g == 0 else bases[g] + 1
            re = bases[g] + counts[g] + 1
            sa[rb:re] = sas[g][rb:re]
            part = np.zeros(n, dtype=np.uint8)
            ob, oe, s = e.shard_bwt(rb, re, part)
            bwt[ob:oe] = part[ob:oe]
            assert sentinel in (None, s)
            sentinel = s
        assert np.array_equal(sa, want), (n, world)
        wb, ws = oracle.bwt_from_sa(x, want)
        assert sentinel == ws and np.array_equal(bwt, wb)
        return engs
    except Exception:
        for e in engs:
            e.close()
        raise


@pytest.mark.parametrize("world", [2, 3, 5])
@pytest.mark.parametrize("family,n", [("markov3", 40000), ("acgt_rep", 30011), ("abcabca", 9000), ("zeros", 3000), ("fib", 10000),
                                      ("periodic7", 5000), ("rand", 64)])
def test_peer_isa_lockstep(oracle, world, family, n, monkeypatch):
    from cases import gen
    from msufsort_b200.api import Library
    monkeypatch.setenv("B200SA_GROUPSORT_TINY", "4")      # reach the CTA and the radix paths at these sizes too
    monkeypatch.setenv("B200SA_GROUPSORT_MEDIUM", "64")
    if world != 3:                                        # bucketed ISA apply (world 3 keeps the direct one)
        monkeypatch.setenv("B200SA_ISA_DIRECT_BYTES", "0")
        monkeypatch.setenv("B200SA_ISA_MIN_UPDATES", "1")
    lib = Library(os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so"))
    engs = _peer_lockstep(lib, oracle, gen(family, n), world)
    try:
        # a second text of another size through the same contexts: mappings are re-established
        x2 = gen("markov3", n // 2 + 7)
        handles = b"".join(e.shard_peer_export(x2.size) for e in engs)
        for g, e in enumerate(engs):
            e.shard_peer_attach(g, world, max(0, ((x2.size + world - 1) // world - 1).bit_length()), x2.size, handles)
        for e in engs:
            e.shard_peer_detach()
    finally:
        for e in engs:
            e.close()


def test_peer_isa_lockstep_fuzz(oracle):
    """random small texts over tiny alphabets (deep rounds, huge groups, empty key ranges) on 2..6 GPUs-in-one-process"""
    from hypothesis import HealthCheck, given, settings, strategies as st
    from msufsort_b200.api import Library
    lib = Library(os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so"))

    @settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
    @given(text=st.lists(st.integers(0, 2), min_size=2, max_size=400), world=st.integers(2, 6))
    def run(text, world):
        engs = _peer_lockstep(lib, oracle, np.array(text, dtype=np.uint8), world)
        for e in engs:
            e.close()

    run()


# ---- the control plane of the C++ round loop between PROCESSES: a POSIX shared-memory segment (csrc/comm.cuh) ----------
def _comm_worker(rank, world, name, q, fail):
    from msufsort_b200.api import B200SAError, Comm, Library
    lib = Library(os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so"))
    c = Comm.shared_memory(name, rank, world, library=lib)   # the ranks meet here once (default deadline: slow process starts are fine)
    try:
        if fail:
            c.set_timeout_ms(300)
            if rank == 1:
                q.put((rank, "left"))        # never shows up at the barrier
                return
            try:
                c.barrier()
                q.put((rank, "no error"))
            except B200SAError as e:
                q.put((rank, "code %d" % e.code))
            return
        for i in range(3000):
            assert c.allreduce_sum(rank * 1000 + i) == sum(r * 1000 + i for r in range(world))
            if i % 7 == 0:
                c.barrier()
        q.put((rank, "ok"))
    finally:
        c.close()


@pytest.mark.parametrize("world", [2, 3])
def test_shm_comm_between_processes(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    name = "/b200sa_test_%d_%s" % (os.getpid(), os.urandom(4).hex())
    procs = [ctx.Process(target=_comm_worker, args=(r, world, name, q, False)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [(r, "ok") for r in range(world)]


def test_shm_comm_times_out_instead_of_hanging():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    name = "/b200sa_test_%d_%s" % (os.getpid(), os.urandom(4).hex())
    procs = [ctx.Process(target=_comm_worker, args=(r, 2, name, q, True)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert got[1] == "left" and got[0] == "code 6"   # B200SA_ECOMM


# ---- the same control plane between THREADS (b200sa_group_*), under ThreadSanitizer ------------------------------------------
@pytest.mark.parametrize("sanitize", [False, True])
def test_comm_stress_between_threads(sanitize, tmp_path):
    """tests/cpp/comm_stress.cpp: barriers, sums and all-gathers in a row on 2, 3 and 8 threads, plain stores that only the
    barrier publishes, and a rank that fails while its peers wait (they leave with B200SA_ECOMM).  The sanitized build
    reports a data race or a missing acquire / release pair in comm.cuh as a failure."""
    import shutil
    import subprocess
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "comm_stress")
    cmd = [cxx, "-std=c++17", "-O1", "-g", "-DB200SA_EMU", "-I" + os.path.join(ROOT, "tests", "emu"),
           "-I" + os.path.join(ROOT, "msufsort_b200", "csrc"), os.path.join(ROOT, "tests", "cpp", "comm_stress.cpp"), "-pthread", "-o", exe]
    if sanitize:
        cmd.insert(1, "-fsanitize=thread")
    built = subprocess.run(cmd, capture_output=True, text=True)
    if built.returncode != 0 and sanitize and "tsan" in built.stderr:
        pytest.skip("ThreadSanitizer runtime not installed")
    assert built.returncode == 0, built.stderr[-2000:]
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1")
    out = subprocess.run([exe, "150" if sanitize else "2000"], capture_output=True, text=True, timeout=600, env=env)
    if sanitize and "unexpected memory mapping" in out.stderr:
        pytest.skip("ThreadSanitizer cannot run under this kernel's address-space layout")
    assert out.returncode == 0 and "comm_stress: ok" in out.stdout, (out.stdout + out.stderr)[-3000:]
