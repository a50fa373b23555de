"""GPU tier, N > 1: one text sharded over 2 (or more) B200s with NCCL; skipped on single-GPU boxes."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, family, n, q, isa="owner"):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from cases import gen
        from msufsort_b200.api import Engine
        from msufsort_b200.sharded import ShardedSorter
        eng = Engine(rank)
        x = gen(family, n)
        d_text = torch.from_numpy(x.copy()).cuda()
        sorter = ShardedSorter(eng, isa=isa)
        res = sorter.suffix_array_bwt(d_text)
        sa = sorter.gather_sa(res).cpu().numpy()
        bwt = sorter.gather_bwt(res).cpu().numpy()
        back = sorter.inverse_bwt(sorter.gather_bwt(res), res.sentinel)
        assert bool((back.cpu() == d_text.cpu()).all()), "sharded inverse BWT did not restore the text"
        if isa == "peer":
            # without the final all-gather every rank holds its slice of the text; the slices tile [0, n)
            out, b, e = sorter.inverse_bwt(sorter.gather_bwt(res), res.sentinel, gather_all=False)
            assert bool(torch.equal(out[b:e], d_text[b:e]))
            tot = torch.tensor([e - b], dtype=torch.int64, device="cuda")
            dist.all_reduce(tot)
            assert int(tot.item()) == n
        counts = sorter.owned_counts(res)
        if rank == 0:
            q.put((sa, bwt, res.sentinel, counts, res.rounds))
        dist.barrier()
        sorter.close()
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("isa", ["peer", "owner", "replicated"])
@pytest.mark.parametrize("family,n", [("markov3", (1 << 22) + 5), ("acgt_rep", 1 << 22), ("rand", 1 << 20), ("abcabca", 1 << 20),
                                      ("fib", 1 << 19), ("zeros", 1 << 18)])
def test_sharded_nccl_matches_oracle(oracle, family, n, isa):
    import torch.multiprocessing as mp
    from cases import gen
    world = min(_ngpus(), 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, family, n, q, isa)) for r in range(world)]
    for p in procs:
        p.start()
    sa, bwt, sentinel, counts, rounds = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    x = gen(family, n)
    want = oracle.sa(x)
    assert sum(counts) == n
    assert np.array_equal(sa, want)
    wb, ws = oracle.bwt_from_sa(x, want)
    assert sentinel == ws and np.array_equal(bwt, wb)
