"""CPU tests of the multi-GPU host logic (world_size 2, gloo): row sharding + shard gather layout."""
import os
import socket

import numpy as np
import pytest

from macrodna_b200 import dist as mdist


def test_row_shard_covers_rows_once():
    for M in (1, 2, 7, 50, 50000, 50001):
        for P in (1, 2, 3, 4, 8):
            spans = [mdist.row_shard(M, P, r) for r in range(P)]
            rows = [i for lo, hi in spans for i in range(lo, hi)] if M < 1000 else None
            assert spans[0][0] == 0 and spans[-1][1] == M
            assert all(spans[i][1] == spans[i + 1][0] for i in range(P - 1))
            per = -(-M // P)
            assert all(hi - lo == per for lo, hi in spans[:-1] if hi < M)
            if rows is not None:
                assert rows == list(range(M))
    assert [mdist.replicate_owner(r, 4) for r in range(6)] == [0, 1, 2, 3, 0, 1]


def _worker(rank, world, port, M, cols, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(M * cols, dtype=torch.float64).reshape(M, cols)
    lo, hi = mdist.row_shard(M, world, rank)
    per = -(-M // world)
    local = torch.full((per, cols), -1.0, dtype=torch.float64)
    local[: hi - lo] = full[lo:hi]
    got = mdist.gather_rows(local, M, world)
    q.put((rank, bool(torch.equal(got, full))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("M", [10, 11])
def test_gather_rows_world2_gloo(M):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, M, 6, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_replicate_partition_and_accuracy():
    """Config 4 is 'replicas only': replicate r belongs to rank r mod P, every replicate to exactly one rank; the
    per-replicate accuracy is the reference's (clonal_proportions_resampling.py:191-201)."""
    import numpy as np

    from macrodna_b200 import dist as mdist

    for world in (1, 2, 4, 8):
        owners = [mdist.replicate_owner(r, world) for r in range(1000)]
        counts = np.bincount(owners, minlength=world)
        assert counts.sum() == 1000 and counts.max() - counts.min() <= 1
    rna_clone = np.array([0, 0, 1, 1, 2])
    dna_clone = np.array([0, 1, 2, 2])
    cols = np.array([3, 0, 1])            # the replicate holds DNA cells 3, 0, 1 (clones 2, 0, 1)
    assign = np.array([1, 0, 2, 2, 0])    # positions in cols -> clones 0, 2, 1, 1, 2
    assert mdist.replicate_accuracy(assign, cols, rna_clone, dna_clone) == 4 / 5
