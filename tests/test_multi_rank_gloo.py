"""CPU, world_size 2 over gloo: the N>1 plumbing (index sharding, idx bases, max-over-ranks timing,
concatenation in pair order).  The per-rank alignment itself is stood in by the CPU oracle here —
this test is about the host-side logic around the GPU call, which needs no GPU."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import aim_b200 as A
from aim_b200 import shard
from oracle import oracle as O

P = 600
KW = dict(max_score=30, read_size=168, backtrace=True, reduce=True)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first = shard.weak_first_pair(rank, P)
    plen, tlen, pats, txts = A.generate_pairs(4, P, 150, 0.04, 168, first_pair=first, nthreads=1)
    res, _ = O.align("wfa", plen, tlen, pats, txts, **KW)
    out = np.zeros(P, A.RESULT_DTYPE)
    out["score"], out["idx"] = res["score"], np.arange(P) + first
    t = shard.max_over_ranks(1.0 + rank)  # slowest rank defines the step time
    allres = shard.gather_in_pair_order(out, world, rank)
    dist.barrier()
    if rank == 0:
        q.put((t, allres["score"].tolist(), allres["idx"].tolist()))
    dist.destroy_process_group()


def test_two_ranks_shard_and_concatenate():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    t, scores, idx = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t == 2.0
    assert idx == list(range(2 * P))
    plen, tlen, pats, txts = A.generate_pairs(4, 2 * P, 150, 0.04, 168, nthreads=1)
    res, _ = O.align("wfa", plen, tlen, pats, txts, **KW)
    assert scores == res["score"].tolist()


@pytest.mark.parametrize("total,world", [(10, 3), (8, 8), (5, 8), (1000003, 4), (0, 2)])
def test_rank_slices_tile_the_batch(total, world):
    got = [shard.rank_slice(total, world, r) for r in range(world)]
    pos = 0
    for first, cnt in got:
        assert first == min(pos, total) and cnt >= 0
        pos += cnt
    assert pos == total
