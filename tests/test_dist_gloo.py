"""CPU, world size 2, gloo: the host-side multi-GPU logic (row blocks, one-shot weight broadcast, output all-gather)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from shapeformer_b200 import dist as sdist


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        params = [torch.randn(5, 3, generator=g), torch.randn(7, generator=g), torch.randn(2, 2, 2, generator=g)]
        mine = [p.clone() if rank == 0 else torch.zeros_like(p) for p in params]
        sdist.broadcast_parameters(mine, src=0)
        ok = all(torch.equal(a, b) for a, b in zip(mine, params))
        lo, hi = sdist.row_block(12, rank, world, group=4)
        local = torch.arange(lo, hi, dtype=torch.int64)[:, None].repeat(1, 3)
        if hi - lo != 8 - 4 * rank:      # 3 groups of 4 over 2 ranks -> 8 + 4 rows: pad to equal shapes for all_gather
            ok = False
        pad = torch.full((8, 3), -1, dtype=torch.int64)
        pad[:hi - lo] = local
        rows = torch.cat([t[t[:, 0] >= 0] for t in sdist.gather_rows(pad)])
        ok = ok and torch.equal(rows[:, 0], torch.arange(12))
        # unequal blocks gathered directly (counts exchanged, padded internally, trimmed per rank)
        parts = sdist.gather_rows(local)
        ok = ok and [p.shape[0] for p in parts] == [8, 4] and torch.equal(torch.cat(parts)[:, 0], torch.arange(12))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_two_rank_plumbing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert res == [(0, True), (1, True)]


def test_row_block_partitions():
    for total, world, group in [(64, 8, 4), (512, 8, 4), (12, 5, 1), (64, 3, 4)]:
        spans = [sdist.row_block(total, r, world, group) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all((hi - lo) % group == 0 for lo, hi in spans)
