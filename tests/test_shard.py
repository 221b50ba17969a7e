"""Multi-GPU host logic on CPU: batches cut from a read stream, dealt to ranks, merged back in input order.
The world_size-2 test runs two real processes over torch.distributed (gloo) with the CPU oracle standing in for the
per-GPU engine; it checks that the union of the ranks' outputs equals the single-process result, in order."""
import os
import sys

import numpy as np
import pytest

from hifimeth_b200 import synth
import shard_model as shard

from conftest import ROOT


def test_cut_batches_bounds_and_coverage():
    rng = np.random.default_rng(0)
    lens = rng.integers(300, 25000, size=257)
    batches = shard.cut_batches(lens, max_reads=16, max_bases=120_000)
    assert [b.seq for b in batches] == list(range(len(batches)))
    assert sum(b.count for b in batches) == len(lens)
    pos = 0
    for b in batches:
        assert b.first == pos and 1 <= b.count <= 16
        assert b.bases == int(lens[pos:pos + b.count].sum())
        assert b.bases <= 120_000 or b.count == 1
        pos += b.count
    assert shard.cut_batches([], 4, 100) == []
    assert [b.count for b in shard.cut_batches([500_000, 10], 4, 1000)] == [1, 1]  # oversized read: its own batch


def test_assign_and_merge():
    batches = shard.cut_batches([1000] * 23, max_reads=3, max_bases=10_000)
    parts = [[(b.seq, b.first) for b in shard.assign(batches, 4, r)] for r in range(4)]
    assert sorted(s for p in parts for s, _ in p) == list(range(len(batches)))
    assert shard.merge_in_order(parts) == [b.first for b in batches]
    with pytest.raises(ValueError):
        shard.merge_in_order([parts[0], parts[0]])
    with pytest.raises(ValueError):
        shard.merge_in_order(parts[:3])


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist

    from oracle import hmoracle

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch, _ = synth.make_reads(11, (1000, 1600), seed=77, short_every=5)
    lens = np.diff(batch.base_off.astype(np.int64))
    batches = shard.cut_batches(lens, max_reads=2, max_bases=4000)
    sites = hmoracle.oracle().batch_sites(batch, 7)  # stand-in for the per-GPU engine (site lists are bit-exact on both)
    mine = [(b.seq, [sites[r]["qoff"].tolist() for r in range(b.first, b.first + b.count)]) for b in shard.assign(batches, world, rank)]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)  # results travel to the ordered writer; the data path itself has no collective
    if rank == 0:
        merged = [q_ for payload in shard.merge_in_order(gathered) for q_ in payload]
        q.put(merged == [s["qoff"].tolist() for s in sites])
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_reproduce_single_process_order():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
