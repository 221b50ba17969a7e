"""Host-side sharding of a read stream over the GPUs of one box.

Reads are independent units (no cross-read state anywhere in src/app/hifimeth/mod_main.cpp:180-212), so the path shards with
no exchange step: the input is cut into contiguous read batches (bounded by reads and bases, like the reference's outer
batch of `-b` reads, src/corelib/sam_batch.hpp:38-54), every batch carries a sequence number, batches are dealt to one
worker per GPU, and a single ordered writer re-emits records in input order (the reference sorts by read id,
mod_main.cpp:353-362).  No collective is involved; weights are replicated.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Sequence


@dataclass(frozen=True)
class ReadRange:
    seq: int     # batch sequence number = position in the input stream
    first: int   # first read (index in the input stream)
    count: int   # reads in the batch
    bases: int   # sum of l_qseq


def cut_batches(read_lengths: Sequence[int], max_reads: int, max_bases: int) -> List[ReadRange]:
    """Contiguous batches of <= max_reads reads and <= max_bases bases, in input order.  A read longer than max_bases gets
    a batch of its own (the engine rejects it with HM_ERR_ARG, as one oversized record would be in the reference's buffers)."""
    if max_reads < 1 or max_bases < 1:
        raise ValueError("max_reads and max_bases must be positive")
    out, first, n, b = [], 0, 0, 0
    for i, l in enumerate(read_lengths):
        l = int(l)
        if n and (n == max_reads or b + l > max_bases):
            out.append(ReadRange(len(out), first, n, b))
            first, n, b = i, 0, 0
        n += 1
        b += l
    if n:
        out.append(ReadRange(len(out), first, n, b))
    return out


def assign(batches: Sequence[ReadRange], world: int, rank: int) -> List[ReadRange]:
    """Static deal of batches to ranks: batch k goes to rank k % world (what a work queue yields when batches cost the same)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return [b for b in batches if b.seq % world == rank]


def merge_in_order(parts: Iterable[Iterable[tuple]]) -> list:
    """parts: per rank, an iterable of (seq, payload).  Returns payloads in sequence order; raises if a batch is missing or
    duplicated (the ordered writer must emit every input record exactly once)."""
    got = {}
    for part in parts:
        for seq, payload in part:
            if seq in got:
                raise ValueError(f"batch {seq} produced twice")
            got[seq] = payload
    if sorted(got) != list(range(len(got))):
        raise ValueError("missing batch in the merged stream")
    return [got[k] for k in range(len(got))]
