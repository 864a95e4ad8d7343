"""Multi-GPU: shard (locus model, read batch) units over the ranks of one box.

Units are independent (the reference loops over them sequentially, genome_analyzer.py:280), so
there is NO collective on the data path: every rank decodes its own loci on its own GPU and
the per-locus results are gathered on the host at the end.  The only
cross-rank concern is load balance: loci are assigned longest-processing-time-first on the
estimated cost  sum_reads(len) * edges(locus).
"""
from __future__ import annotations


def locus_cost(n_edges, read_lengths):
    return float(n_edges) * float(sum(read_lengths))


def lpt_assign(costs, world_size):
    """rank of every unit: heaviest units first, each to the currently lightest rank."""
    load = [0.0] * world_size
    owner = [0] * len(costs)
    for i in sorted(range(len(costs)), key=lambda k: (-costs[k], k)):
        r = min(range(world_size), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += costs[i]
    return owner, load


def my_units(owner, rank):
    return [i for i, r in enumerate(owner) if r == rank]


def gather_results(local, owner, rank, world_size, group=None):
    """Host gather of per-unit results: ``local`` maps unit index -> result for the units this
    rank owns.  Returns the full list (unit order) on every rank.  Uses torch.distributed when
    more than one rank is running (gloo or nccl); a no-op otherwise."""
    if world_size == 1:
        return [local[i] for i in range(len(owner))]
    import torch.distributed as dist
    parts = [None] * world_size
    dist.all_gather_object(parts, local, group=group)
    out = [None] * len(owner)
    for r, part in enumerate(parts):
        for i, v in part.items():
            if owner[i] != r:
                raise RuntimeError("rank %d returned a unit it does not own" % r)
            out[i] = v
    if any(v is None for v in out):
        raise RuntimeError("some units were not decoded by any rank")
    return out


def gathered_row_of_every_read(owner, reads_per_unit, rows_max):
    """Where the reads of every unit sit in the table an ``all_gather`` of the per-rank result rows gives.

    Each rank decodes its units in ascending unit order and writes one result row per read, padded to
    ``rows_max`` rows per rank; the gathered table is rank-major.  Returns an int64 array ``idx`` with one
    entry per read in UNIT order: row ``idx[k]`` of the gathered table is read ``k``."""
    import numpy as np
    owner = np.asarray(owner)
    n = np.asarray(reads_per_unit, dtype=np.int64)
    first = np.zeros(len(n) + 1, dtype=np.int64)
    np.cumsum(n, out=first[1:])
    idx = np.empty(int(first[-1]), dtype=np.int64)
    for r in range(int(owner.max()) + 1 if len(owner) else 0):
        pos = r * int(rows_max)
        for u in np.nonzero(owner == r)[0]:
            idx[first[u]:first[u + 1]] = pos + np.arange(n[u])
            pos += int(n[u])
        if pos > (r + 1) * int(rows_max):
            raise ValueError("rank %d holds more rows than rows_max" % r)
    return idx
