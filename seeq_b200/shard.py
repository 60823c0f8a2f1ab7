"""Multi-GPU driver logic: newline-aligned sharding and the three tiny exchanges.

The matching path is embarrassingly parallel over lines (the reference resets
its state per line, libseeq.c:237-247; the only cross-line state is the line
counter, seeq.c:377).  One process per GPU scans its own newline-aligned byte
range; the ranks then exchange

  (i)   per-shard counted-line totals -> exclusive prefix = global line base,
  (ii)  match / record counts          -> sum,
  (iii) record arrays                  -> ordered concatenation (gather).

No data-path collective exists: the messages are a few integers (NCCL over
NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_ranges(buf: np.ndarray, world: int):
    """[(begin, end)] per rank: boundary k*n/world moved to just after the next '\\n'.

    Pure-numpy statement of sqbShardRange() (csrc/sqb_engine.cu); tests check
    that the two agree.
    """
    n = int(buf.size)
    cuts = [0]
    for r in range(1, world):
        p = n * r // world
        if p > 0 and p < n and buf[p - 1] != 0x0A:
            nl = np.flatnonzero(buf[p:] == 0x0A)
            p = p + int(nl[0]) + 1 if nl.size else n
        cuts.append(max(p, cuts[-1]))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def exchange(nlines: int, nmatched: int, nrecs: int, dist=None, device=None):
    """-> (line_base of this rank, total lines, total matched, total records).

    `dist` is torch.distributed (initialised) or None for a single process.
    """
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return 0, nlines, nmatched, nrecs
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.tensor([nlines, nmatched, nrecs], dtype=torch.int64, device=device)
    table = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(table, mine)
    table = torch.stack(table).cpu().numpy()
    base = int(table[:rank, 0].sum())
    return base, int(table[:, 0].sum()), int(table[:, 1].sum()), int(table[:, 2].sum())


def gather_records(recs: np.ndarray, line_base: int, dist=None):
    """Rebase shard-local record line indices and gather them in rank order on rank 0.

    recs: structured array with fields line/start/end/dist (binding.REC_DTYPE).
    Returns the concatenated (n,4) int64 array on rank 0, None elsewhere.
    """
    out = np.stack([recs["line"].astype(np.int64) + line_base, recs["start"].astype(np.int64),
                    recs["end"].astype(np.int64), recs["dist"].astype(np.int64)], axis=1) \
        if recs.size else np.zeros((0, 4), dtype=np.int64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return out
    parts = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(out, parts, dst=0)
    if dist.get_rank() != 0:
        return None
    return np.concatenate(parts, axis=0)
