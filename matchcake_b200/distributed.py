"""Sharding of the embarrassingly parallel axes over the GPUs of one node (one process per GPU).

The path has exactly one exchange step: an all-gather of equally sized row blocks (NCCL over NVLink on GPUs;
``gloo`` in the CPU tests).  No reductions, no point-to-point.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np
import torch


def world() -> Tuple[int, int]:
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def even_ranges(n: int, parts: int) -> List[Tuple[int, int]]:
    """Contiguous split of ``range(n)`` into ``parts`` nearly equal pieces (batch axis of expvals / probs)."""
    base, rem = divmod(n, parts)
    out, s = [], 0
    for p in range(parts):
        e = s + base + (1 if p < rem else 0)
        out.append((s, e))
        s = e
    return out


def gram_row_cost(n0: int, n1: int, full: bool = False) -> np.ndarray:
    """Number of evaluated cells in each Gram row under the reference's triangle rule (gram_matrix.py:191-195)."""
    i = np.arange(n0)
    if full:
        return np.full(n0, n1, dtype=np.int64)
    if n0 < n1:
        return np.maximum(n1 - 1 - i, 0).astype(np.int64)
    return np.minimum(i, n1).astype(np.int64)


def balanced_row_ranges(n0: int, n1: int, parts: int, full: bool = False) -> List[Tuple[int, int]]:
    """Contiguous row blocks with (nearly) equal numbers of evaluated cells."""
    cost = gram_row_cost(n0, n1, full)
    csum = np.concatenate([[0], np.cumsum(cost)])
    total = csum[-1]
    bounds = [0]
    for p in range(1, parts):
        bounds.append(int(np.searchsorted(csum, total * p / parts, side="left")))
    bounds.append(n0)
    bounds = np.maximum.accumulate(np.clip(bounds, 0, n0))
    return [(int(bounds[p]), int(bounds[p + 1])) for p in range(parts)]


def folded_row_blocks(n: int, parts: int) -> List[Tuple[Tuple[int, int], Tuple[int, int]]]:
    """Square triangular Gram: rank r gets a block of rows from the top AND the mirrored block from the bottom
    (rows taken in the order 0, n-1, 1, n-2, ...), so that every rank holds the same number of ROWS (equal-size
    all-gather without padding) and the same number of evaluated CELLS (row i costs i, row n-1-i costs n-1-i).
    Returns, per rank, ((top_begin, top_end), (bottom_begin, bottom_end))."""
    c = -(-n // parts)
    out = []
    for r in range(parts):
        k0, k1 = min(r * c, n), min((r + 1) * c, n)
        top = ((k0 + 1) // 2, (k1 + 1) // 2)
        bottom = (n - k1 // 2, n - k0 // 2)
        out.append((top, bottom))
    return out


def folded_rows_in_place(compute_rows: Callable[[int, int, torch.Tensor], None], n: int, n_cols: int,
                         dtype: torch.dtype, device, timeline: Optional[list] = None) -> torch.Tensor:
    """Triangular square Gram over the ranks WITHOUT staging copies: every rank computes its top and bottom row block
    directly into the rows of the final (n, n_cols) matrix (``compute_rows(begin, end, out_rows)``), then two
    all-gathers assemble it -- the top halves in place (ascending rows with ascending rank: slot ``rank`` of the
    output IS the input), the bottom halves (descending rows with ascending rank) with the row-block views of the
    final matrix as the output list.  Needs n divisible by 2 * world_size (equal blocks); the caller falls back to
    :func:`sharded_folded_rows` otherwise."""
    import torch.distributed as dist

    rank, ws = world()
    h = n // (2 * ws)
    assert n == 2 * ws * h and ws > 1
    (t0, t1), (b0, b1) = folded_row_blocks(n, ws)[rank]
    assert (t0, t1) == (rank * h, (rank + 1) * h) and (b0, b1) == (n - (rank + 1) * h, n - rank * h)
    k = torch.empty((n, n_cols), dtype=dtype, device=device)
    mark = (lambda name: timeline.append((name, _event()))) if timeline is not None else (lambda name: None)
    mark("begin")
    compute_rows(t0, t1, k[t0:t1])
    mark("top_rows")
    compute_rows(b0, b1, k[b0:b1])
    mark("bottom_rows")
    # Both exchanges AFTER both kernels.  Measured at N = 8 (C3): issuing the top block's all-gather asynchronously under
    # the bottom kernel made the step 47.0 ms instead of 37.1 -- the collective's kernel holds SMs while it waits for the
    # slowest peer, and the persistent one-block-per-SM Gram grid then runs its share of those SMs' work late.
    dist.all_gather_into_tensor(k[: n // 2], k[t0:t1])                      # in place: slot `rank` is the input
    dist.all_gather([k[n - (r + 1) * h: n - r * h] for r in range(ws)], k[b0:b1])
    mark("all_gather")
    return k


def _event():
    e = torch.cuda.Event(enable_timing=True) if torch.cuda.is_available() else None
    if e is not None:
        e.record()
    return e


def gather_row_ranges(k: torch.Tensor, ranges: List[Tuple[int, int]]) -> torch.Tensor:
    """``k`` is the full (n_rows, ...) matrix in which this rank has filled rows ranges[rank]; fill in everybody
    else's rows.  NCCL: one all-gather whose outputs are the row-block views of ``k`` (unequal block sizes are fine);
    gloo (CPU tests) needs equal sizes, so the blocks are padded there."""
    import torch.distributed as dist

    rank, ws = world()
    if ws == 1:
        return k
    s, e = ranges[rank]
    if dist.get_backend() == "nccl":
        dist.all_gather([k[a:b] for a, b in ranges], k[s:e])
        return k
    full = all_gather_rows(k[s:e], ranges, k.shape[0])
    k.copy_(full)
    return k


def sharded_folded_rows(compute_rows: Callable[[int, int], torch.Tensor], n: int) -> torch.Tensor:
    """Triangular square Gram over the ranks: each rank computes its two row blocks (``compute_rows(begin, end)`` ->
    (end - begin, n) tensor), ONE equal-size all-gather assembles the (n, n) result."""
    import torch.distributed as dist

    rank, ws = world()
    if ws == 1:
        return compute_rows(0, n)
    blocks = folded_row_blocks(n, ws)
    c = -(-n // ws)
    (t0, t1), (b0, b1) = blocks[rank]
    parts = [compute_rows(t0, t1)] if t1 > t0 else []
    if b1 > b0:
        parts.append(compute_rows(b0, b1))
    if not parts:  # more ranks than rows: this rank only takes part in the collective
        parts = [compute_rows(0, 0)]
    ref = parts[0]
    local = torch.zeros((c,) + tuple(ref.shape[1:]), dtype=ref.dtype, device=ref.device)
    o = 0
    for part in parts:
        local[o:o + part.shape[0]] = part
        o += part.shape[0]
    gathered = torch.empty((ws * c,) + tuple(ref.shape[1:]), dtype=ref.dtype, device=ref.device)
    dist.all_gather_into_tensor(gathered, local)
    out = torch.empty((n,) + tuple(ref.shape[1:]), dtype=ref.dtype, device=ref.device)
    for r, ((s0, e0), (s1, e1)) in enumerate(blocks):
        out[s0:e0] = gathered[r * c: r * c + (e0 - s0)]
        out[s1:e1] = gathered[r * c + (e0 - s0): r * c + (e0 - s0) + (e1 - s1)]
    return out


def all_gather_rows(local: torch.Tensor, ranges: List[Tuple[int, int]], n_rows: int) -> torch.Tensor:
    """Assemble a (n_rows, ...) tensor from per-rank row blocks ``local`` = rows ranges[rank] with ONE all-gather
    (blocks are padded to the largest block so that the collective is a plain equal-size all-gather)."""
    import torch.distributed as dist

    rank, ws = world()
    if ws == 1:
        return local
    max_rows = max(e - s for s, e in ranges)
    pad = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    gathered = torch.empty((ws * max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, pad.contiguous())
    out = torch.empty((n_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r, (s, e) in enumerate(ranges):
        out[s:e] = gathered[r * max_rows: r * max_rows + (e - s)]
    return out


def sharded_rows(compute_rows: Callable[[int, int], torch.Tensor], ranges: List[Tuple[int, int]], n_rows: int) -> torch.Tensor:
    """Run ``compute_rows(begin, end)`` for this rank's block and all-gather the result."""
    rank, ws = world()
    s, e = ranges[rank]
    return all_gather_rows(compute_rows(s, e), ranges, n_rows)
