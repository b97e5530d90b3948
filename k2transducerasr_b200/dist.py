"""Multi-GPU plumbing: one process per GPU, utterances / streams sharded in contiguous blocks, no
data-path collective (streams are independent: SURVEY.md section 8e). The only exchange is a gather of the
final (tokens, timestamps, score) per stream for reporting, done once per batch with torch.distributed
(NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of `n_units` streams owned by `rank` (first ranks take the remainder)."""
    base, rem = divmod(n_units, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_results(tokens: Sequence[Sequence[int]], ts: Sequence[Sequence[int]], scores: Optional[Sequence[float]],
                 max_len: int) -> torch.Tensor:
    """[n, 2 + 2*max_len] int32: count, score bits, tokens..., timestamps..."""
    n = len(tokens)
    out = np.zeros((n, 2 + 2 * max_len), np.int32)
    for i in range(n):
        k = len(tokens[i])
        if k > max_len:
            raise ValueError("max_len too small for the gathered results")
        out[i, 0] = k
        out[i, 1] = np.float32(scores[i] if scores is not None else 0.0).view(np.int32)
        out[i, 2:2 + k] = tokens[i]
        out[i, 2 + max_len:2 + max_len + k] = ts[i]
    return torch.from_numpy(out)


def unpack_results(packed: torch.Tensor, max_len: int):
    a = packed.cpu().numpy()
    toks, tss, scores = [], [], []
    for row in a:
        k = int(row[0])
        scores.append(float(row[1:2].view(np.float32)[0]))
        toks.append(row[2:2 + k].tolist())
        tss.append(row[2 + max_len:2 + max_len + k].tolist())
    return toks, tss, scores


def gather_results(tokens, ts, scores, n_total: int, max_len: int, device: Optional[torch.device] = None,
                   group=None):
    """All-gather every rank's shard of results; returns (tokens, ts, scores) for all n_total streams in
    global stream order on every rank. Shards follow shard_range()."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    biggest = max(b - a for a, b in sizes)
    mine = pack_results(tokens, ts, scores, max_len)
    pad = torch.zeros((biggest, mine.shape[1]), dtype=torch.int32)
    pad[: mine.shape[0]] = mine
    if device is not None:
        pad = pad.to(device)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    all_t, all_s, all_sc = [], [], []
    for r, (a, b) in enumerate(sizes):
        t, s, sc = unpack_results(bufs[r][: b - a], max_len)
        all_t += t; all_s += s; all_sc += sc
    return all_t, all_s, all_sc
