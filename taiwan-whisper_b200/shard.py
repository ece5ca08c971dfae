"""Multi-GPU: shard the utterance manifest, no collective on the data path, one final gather.

Replaces the per-batch `accelerator.pad_across_processes` + `gather_for_metrics`
(ref: training/run_pseudo_labelling.py:919-922) and the batch-interleaved sharding of
`accelerator.prepare(eval_loader)` (ref: :906) with contiguous ranges per rank — the
per-rank-output + ordered merge pattern the reference itself uses in
ref: dataset/cool_dataset.py:216 / dataset/test_cool_dataset.sh:25.
Works with any torch.distributed backend (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of manifest rows owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_token_rows(tokens: torch.Tensor, lengths: torch.Tensor, pad_id: int, n_total: Optional[int] = None):
    """All ranks pass their [n_local, L] int32 token rows + [n_local] lengths (n_local may differ by one, and be
    zero).  Returns (tokens [N, L], lengths [N]) in manifest order on every rank (all_gather)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tokens, lengths
    world = dist.get_world_size()
    dev = tokens.device
    n_local = torch.tensor([tokens.shape[0], tokens.shape[1]], device=dev, dtype=torch.int64)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    n_max = max(int(s[0]) for s in sizes)
    L = max(int(s[1]) for s in sizes)
    buf = torch.full((n_max, L + 1), pad_id, dtype=torch.int32, device=dev)
    if tokens.shape[0]:
        buf[:tokens.shape[0], :tokens.shape[1]] = tokens.to(torch.int32)
        buf[:tokens.shape[0], L] = lengths.to(torch.int32)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    rows = torch.cat([o[:int(s[0])] for o, s in zip(out, sizes)])
    if n_total is not None and rows.shape[0] != n_total:
        raise RuntimeError(f"gathered {rows.shape[0]} rows, manifest has {n_total}")
    return rows[:, :L], rows[:, L]


def run_manifest(n_clips: int, batch: int, load_batch: Callable[[int, int], torch.Tensor],
                 transcribe: Callable[[torch.Tensor], Tuple[torch.Tensor, torch.Tensor]], pad_id: int,
                 rank: int = 0, world: int = 1, gather: bool = True):
    """Pseudo-labels manifest rows [0, n_clips): each rank walks its contiguous shard in batches
    (`load_batch(lo, hi)` -> host int16 PCM [hi-lo, 480000]; `transcribe(pcm)` -> (tokens, lengths)),
    then one gather.  Returns (tokens, lengths) for the whole manifest (manifest order)."""
    lo, hi = shard_bounds(n_clips, rank, world)
    toks, lens = [], []
    for s in range(lo, hi, batch):
        e = min(hi, s + batch)
        t, l = transcribe(load_batch(s, e))
        toks.append(t.clone())
        lens.append(l.clone())
    if toks:
        tokens, lengths = torch.cat(toks), torch.cat(lens)
    else:
        tokens, lengths = torch.zeros((0, 1), dtype=torch.int32), torch.zeros((0,), dtype=torch.int32)
    if gather and world > 1:
        if dist.get_backend() == "nccl":
            tokens, lengths = tokens.cuda(), lengths.cuda()
        return gather_token_rows(tokens, lengths, pad_id, n_clips)
    return tokens, lengths
