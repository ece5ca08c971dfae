"""Chunked long-form front end: cut a long recording into 30 s windows with a stride on each side
and compute their log-mel features on the GPU *without materialising the windows* — window b is row b of a
strided view of the recording (row pitch = chunk - stride_left - stride_right samples), which is exactly the
`pcm_stride` / `n_valid` interface of tw_logmel.

Window placement and the (chunk_len, stride_left, stride_right) bookkeeping restate
ref: training/flax/distil_whisper/pipeline.py:224-254 (chunk_iter_with_batch) and :325-335 (stride = chunk/6).
`stitch_windows` merges the per-window token streams of overlapping windows the way `tokenizer._decode_asr` does for
the no-timestamp case (ref :353-375 -> HF tokenization_whisper.py `_find_longest_common_sequence`); ids -> text stays
with the tokenizer (third-party vocabulary data).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from . import lib as _lib
from .configs import N_FRAMES, N_SAMPLES, SAMPLING_RATE


def chunk_plan(n_samples: int, chunk_len: int = N_SAMPLES, stride_left: Optional[int] = None,
               stride_right: Optional[int] = None) -> Tuple[np.ndarray, List[Tuple[int, int, int]]]:
    """Window start indices and per-window (chunk_len_actual, stride_left, stride_right)."""
    if stride_left is None:
        stride_left = round(chunk_len / 6)
    if stride_right is None:
        stride_right = stride_left
    if chunk_len < stride_left + stride_right:
        raise ValueError("Chunk length must be superior to stride length")
    step = chunk_len - stride_left - stride_right
    starts = np.arange(0, n_samples, step)
    ends = starts + chunk_len
    sl = np.where(starts == 0, 0, stride_left)
    is_last = np.where(stride_right > 0, ends > n_samples, ends >= n_samples)
    sr = np.where(is_last, 0, stride_right)
    lens = np.minimum(ends, n_samples) - starts
    return starts, [(int(l), int(a), int(b)) for l, a, b in zip(lens, sl, sr)]


def chunked_log_mel(pcm: torch.Tensor, n_mel: int, chunk_len: int = N_SAMPLES, stride_left: Optional[int] = None,
                    stride_right: Optional[int] = None):
    """pcm: 1-D int16 / float32 CUDA tensor (one recording, 16 kHz).  Returns (features [n_windows, n_mel, 3000] f32,
    strides list) — the `{"stride": ..., "input_features": ...}` items the reference's chunker yields."""
    if pcm.dim() != 1 or not pcm.is_cuda or pcm.dtype not in (torch.int16, torch.float32):
        raise ValueError("chunked_log_mel expects a 1-D int16/float32 CUDA tensor")
    if chunk_len != N_SAMPLES:
        raise NotImplementedError("the log-mel kernel is specialised for 30 s windows")
    starts, strides = chunk_plan(pcm.shape[0], chunk_len, stride_left, stride_right)
    n = len(starts)
    step = int(starts[1] - starts[0]) if n > 1 else chunk_len
    dev = pcm.device.index if pcm.device.index is not None else torch.cuda.current_device()
    ctx = _lib.Context.get(dev)
    pcm = pcm.contiguous()
    out = torch.empty((n, n_mel, N_FRAMES), dtype=torch.float32, device=pcm.device)
    n_valid = torch.tensor([s[0] for s in strides], dtype=torch.int32, device=pcm.device)
    dt = _lib.TW_I16 if pcm.dtype == torch.int16 else _lib.TW_F32
    with torch.cuda.device(dev):
        # row b of the "batch" starts at sample b*step of the recording: zero-copy windowing
        ctx.check(ctx.lib.tw_logmel(ctx.handle, pcm.data_ptr(), dt, step, n_valid.data_ptr(), n, n_mel, out.data_ptr(),
                                    torch.cuda.current_stream(pcm.device).cuda_stream))
    return out, strides


def stitch_windows(sequences: List[List[int]]) -> List[int]:
    """Merge the token streams of consecutive overlapping windows into one stream (no-timestamp mode).

    Restates HF `_find_longest_common_sequence` (tokenization_whisper.py; called by `_decode_asr`, which
    ref: training/flax/distil_whisper/pipeline.py:353-375 calls): for each next window slide it over the tail of what
    has been kept so far; an alignment of overlap length i scores (#equal tokens) / i + i / 10000 (the epsilon prefers
    longer overlaps) and needs at least two equal tokens; at the best alignment keep the left stream up to the middle of
    the overlap and continue with the right stream from the middle of its side of the overlap."""
    if not sequences:
        return []
    left = list(sequences[0])
    total: List[int] = []
    for right in sequences[1:]:
        right = list(right)
        n_l, n_r = len(left), len(right)
        best, best_idx = 0.0, (n_l, n_l, 0, 0)
        la, ra = np.asarray(left, dtype=np.int64), np.asarray(right, dtype=np.int64)
        for i in range(1, n_l + n_r):
            l0, l1 = max(0, n_l - i), min(n_l, n_l + n_r - i)
            r0, r1 = max(0, i - n_l), min(n_r, i)
            matches = int(np.sum(la[l0:l1] == ra[r0:r1]))
            score = matches / i + i / 10000.0
            if matches > 1 and score > best:
                best, best_idx = score, (l0, l1, r0, r1)
        l0, l1, r0, r1 = best_idx
        total.extend(left[:(l0 + l1) // 2])
        left = right[(r0 + r1) // 2:]
    total.extend(left)
    return total
