"""Deterministic synthetic 30 s / 16 kHz clips (there is no network for datasets).

x = 0.05 N(0,1) + sum_k a_k sin(2 pi (f_k t + c_k t^2 / 2)) env(t), gated at 4 Hz with exact-zero
spans, clipped to [-1, 1) and stored as int16; both the CUDA path and the checker consume the
same dequantised samples (x_i16 / 32768 in f32).  Stands in for the 16 kHz flac segments the
reference reads (ref: prefiltering/validator_inference.py:119-140).
"""
from __future__ import annotations

import numpy as np

from .configs import N_SAMPLES, SAMPLING_RATE


def synth_clip(clip_id: int, seed: int = 0, n_samples: int = N_SAMPLES) -> np.ndarray:
    rng = np.random.default_rng(seed * 1_000_003 + clip_id)
    t = np.arange(n_samples, dtype=np.float64) / SAMPLING_RATE
    x = 0.05 * rng.standard_normal(n_samples)
    for _ in range(3):
        f = rng.uniform(80.0, 4000.0)
        c = rng.uniform(-50.0, 50.0)
        a = rng.uniform(0.02, 0.2)
        x += a * np.sin(2.0 * np.pi * (f * t + 0.5 * c * t * t))
    phase = rng.uniform(0.0, 1.0)
    env = 0.5 - 0.5 * np.cos(2.0 * np.pi * (4.0 * t + phase))
    x *= env
    # 20 % exact-zero spans (hard silences, as VAD-cut audio has)
    n_blocks = 50
    blk = n_samples // n_blocks
    for b in rng.choice(n_blocks, size=n_blocks // 5, replace=False):
        x[b * blk:(b + 1) * blk] = 0.0
    q = np.round(np.clip(x, -1.0, 1.0 - 1.0 / 32768.0) * 32768.0)
    return q.astype(np.int16)


def synth_batch(first_id: int, count: int, seed: int = 0, n_samples: int = N_SAMPLES) -> np.ndarray:
    return np.stack([synth_clip(first_id + i, seed, n_samples) for i in range(count)])


def dequantise(pcm_i16: np.ndarray) -> np.ndarray:
    return pcm_i16.astype(np.float32) / np.float32(32768.0)


def edge_case_clips() -> dict[str, np.ndarray]:
    """Parity edge cases: silence, short clip (zero-padded), full-scale square wave, impulse."""
    out = {}
    out["silence"] = np.zeros(N_SAMPLES, np.int16)
    short = np.zeros(N_SAMPLES, np.int16)
    short[:5 * SAMPLING_RATE] = synth_clip(7, 3, 5 * SAMPLING_RATE)
    out["short5s"] = short
    sq = np.where((np.arange(N_SAMPLES) // 40) % 2 == 0, 32767, -32768).astype(np.int16)
    out["square"] = sq
    imp = np.zeros(N_SAMPLES, np.int16)
    imp[123_457] = 32767
    out["impulse"] = imp
    return out
