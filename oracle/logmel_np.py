"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement of the Whisper log-mel front end.  Follows
  transformers/models/whisper/feature_extraction_whisper.py:135-164 (_torch_extract_fbank_features,
  the path HF takes whenever torch is importable) and :105-133 (numpy twin),
  transformers/audio_utils.py:263-333 (slaney mel scale), :356-375 (triangular filters),
  :453-545 (mel_filter_bank), and the in-reference twin
  ref: training/flax/distil_whisper/pipeline.py:40-58.
The FFT is done in float64 (as HF's numpy path does) and the result cast to float32; HF's torch
path does the STFT in float32 — the two agree to ~1e-5 (HF's own stated tolerance), the
contract for the CUDA kernel is 1e-4 abs.
"""
from __future__ import annotations

import numpy as np

N_FFT = 400
HOP = 160
N_SAMPLES = 480_000
N_FRAMES = 3000


def hertz_to_mel_slaney(freq):
    freq = np.asarray(freq, dtype=np.float64)
    mels = 3.0 * freq / 200.0
    logstep = 27.0 / np.log(6.4)
    log_region = freq >= 1000.0
    with np.errstate(divide="ignore"):
        mels = np.where(log_region, 15.0 + np.log(np.maximum(freq, 1e-300) / 1000.0) * logstep, mels)
    return mels


def mel_to_hertz_slaney(mels):
    mels = np.asarray(mels, dtype=np.float64)
    freq = 200.0 * mels / 3.0
    logstep = np.log(6.4) / 27.0
    log_region = mels >= 15.0
    return np.where(log_region, 1000.0 * np.exp(logstep * (mels - 15.0)), freq)


def mel_filter_bank(n_mel: int, n_freq: int = 201, sr: int = 16000, fmax: float = 8000.0) -> np.ndarray:
    """[n_freq, n_mel] float64 — slaney scale, slaney (area) normalisation."""
    mel_pts = np.linspace(hertz_to_mel_slaney(0.0), hertz_to_mel_slaney(fmax), n_mel + 2)
    filter_freqs = mel_to_hertz_slaney(mel_pts)
    fft_freqs = np.linspace(0, sr // 2, n_freq)
    diff = np.diff(filter_freqs)
    slopes = filter_freqs[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / diff[:-1]
    up = slopes[:, 2:] / diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    fb *= (2.0 / (filter_freqs[2:n_mel + 2] - filter_freqs[:n_mel]))[None, :]
    return fb


def hann_periodic(n: int = N_FFT) -> np.ndarray:
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def pad_or_trim(x: np.ndarray, n: int = N_SAMPLES) -> np.ndarray:
    """ref: prefiltering/validator_inference.py:131-137 and HF __call__ (:296-303)."""
    x = np.asarray(x, dtype=np.float32)
    if x.shape[-1] >= n:
        return x[..., :n]
    pad = [(0, 0)] * (x.ndim - 1) + [(0, n - x.shape[-1])]
    return np.pad(x, pad)


def log_mel(x: np.ndarray, n_mel: int) -> np.ndarray:
    """x: [480000] or [B,480000] float32 in [-1,1] -> [B?, n_mel, 3000] float32."""
    x = np.asarray(x, dtype=np.float32)
    if x.ndim == 2:
        return np.stack([log_mel(r, n_mel) for r in x])
    assert x.shape[0] == N_SAMPLES, "pad_or_trim first"
    xp = np.pad(x.astype(np.float64), (N_FFT // 2, N_FFT // 2), mode="reflect")
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(N_FRAMES)[:, None]     # frame 3000 is dropped
    frames = xp[idx] * hann_periodic()[None, :]
    spec = np.fft.rfft(frames, axis=1)                                       # [3000, 201]
    power = spec.real ** 2 + spec.imag ** 2
    fb = mel_filter_bank(n_mel).astype(np.float32).astype(np.float64)        # HF stores the bank as f32
    mel = fb.T @ power.T                                                     # [n_mel, 3000]
    logs = np.log10(np.maximum(mel, 1e-10))
    logs = np.maximum(logs, logs.max() - 8.0)                                # per-clip max
    return ((logs + 4.0) / 4.0).astype(np.float32)
