"""Probe: pipelined batch loop (encoder of batch i+1 on a small SM partition while batch i decodes) against the sequential
transcribe_pcm loop: large-v3, 64 clips per batch, pinned host PCM in, host ids out.  usage: pipeline_probe.py [n_enc_sms ...]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taiwan_whisper_b200.configs import SHAPES  # noqa: E402
from taiwan_whisper_b200.hf_compat import build_hf_model  # noqa: E402
from taiwan_whisper_b200.host import B200WhisperForConditionalGeneration  # noqa: E402
from taiwan_whisper_b200.synth import synth_batch  # noqa: E402

sh = SHAPES["large-v3"]
B, NB, ML = 64, 5, 256
n_enc = int(sys.argv[1]) if len(sys.argv) > 1 else 24
with torch.device("cuda"):
    hf = build_hf_model(sh, seed=1234)
m = B200WhisperForConditionalGeneration.from_hf(hf, dtype=torch.bfloat16, max_batch=B, output_layout="5.x")
del hf
pool = torch.from_numpy(synth_batch(0, B)).pin_memory()
batches = [pool for _ in range(NB)]
out_tok = torch.empty((B, ML - 4), dtype=torch.int32).pin_memory()
out_len = torch.empty((B,), dtype=torch.int32).pin_memory()
m.transcribe_pcm(pool, ML, out_tokens=out_tok, out_lengths=out_len)
torch.cuda.synchronize()
t0 = time.perf_counter()
for b in batches:
    m.transcribe_pcm(b, ML, out_tokens=out_tok, out_lengths=out_len)
torch.cuda.synchronize()
seq = (time.perf_counter() - t0) / NB
ref = out_tok.clone()
print(f"sequential: {seq*1000:.1f} ms per batch = {B*30/seq:.1f} audio-s/s  stages {m.last_stage_ms()}")
print("pipeline SMs (encoder, decode):", m.enable_pipeline(n_enc))
stamps = []
agree = 0.0
for toks, lens in m.transcribe_batches(batches + batches[:3], ML):
    torch.cuda.synchronize()
    stamps.append(time.perf_counter())
    agree = (toks == ref).float().mean().item()
iv = [b - a for a, b in zip(stamps[1:], stamps[2:])]          # steady state: intervals between decode completions after the first two
pip = sorted(iv)[len(iv) // 2]
print(f"pipelined ({n_enc} SMs): intervals {[round(x*1000) for x in iv]} ms, median {pip*1000:.1f} ms per batch = {B*30/pip:.1f} audio-s/s  "
      f"stages {m.last_stage_ms()}  free-running token agreement with sequential {agree:.3f}")
