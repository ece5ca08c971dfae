"""Smoke + timing of the other BASELINE.json configs on one GPU (random-init, synthetic audio):
  configs[3] distil student (large-v3 encoder, 2 decoder layers), batch 128
  configs[4] whisper-medium validator, batch 32, return_timestamps=True, max_length 448
  configs[0] whisper-tiny batch 1"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taiwan_whisper_b200.hf_compat import build_hf_model  # noqa: E402
from taiwan_whisper_b200.configs import SHAPES  # noqa: E402
from taiwan_whisper_b200.host import B200WhisperForConditionalGeneration  # noqa: E402
from taiwan_whisper_b200.synth import synth_batch  # noqa: E402

out = {}
for name, B, max_length, ts in (("distil-large-v3", 128, 256, False), ("medium", 32, 448, True), ("tiny", 1, 64, False)):
    sh = SHAPES[name]
    with torch.device("cuda"):
        hf = build_hf_model(sh, seed=1234)
    m = B200WhisperForConditionalGeneration.from_hf(hf, dtype=torch.bfloat16, max_batch=B)
    del hf
    pcm = torch.from_numpy(synth_batch(0, min(B, 16))).repeat((B + 15) // 16, 1)[:B].contiguous().pin_memory()
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        toks, lens = m.transcribe_pcm(pcm, max_length, return_timestamps=ts)
        dt = time.perf_counter() - t0
    out[name] = {"batch": B, "max_length": max_length, "timestamps": ts, "sec_per_batch": dt, "rtfx": B * 30.0 / dt,
                 "stage_ms": m.last_stage_ms(), "mean_len": float(lens.float().mean()), "first_tokens": toks[0, :6].tolist()}
    print(name, json.dumps(out[name]))
    m.close()
    torch.cuda.empty_cache()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
