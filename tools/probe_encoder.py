"""Probe: encoder stage time (large-v3, 64 clips) — used for A/B runs of two builds on one box."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taiwan_whisper_b200.hf_compat import build_hf_model  # noqa: E402
from taiwan_whisper_b200.configs import SHAPES  # noqa: E402
from taiwan_whisper_b200.host import B200WhisperForConditionalGeneration  # noqa: E402

sh = SHAPES["large-v3"]
B = 64
with torch.device("cuda"):
    hf = build_hf_model(sh, seed=1234)
m = B200WhisperForConditionalGeneration.from_hf(hf, dtype=torch.bfloat16, max_batch=B)
del hf
mel = torch.randn((B, 128, 3000), device="cuda") * 0.3
for _ in range(2):
    m.encode(mel)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(4):
    e0.record()
    m.encode(mel)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"encoder large-v3 B=64 : " + " ".join(f"{t:.1f}" for t in ts) + " ms")
