"""log-mel kernel alone at the bench shape (64 clips of int16 PCM, 128 mel) for timing / ncu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taiwan_whisper_b200.host import log_mel  # noqa: E402
from taiwan_whisper_b200.synth import synth_batch  # noqa: E402

B = 64
pcm = torch.from_numpy(synth_batch(0, 8)).repeat(B // 8, 1).cuda()
for _ in range(3):
    out = log_mel(pcm, None, 128)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = log_mel(pcm, None, 128)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
byt = B * (480000 * 2 + 128 * 3000 * 4)
print(f"log-mel B={B}: {ms*1000:.1f} us per batch, {ms*1000/B:.2f} us/clip, {byt/ms/1e6:.0f} GB/s algorithmic")
