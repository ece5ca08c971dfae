"""Encoder at large-v3 width (d=1280, ffn=5120, 20 heads, 128 mel) with 2 layers, batch 16, for an ncu launch list."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taiwan_whisper_b200.hf_compat import build_hf_model  # noqa: E402
from taiwan_whisper_b200.configs import WhisperShape  # noqa: E402
from taiwan_whisper_b200.host import B200WhisperForConditionalGeneration, log_mel  # noqa: E402
from taiwan_whisper_b200.synth import synth_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sh = WhisperShape("lv3-2enc", 128, 1280, 5120, 20, 2, 1, 51866)
with torch.device("cuda"):
    hf = build_hf_model(sh, seed=1)
m = B200WhisperForConditionalGeneration.from_hf(hf, dtype=torch.bfloat16, max_batch=B)
pcm = torch.from_numpy(synth_batch(0, 4)).repeat(B // 4, 1).cuda()
for _ in range(2):
    mel = log_mel(pcm, None, 128)
    enc = m.encode(mel)
torch.cuda.synchronize()
print("ok", enc.shape)
