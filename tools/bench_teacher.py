"""Teacher forward of the distillation step (tw_decoder_logits) at large-v3 size: B clips x T decoder positions."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taiwan_whisper_b200.hf_compat import build_hf_model  # noqa: E402
from taiwan_whisper_b200.configs import SHAPES  # noqa: E402
from taiwan_whisper_b200.host import B200WhisperForConditionalGeneration  # noqa: E402

B = int(os.environ.get("B", "64"))
T = int(os.environ.get("T", "128"))
sh = SHAPES[os.environ.get("MODEL", "large-v3")]
with torch.device("cuda"):
    hf = build_hf_model(sh, seed=1234)
m = B200WhisperForConditionalGeneration.from_hf(hf, dtype=torch.bfloat16, max_batch=B)
del hf
enc = (torch.randn((B, 1500, sh.d_model), device="cuda") * 0.5).bfloat16()
ids = torch.randint(0, 50000, (B, T), device="cuda")
for _ in range(2):
    lg = m.decoder_logits(enc, ids)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    lg = m.decoder_logits(enc, ids)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
d, f, L, V = sh.d_model, sh.ffn, sh.dec_layers, sh.vocab
M = B * T
flops = L * (2 * M * d * (3 * d + d + d + d) + 4 * M * d * f + 4 * B * T * T * d / 2 + 4 * B * T * 1500 * d + 4 * B * 1500 * d * d) + 2 * M * d * V
print(f"teacher forward {sh.name if hasattr(sh, 'name') else ''} B={B} T={T}: {ms:.2f} ms  ({flops / ms / 1e9:.0f} TFLOP/s incl. the cross-K/V projection), "
      f"{M / (ms / 1e3):.0f} label positions/s, logits finite: {bool(torch.isfinite(lg).all())}")
