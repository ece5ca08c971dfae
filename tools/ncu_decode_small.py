"""Short decode run at large-v3 width (d=1280, ffn=5120, 20 heads, vocab 51866, batch 64 — or argv 2 rows, e.g. 192 = three
merged batches) with only 1 encoder and 2 decoder layers, for an ncu launch list of the decode-step kernels (TWB200_GRAPH=0 so every launch is visible)."""
import os
import sys

os.environ.setdefault("TWB200_GRAPH", "0")
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taiwan_whisper_b200.hf_compat import build_hf_model  # noqa: E402
from taiwan_whisper_b200.configs import WhisperShape  # noqa: E402
from taiwan_whisper_b200.host import B200WhisperForConditionalGeneration  # noqa: E402

sh = WhisperShape("lv3-2dec", 128, 1280, 5120, 20, 1, 2, 51866)
with torch.device("cuda"):
    hf = build_hf_model(sh, seed=1)
ROWS = int(sys.argv[2]) if len(sys.argv) > 2 else 64
m = B200WhisperForConditionalGeneration.from_hf(hf, dtype=torch.bfloat16, max_batch=ROWS)
enc = (torch.randn((ROWS, 1500, 1280), device="cuda") * 0.5).bfloat16()
prompt = m._init_tokens("zh", "transcribe", False)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
toks, lens = m.decode(enc, prompt, len(prompt) + n, False)
torch.cuda.synchronize()
print("ok", toks.shape)
