"""In-graph cost of a decoder layer per token at large-v3 width, by row count: decode-only runs (CUDA graph + PDL) of a model
with 1 encoder and N_DEC decoder layers, swept over row counts (ROWS=64,128) and TWB200_SKIP masks (MASKS=0,2 restricts them; TOKENS=252 for the bench's positions).
    python tools/decode_costs.py
Leaving a kernel out gives its marginal cost inside the step (profiles/r02_decode_step_costs.md); results of masked runs are garbage."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_DEC, TOKENS = 8, int(os.environ.get("TOKENS", "124"))

import torch
sys.path.insert(0, ROOT)
from taiwan_whisper_b200.hf_compat import build_hf_model  # noqa: E402
from taiwan_whisper_b200.configs import WhisperShape  # noqa: E402
from taiwan_whisper_b200.host import B200WhisperForConditionalGeneration  # noqa: E402

names = {0: "nothing", 1: "3 LayerNorms", 2: "self-attention", 4: "cross stream + combine", 8: "QKV", 16: "self out-proj", 32: "cross q",
         64: "cross out-proj", 128: "fc1", 256: "fc2", 504: "all six GEMMs"}
rows_list = [int(x) for x in os.environ.get("ROWS", "64,128").split(",")]
if os.environ.get("MASKS"):
    names = {int(k): names.get(int(k), f"mask {k}") for k in os.environ["MASKS"].split(",")}
sh = WhisperShape("lv3-8dec", 128, 1280, 5120, 20, 1, N_DEC, 51866)
with torch.device("cuda"):
    hf = build_hf_model(sh, seed=1)
res = {}
for rows in rows_list:
    m = B200WhisperForConditionalGeneration.from_hf(hf, dtype=torch.bfloat16, max_batch=rows)
    enc = (torch.randn((rows, 1500, 1280), device="cuda") * 0.5).bfloat16()
    prompt = m._init_tokens("zh", "transcribe", False)
    steps = len(prompt) + TOKENS - 1
    for mask in names:
        os.environ["TWB200_SKIP"] = str(mask)        # read by the library at every decode call
        m.decode(enc, prompt, len(prompt) + TOKENS, False)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            m.decode(enc, prompt, len(prompt) + TOKENS, False)
            torch.cuda.synchronize()
            best = min(best, m.last_stage_ms()["decode"])
        res[(rows, mask)] = 1000.0 * best / steps / N_DEC
    os.environ["TWB200_SKIP"] = "0"
    m.close()
    del enc
print("| left out | " + " | ".join(f"{r} rows: us / layer / token (marginal)" for r in rows_list) + " |")
print("|---|" + "---|" * len(rows_list))
for mask in names:
    print(f"| {names[mask]} | " + " | ".join(f"{res[(r, mask)]:.1f} ({res[(r, 0)] - res[(r, mask)]:.1f})" for r in rows_list) + " |", flush=True)
