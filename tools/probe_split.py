"""Probe: decode time of one 64-clip large-v3 batch with the decoder layers split into 1..4 sub-batches
(TWB200_SPLIT is read when the model is loaded, so one model instance per setting)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taiwan_whisper_b200.hf_compat import build_hf_model  # noqa: E402
from taiwan_whisper_b200.configs import SHAPES  # noqa: E402
from taiwan_whisper_b200.host import B200WhisperForConditionalGeneration  # noqa: E402

settings = sys.argv[1:] or ["1", "2", "3", "4"]
B = int(os.environ.get("PROBE_B", "64"))
sh = SHAPES["large-v3"]
with torch.device("cuda"):
    hf = build_hf_model(sh, seed=1234)
ML = 256
enc = (torch.randn((B, 1500, 1280), device="cuda") * 0.5).bfloat16()
for s in settings:
    for kv in s.split(","):
        if "=" in kv:
            k, v = kv.split("=")
            os.environ[k] = v
        else:
            os.environ["TWB200_SPLIT"] = kv
    m = B200WhisperForConditionalGeneration.from_hf(hf, dtype=torch.bfloat16, max_batch=B)
    prompt = m._init_tokens("zh", "transcribe", False)
    tracing = int(os.environ.get("TWB200_TRACE", "-1")) >= 0
    if not tracing:
        m.decode(enc, prompt, ML, False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 1 if tracing else 2
    for _ in range(n):
        toks, lens = m.decode(enc, prompt, ML, False)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    st = m.last_stage_ms()
    print(f"setting {s}: decode+crosskv {dt*1000:.1f} ms  (decode stage {st['decode']:.1f} ms, cross_kv {st['cross_kv']:.1f} ms)", flush=True)
    m.close()
    del m
    torch.cuda.empty_cache()
