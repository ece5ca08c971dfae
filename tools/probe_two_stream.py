"""Probe: does decoding two half-batches (32 + 32 clips) concurrently on two streams beat one 64-clip decode?
(latency-bound small kernels of one half could overlap the bandwidth-bound K/V stream of the other)"""
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taiwan_whisper_b200.hf_compat import build_hf_model  # noqa: E402
from taiwan_whisper_b200.configs import SHAPES  # noqa: E402
from taiwan_whisper_b200.host import B200WhisperForConditionalGeneration  # noqa: E402

sh = SHAPES["large-v3"]
with torch.device("cuda"):
    hf = build_hf_model(sh, seed=1234)
ML = 256
full = B200WhisperForConditionalGeneration.from_hf(hf, dtype=torch.bfloat16, max_batch=64)
prompt = full._init_tokens("zh", "transcribe", False)
enc = (torch.randn((64, 1500, 1280), device="cuda") * 0.5).bfloat16()


def timed(fn, n=2):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


t_full = timed(lambda: full.decode(enc, prompt, ML, False))
print(f"one model, B=64: {t_full*1000:.1f} ms")
full.close()
torch.cuda.empty_cache()
halves = [B200WhisperForConditionalGeneration.from_hf(hf, dtype=torch.bfloat16, max_batch=32) for _ in range(2)]
del hf
streams = [torch.cuda.Stream() for _ in range(2)]
encs = [enc[:32].contiguous(), enc[32:].contiguous()]


def seq():
    for m, e in zip(halves, encs):
        m.decode(e, prompt, ML, False)


def conc():
    def work(i):
        with torch.cuda.stream(streams[i]):
            halves[i].decode(encs[i], prompt, ML, False)
    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for s in streams:
        s.synchronize()


print(f"two models, B=32 each, sequential: {timed(seq)*1000:.1f} ms")
print(f"two models, B=32 each, two streams: {timed(conc)*1000:.1f} ms")
