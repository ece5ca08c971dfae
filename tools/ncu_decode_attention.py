"""Four launches of the dominant kernel (decode cross-attention K/V streaming) at the bench shape
(large-v3: B=64 clips [argv 1; 128 = two merged batches], Tk=1500, H=20, bf16 => 491.5 MB per launch per 64 clips) for `ncu --set full`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taiwan_whisper_b200 import lib as twlib  # noqa: E402

ctx = twlib.Context.get(0)
B, Tk, H = (int(sys.argv[1]) if len(sys.argv) > 1 else 64), 1500, 20
d = H * 64
kvs = [torch.randn((B, Tk, 2 * d), device="cuda").bfloat16() for _ in range(2)]
q = (torch.randn((B, d), device="cuda") * 0.125).bfloat16()
out = torch.empty((B, d), device="cuda", dtype=torch.bfloat16)
for i in range(4):
    ctx.check(ctx.lib.tw_debug_decode_attention(ctx.handle, q.data_ptr(), d, kvs[i % 2].data_ptr(), Tk * 2 * d, Tk, B, H,
                                                twlib.TW_BF16, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("ok", float(out.float().abs().sum()))
