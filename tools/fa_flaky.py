import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taiwan_whisper_b200 import lib as twlib
ctx = twlib.Context.get(0)
def ref_attn(qkv, B, S, H):
    d = H * 64
    x = qkv.float().view(B, S, 3, H, 64)
    sc = torch.einsum("bqhd,bkhd->bhqk", x[:, :, 0], x[:, :, 1])
    return torch.einsum("bhqk,bkhd->bqhd", torch.softmax(sc, -1), x[:, :, 2]).reshape(B * S, d)
bad = 0
N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for it in range(N):
    for (B, S, H) in ((1, 1500, 20),):
        d = H * 64
        g = torch.Generator(device="cuda").manual_seed(it * 7 + S + H)
        qkv = torch.randn((B * S, 3 * d), device="cuda", generator=g)
        qkv[:, :d] *= 0.3
        qkv = qkv.bfloat16()
        out = torch.zeros((B * S, d), device="cuda", dtype=torch.bfloat16)
        ctx.check(ctx.lib.tw_debug_encoder_attention(ctx.handle, qkv.data_ptr(), out.data_ptr(), B, S, H, twlib.TW_BF16, 1, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        err = (out.float() - ref_attn(qkv, B, S, H)).abs().max().item()
        if err > 2e-2:
            bad += 1
            rows = ((out.float() - ref_attn(qkv, B, S, H)).abs().amax(1) > 2e-2).nonzero().flatten()
            print("bad", it, (B, S, H), err, "rows", rows[:8].tolist(), "n", rows.numel(), "cols", ((out.float() - ref_attn(qkv, B, S, H)).abs().amax(0) > 2e-2).nonzero().flatten()[:8].tolist())
print("failures", bad, "of", N)
