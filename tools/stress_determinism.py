"""Race detector: every kernel without floating-point atomics must return bit-identical output when launched again on the same
input.  Runs each kernel-level entry point N times at shapes that fill the GPU (two CTAs per SM where the kernel allows it) and
counts launches whose output differs from the first.  usage: stress_determinism.py [N]   (found the two tensor-memory hazards
of the flash-attention kernel this way: ~1 % of the launches at 20 heads x 1500 keys)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taiwan_whisper_b200 import lib as twlib  # noqa: E402
from taiwan_whisper_b200.host import log_mel  # noqa: E402

ctx = twlib.Context.get(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
g = torch.Generator(device="cuda").manual_seed(7)
results = {}


def check(name, fn):
    ref = fn().clone()
    torch.cuda.synchronize()
    bad = 0
    for _ in range(N):
        out = fn()
        torch.cuda.synchronize()
        if not torch.equal(out, ref):
            bad += 1
    results[name] = bad
    print(f"{name:55s} {bad} of {N} launches differ", flush=True)


def gemm(M, Nn, K, mode, impl):
    A = (torch.randn((M, K), device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn((Nn, K), device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn((Nn,), device="cuda", generator=g) * 0.1
    C = torch.empty((M, Nn), device="cuda", dtype=torch.bfloat16 if mode in (0, 1) else torch.float32)

    def fn():
        ctx.check(ctx.lib.tw_debug_gemm(ctx.handle, A.data_ptr(), W.data_ptr(), bias.data_ptr(), C.data_ptr(), M, Nn, K, twlib.TW_BF16, mode,
                                        None, 1, impl, st()))
        return C
    return fn


check("gemm_tc 12000x3840x1280 store", gemm(12000, 3840, 1280, 0, 1))
check("gemm_tc 12000x5120x1280 gelu", gemm(12000, 5120, 1280, 1, 1))
check("gemm_tc 12000x1280x5120 store", gemm(12000, 1280, 5120, 0, 1))
check("gemm_tc_skinny 64x3840x1280 store", gemm(64, 3840, 1280, 0, 3))
check("gemm_tc_skinny 64x5120x1280 gelu", gemm(64, 5120, 1280, 1, 3))
check("gemm_tc_skinny 64x51866x1280 f32", gemm(64, 51866, 1280, 4, 3))


def enc_attn(B, S, H):
    d = H * 64
    qkv = torch.randn((B * S, 3 * d), device="cuda", generator=g)
    qkv[:, :d] *= 0.3
    qkv = qkv.bfloat16()
    out = torch.empty((B * S, d), device="cuda", dtype=torch.bfloat16)

    def fn():
        ctx.check(ctx.lib.tw_debug_encoder_attention(ctx.handle, qkv.data_ptr(), out.data_ptr(), B, S, H, twlib.TW_BF16, 1, st()))
        return out
    return fn


check("encoder_attention_tc 1x1500x20", enc_attn(1, 1500, 20))
check("encoder_attention_tc 4x1500x20", enc_attn(4, 1500, 20))
check("encoder_attention_tc 3x1500x6", enc_attn(3, 1500, 6))


def causal_attn(B, S, H):
    d = H * 64
    q = (torch.randn((B * S, d), device="cuda", generator=g) * 0.3).bfloat16()
    kv = torch.randn((B * S, 2 * d), device="cuda", generator=g).bfloat16()
    out = torch.empty((B * S, d), device="cuda", dtype=torch.bfloat16)

    def fn():
        ctx.check(ctx.lib.tw_debug_attention(ctx.handle, q.data_ptr(), d, 0, kv.data_ptr(), 2 * d, 0, d, out.data_ptr(), B, S, S, H, twlib.TW_BF16,
                                             1, 1, st()))
        return out
    return fn


check("attention_tc causal 16x448x20", causal_attn(16, 448, 20))


def dec_attn(entry, B, Tk, H):
    d = H * 64
    kv = torch.randn((B, Tk, 2 * d), device="cuda", generator=g).bfloat16()
    q = (torch.randn((B, d), device="cuda", generator=g) * 0.3).bfloat16()
    out = torch.empty((B, d), device="cuda", dtype=torch.bfloat16)

    def fn():
        ctx.check(getattr(ctx.lib, entry)(ctx.handle, q.data_ptr(), d, kv.data_ptr(), Tk * 2 * d, Tk, B, H, twlib.TW_BF16, out.data_ptr(), st()))
        return out
    return fn


check("decode_attention_stream 64x1500x20", dec_attn("tw_debug_decode_attention", 64, 1500, 20))
check("decode_attention_stream 7x1500x6", dec_attn("tw_debug_decode_attention", 7, 1500, 6))
check("self_attention_decode 64x200x20", dec_attn("tw_debug_self_attention", 64, 200, 20))


def absorbed(B, Tk, H):
    d = H * 64
    enc = torch.randn((B * Tk, d), device="cuda", generator=g).bfloat16()
    qt = torch.zeros((B * H + 24, d), device="cuda", dtype=torch.bfloat16)
    qt[:B * H] = (torch.randn((B * H, d), device="cuda", generator=g) * 4.0 / d ** 0.5).bfloat16()
    out = torch.empty((B, H * d), device="cuda", dtype=torch.bfloat16)

    def fn():
        ctx.check(ctx.lib.tw_debug_absorbed_attention(ctx.handle, qt.data_ptr(), enc.data_ptr(), Tk, B, H, out.data_ptr(), None, None, 0, st()))
        return out
    return fn


check("absorbed_attention 64x1500x20", absorbed(64, 1500, 20))

pcm = torch.randint(-20000, 20000, (16, 480000), device="cuda", generator=g, dtype=torch.int16)
check("logmel 16 clips x 128 mel", lambda: log_mel(pcm, None, 128))
print("TOTAL", sum(results.values()))
