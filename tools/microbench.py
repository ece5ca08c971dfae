"""Kernel microbenchmarks through the C-ABI debug entry points (CUDA events, L2 flushed between
iterations by rotating over buffers larger than L2).  Usage: python tools/microbench.py [what ...]"""
import os
import sys
import json

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taiwan_whisper_b200 import lib as twlib  # noqa: E402

dev = torch.device("cuda:0")
ctx = twlib.Context.get(0)
st = lambda: torch.cuda.current_stream().cuda_stream
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
TF = 1637.4


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_decode_attention():
    H, d, B = 20, 1280, 64
    for Tk, nbuf in ((1500, 4), (128, 8), (16, 8)):
        kvs = [torch.randn((B, Tk, 2 * d), device=dev).bfloat16() for _ in range(nbuf)]     # 491 MB each at Tk=1500
        q = torch.randn((B, d), device=dev).bfloat16() * 0.125
        out = torch.empty((B, d), device=dev, dtype=torch.bfloat16)

        def fn(i):
            kv = kvs[i % nbuf]
            ctx.check(ctx.lib.tw_debug_decode_attention(ctx.handle, q.data_ptr(), d, kv.data_ptr(), Tk * 2 * d, Tk, B, H,
                                                        twlib.TW_BF16, out.data_ptr(), st()))
        ms = timeit(fn)
        bytes_ = B * Tk * 2 * d * 2
        # check against torch
        kv = kvs[0].float().view(B, Tk, 2, H, 64)
        s = torch.einsum("bhd,bthd->bht", q.float().view(B, H, 64), kv[:, :, 0])
        ref = torch.einsum("bht,bthd->bhd", torch.softmax(s, -1), kv[:, :, 1]).reshape(B, d)
        fn(0)
        torch.cuda.synchronize()
        err = (out.float() - ref).abs().max().item()
        print(f"decode_attention Tk={Tk}: {ms*1000:.1f} us (partial+combine)  {bytes_/ms/1e6:.0f} GB/s  {bytes_/ms/1e6/HBM:.2%} of HBM  maxerr {err:.2e}")
        del kvs


def bench_gemm():
    shapes = [(64, 1280, 1280), (64, 3840, 1280), (64, 5120, 1280), (64, 1280, 5120), (64, 51866, 1280),
              (96000, 1280, 1280), (96000, 3840, 1280), (96000, 5120, 1280), (96000, 1280, 5120), (96000, 2560, 1280)]
    for (M, N, K) in shapes:
        nbuf = 1 if M > 1000 else 6
        A = (torch.randn((M, K), device=dev) * 0.5).bfloat16()
        Ws = [(torch.randn((N, K), device=dev) * 0.05).bfloat16() for _ in range(nbuf)]
        bias = torch.randn((N,), device=dev)
        Cc = torch.empty((M, N), device=dev, dtype=torch.bfloat16)

        def fn(i):
            W = Ws[i % nbuf]
            ctx.check(ctx.lib.tw_debug_gemm(ctx.handle, A.data_ptr(), W.data_ptr(), bias.data_ptr(), Cc.data_ptr(), M, N, K,
                                            twlib.TW_BF16, 0, None, 1, 1, st()))
        ms = timeit(fn, iters=10)
        fl = 2.0 * M * N * K
        wbytes = N * K * 2
        print(f"gemm_tc M={M} N={N} K={K}: {ms*1000:.1f} us  {fl/ms/1e9:.0f} TFLOP/s ({fl/ms/1e9/TF:.1%})  W-stream {wbytes/ms/1e6:.0f} GB/s")
        del Ws, A, Cc


def bench_encoder_attention():
    B, S, H = 8, 1500, 20
    d = H * 64
    qkv = (torch.randn((B * S, 3 * d), device=dev) * 0.5).bfloat16()
    out = torch.empty((B * S, d), device=dev, dtype=torch.bfloat16)

    for impl, name in ((0, "simt"), (1, "tcgen05")):
        def fn(i):
            ctx.check(ctx.lib.tw_debug_encoder_attention(ctx.handle, qkv.data_ptr(), out.data_ptr(), B, S, H, twlib.TW_BF16, impl, st()))
        ms = timeit(fn, iters=5)
        fl = 4.0 * B * H * S * S * 64
        print(f"encoder_attention({name}) B={B}: {ms:.2f} ms  {fl/ms/1e9:.1f} TFLOP/s")


def bench_small_for_ncu():
    """A handful of launches of the decode-side kernels at large-v3 / B=64 shapes, for
    `ncu --metrics gpu__time_duration.sum` (per-launch device time without host overhead)."""
    B, d, H, ffn = 64, 1280, 20, 5120
    x = (torch.randn((B, d), device=dev) * 0.5).bfloat16()
    for (N, K) in ((3840, 1280), (1280, 1280), (5120, 1280), (1280, 5120), (51866, 1280)):
        A = (torch.randn((B, K), device=dev) * 0.5).bfloat16()
        W = (torch.randn((N, K), device=dev) * 0.05).bfloat16()
        Cc = torch.empty((B, N), device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            ctx.check(ctx.lib.tw_debug_gemm(ctx.handle, A.data_ptr(), W.data_ptr(), None, Cc.data_ptr(), B, N, K, twlib.TW_BF16, 0,
                                            None, 1, 1, st()))
        Cf = torch.zeros((B, N), device=dev)
        for mode, C_ in ((0, Cc), (2, Cf)):
            if K > 1280 and mode != 2:
                continue
            for _ in range(3):
                ctx.check(ctx.lib.tw_debug_gemm(ctx.handle, A.data_ptr(), W.data_ptr(), None, C_.data_ptr(), B, N, K, twlib.TW_BF16,
                                                mode, None, 1, 2, st()))
    for Tk in (8, 128, 256, 1500):
        kv = torch.randn((B, Tk, 2 * d), device=dev).bfloat16()
        q = torch.randn((B, d), device=dev).bfloat16() * 0.125
        out = torch.empty((B, d), device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            ctx.check(ctx.lib.tw_debug_decode_attention(ctx.handle, q.data_ptr(), d, kv.data_ptr(), Tk * 2 * d, Tk, B, H,
                                                        twlib.TW_BF16, out.data_ptr(), st()))
    torch.cuda.synchronize()
    print("small done")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "small_for_ncu":
    bench_small_for_ncu()
    sys.exit(0)

if __name__ == "__main__":
    what = sys.argv[1:] or ["decode_attention", "gemm", "encoder_attention"]
    for w in what:
        globals()["bench_" + w]()
