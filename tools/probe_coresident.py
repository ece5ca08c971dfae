"""Probe: do the small-footprint decode-step kernels run NEXT TO a resident cross-attention streaming kernel?
A long K|V stream (B clips) runs on stream 1; a few GEMMs / self-attention launches go to stream 2 right after it starts.
If they finish long before the stream kernel does, their CTAs shared the SMs with it."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taiwan_whisper_b200 import lib as twlib  # noqa: E402

ctx = twlib.Context.get(0)
lib = ctx.lib
H, d = 20, 1280
Bk = int(os.environ.get("PROBE_BK", "192"))
lite = int(os.environ.get("PROBE_LITE", "1"))
kv = (torch.randn((Bk, 1500, 2 * d), device="cuda") * 0.3).bfloat16()
q = (torch.randn((Bk, d), device="cuda") * 0.3).bfloat16()
out = torch.empty((Bk, d), device="cuda", dtype=torch.bfloat16)
M = 32
A = (torch.randn((M, d), device="cuda") * 0.5).bfloat16()
W = (torch.randn((d, d), device="cuda") * 0.05).bfloat16()
Cc = torch.empty((M, d), device="cuda", dtype=torch.bfloat16)
Tc = 128
skv = (torch.randn((M, 448, 2 * d), device="cuda") * 0.3).bfloat16()
sq = (torch.randn((M, 3 * d), device="cuda") * 0.3).bfloat16()
sout = torch.empty((M, d), device="cuda", dtype=torch.bfloat16)
small = torch.zeros((32, 1280), device="cuda")
NW = 40                                     # distinct 13 MB weight matrices: every launch reads HBM-cold weights
Wbig = (torch.randn((NW, 5120, d), device="cuda") * 0.05).bfloat16()
Cbig = torch.empty((M, 5120), device="cuda", dtype=torch.bfloat16)
wi = [0]
pdl = int(os.environ.get("PROBE_PDL", "0"))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
lib.tw_debug_set_lite(lite)


def stream_kernel(st):
    ctx.check(lib.tw_debug_decode_attention(ctx.handle, q.data_ptr(), d, kv.data_ptr(), 1500 * 2 * d, 1500, Bk, H, twlib.TW_BF16,
                                            out.data_ptr(), st.cuda_stream))


def gemm(st, impl):
    ctx.check(lib.tw_debug_gemm(ctx.handle, A.data_ptr(), W.data_ptr(), None, Cc.data_ptr(), M, d, d, twlib.TW_BF16, 0, None, 1, impl,
                                st.cuda_stream))
    lib.tw_debug_set_lite(lite)          # the impl-4 path clears the flag


def gemm_cold(st):
    wi[0] = (wi[0] + 1) % NW
    lib.tw_debug_set_pdl(pdl)
    ctx.check(lib.tw_debug_gemm(ctx.handle, A.data_ptr(), Wbig[wi[0]].data_ptr(), None, Cbig.data_ptr(), M, 5120, d, twlib.TW_BF16, 1, None,
                                1, 4, st.cuda_stream))
    lib.tw_debug_set_pdl(0)
    lib.tw_debug_set_lite(lite)


def self_attn(st):
    ctx.check(lib.tw_debug_self_attention(ctx.handle, sq.data_ptr(), 3 * d, skv.data_ptr(), 448 * 2 * d, Tc, M, H, twlib.TW_BF16,
                                          sout.data_ptr(), st.cuda_stream))


def torch_small(st):
    with torch.cuda.stream(st):
        small.add_(1.0)


def ev():
    return torch.cuda.Event(enable_timing=True)


def alone(fn, n=5):
    fn(s2)
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record(s2)
    for _ in range(n):
        fn(s2)
    b.record(s2)
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1000 / n


def overlapped(fn, n=3):
    torch.cuda.synchronize()
    t0, t_small, t_big = ev(), ev(), ev()
    t0.record(s1)
    s2.wait_event(t0)
    stream_kernel(s1)
    t_big.record(s1)
    for _ in range(n):
        fn(s2)
    t_small.record(s2)
    torch.cuda.synchronize()
    return t0.elapsed_time(t_small) * 1000, t0.elapsed_time(t_big) * 1000


print(f"pdl={pdl} lite={lite} carveout={os.environ.get('TWB200_CARVEOUT', '1')}  stream kernel: {Bk} clips, alone {alone(stream_kernel, 3):.1f} us "
      f"({Bk * 1500 * 2 * d * 2 / alone(stream_kernel, 3) / 1e6:.2f} TB/s)")
for name, fn in (("gemm lite (impl 4)", lambda st: gemm(st, 4)), ("gemm lite cold 13MB", gemm_cold),
                 ("self-attention", self_attn), ("torch small add_", torch_small)):
    t_alone = alone(fn)
    for rep in range(2):
        ts, tb = overlapped(fn)
        print(f"  {name:22s} alone {t_alone:6.1f} us/launch | 3 launches behind a running stream kernel done at {ts:7.1f} us, "
              f"stream kernel done at {tb:7.1f} us  -> {'CO-RESIDENT' if ts < 0.7 * tb else 'waited'}")
