"""probe: phase timestamps (globaltimer, ns) of the skinny tcgen05 GEMM, CTA 0 and a middle CTA.
Needs libtwb200 built with -DTW_GEMM_TRACE (tools/probes/build_trace.sh)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from taiwan_whisper_b200 import lib as twlib  # noqa: E402

ctx = twlib.Context.get(0)
lib = ctx.lib
dev = torch.device("cuda")
for (M, N, K, mode) in ((64, 3840, 1280, 0), (64, 1280, 1280, 2), (64, 5120, 1280, 1), (64, 1280, 5120, 2)):
    A = (torch.randn((M, K), device=dev) * 0.5).bfloat16()
    Ws = [(torch.randn((N, K), device=dev) * 0.05).bfloat16() for _ in range(4)]
    bias = torch.randn((N,), device=dev)
    Cc = torch.zeros((M, N), device=dev, dtype=torch.bfloat16 if mode in (0, 1) else torch.float32)
    for i in range(4):
        ctx.check(lib.tw_debug_gemm(ctx.handle, A.data_ptr(), Ws[i].data_ptr(), bias.data_ptr(), Cc.data_ptr(), M, N, K, twlib.TW_BF16,
                                    mode, None, 1, 1, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    t = torch.empty(256 * 8, dtype=torch.int64, device=dev)
    # copy from the device trace buffer via a torch view of raw memory
    lib.tw_debug_trace_copy.argtypes = [C.c_void_p]
    lib.tw_debug_trace_copy(C.c_void_p(t.data_ptr()))
    torch.cuda.synchronize()
    tr = t.cpu().view(256, 8)
    names = ["entry", "setup_done", "first_full", "last_mma_issued", "acc_ready", "epi_done", "exit", "-"]
    print(f"M={M} N={N} K={K} mode={mode}")
    base = int(tr[:, 0][tr[:, 0] > 0].min())
    for cta in (0, 20, 39):
        row = tr[cta]
        if int(row[0]) == 0:
            continue
        print("  cta", cta, " ".join(f"{names[i]}={int(row[i]) - base}" for i in range(7) if int(row[i]) > 0))
    last_exit = int(tr[:, 6].max()) - base
    print("  all CTAs done at", last_exit, "ns after the first entry")
