"""Helpers for the -m gpu parity tests (CUDA path through the C ABI vs the oracle)."""
import ctypes as C
import functools

import numpy as np
import torch

from oracle import hf_ref, logmel_np, whisper_np
from taiwan_whisper_b200 import lib as twlib
from taiwan_whisper_b200.configs import SHAPES
from taiwan_whisper_b200.synth import dequantise, synth_batch
from tests.helpers import default_rules, prompt_ids, weights_np


@functools.lru_cache(maxsize=None)
def hf_model(shape_name, seed=1234):
    return hf_ref.build_hf_model(SHAPES[shape_name], seed=seed)


@functools.lru_cache(maxsize=None)
def b200_model(shape_name, dtype_name, max_batch=4, seed=1234):
    from taiwan_whisper_b200.host import B200WhisperForConditionalGeneration
    dtype = {"f32": torch.float32, "bf16": torch.bfloat16}[dtype_name]
    # the oracle is transformers 5.x (generated ids only, seek loop on): the tests ask for that layout explicitly
    return B200WhisperForConditionalGeneration.from_hf(hf_model(shape_name, seed), dtype=dtype, max_batch=max_batch, output_layout="5.x")


@functools.lru_cache(maxsize=None)
def oracle_run(shape_name, n_clips, max_length, timestamps=False, seed=1234):
    """Oracle (numpy restatement) outputs: mel, taps, enc, tokens, per-step post-rules logits."""
    sh = SHAPES[shape_name]
    W = weights_np(hf_model(shape_name, seed))
    pcm = synth_batch(0, n_clips)
    mel = logmel_np.log_mel(dequantise(pcm), sh.n_mel)
    out = []
    for b in range(n_clips):
        taps = []
        enc = whisper_np.encoder_forward(W, mel[b], sh.heads, sh.enc_layers, taps)
        lt = []
        toks = whisper_np.greedy_decode(W, enc, prompt_ids(sh.vocab, timestamps), default_rules(sh.vocab, timestamps),
                                        max_length, sh.heads, sh.dec_layers, logits_tap=lt)
        out.append(dict(taps=taps, enc=enc, tokens=toks, logits=lt))
    return pcm, mel, out


def gemm_debug(A, W, bias, mode, use_tc, out_dtype=None, C_init=None, pos=None, period=1):
    """Calls tw_debug_gemm on cuda tensors; returns C."""
    ctx = twlib.Context.get(torch.cuda.current_device())
    M, K = A.shape
    N = W.shape[0]
    dt = twlib.TW_BF16 if A.dtype == torch.bfloat16 else twlib.TW_F32
    if mode in (0, 1):
        Cc = torch.empty((M, N), dtype=A.dtype, device=A.device)
    else:
        Cc = C_init.clone() if C_init is not None else torch.zeros((M, N), dtype=torch.float32, device=A.device)
    ctx.check(ctx.lib.tw_debug_gemm(ctx.handle, A.data_ptr(), W.data_ptr(), bias.data_ptr() if bias is not None else None,
                                    Cc.data_ptr(), M, N, K, dt, mode, pos.data_ptr() if pos is not None else None, period,
                                    int(use_tc), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return Cc
