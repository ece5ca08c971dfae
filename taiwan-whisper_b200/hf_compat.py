"""Offline construction of the HuggingFace objects the reference scripts hold.

The reference loads `WhisperForConditionalGeneration.from_pretrained(...)` from the Hub
(ref: training/run_pseudo_labelling.py:566-576, prefiltering/validator_inference.py:30); there is no
network here, so bench.py / tools / tests build the same object with random-init weights (HF init under
`torch.manual_seed`) and the generation config a released checkpoint ships as generation_config.json
(token ids: configs.py).  This is host-side plumbing of the product (the B200 model is created *from* such an
object); it contains no arithmetic of the path.
"""
from __future__ import annotations

import torch

from .configs import NON_SPEECH_TOKENS_MULTI, SHAPES, WhisperShape, token_ids


def build_hf_model(shape: WhisperShape | str, seed: int = 1234, dtype=torch.float32, suppress: bool = True,
                   attn_implementation: str = "eager"):
    from transformers import GenerationConfig, WhisperConfig, WhisperForConditionalGeneration

    if isinstance(shape, str):
        shape = SHAPES[shape]
    ids = token_ids(shape.vocab)
    cfg = WhisperConfig(
        vocab_size=shape.vocab, num_mel_bins=shape.n_mel, d_model=shape.d_model,
        encoder_layers=shape.enc_layers, decoder_layers=shape.dec_layers,
        encoder_attention_heads=shape.heads, decoder_attention_heads=shape.heads,
        encoder_ffn_dim=shape.ffn, decoder_ffn_dim=shape.ffn,
        max_source_positions=1500, max_target_positions=shape.max_target,
        pad_token_id=ids.pad, bos_token_id=ids.eos, eos_token_id=ids.eos, decoder_start_token_id=ids.sot,
        attn_implementation=attn_implementation,
    )
    torch.manual_seed(seed)
    model = WhisperForConditionalGeneration(cfg).eval()
    if dtype != torch.float32:
        model = model.to(dtype)
    gc = GenerationConfig(
        decoder_start_token_id=ids.sot, eos_token_id=ids.eos, pad_token_id=ids.pad, bos_token_id=ids.eos,
        max_length=shape.max_target,
    )
    gc.is_multilingual = True
    gc.lang_to_id = dict(ids.lang_to_id)
    gc.task_to_id = {"transcribe": ids.transcribe, "translate": ids.translate}
    gc.no_timestamps_token_id = ids.notimestamps
    gc.prev_sot_token_id = ids.startofprev
    gc.begin_suppress_tokens = [220, ids.eos]
    special = [ids.sot, ids.translate, ids.transcribe, ids.startofprev - 1, ids.startofprev, ids.nospeech]
    gc.suppress_tokens = sorted(set(NON_SPEECH_TOKENS_MULTI[:-4] + special)) if suppress else None
    gc.max_initial_timestamp_index = 50
    gc.alignment_heads = None
    model.generation_config = gc
    return model
