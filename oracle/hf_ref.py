"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference's own implementation of the path, runnable offline: HuggingFace
`WhisperFeatureExtractor` + `WhisperForConditionalGeneration.generate`, exactly the two entry
points the reference calls (ref: training/run_pseudo_labelling.py:739,917-918;
prefiltering/validator_inference.py:57-60,78), with random-init weights
(`torch.manual_seed(seed)`, HF init) and a hand-built generation config (the real checkpoints
ship it as generation_config.json; there is no network here).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from taiwan_whisper_b200.configs import NON_SPEECH_TOKENS_MULTI, SHAPES, WhisperShape, token_ids  # noqa: E402


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


def build_hf_feature_extractor(n_mel: int):
    from transformers import WhisperFeatureExtractor
    return WhisperFeatureExtractor(feature_size=n_mel)


def hf_features(fe, pcm_f32: np.ndarray) -> np.ndarray:
    """[B,480000] f32 -> [B,n_mel,3000] f32 through the reference's call
    (ref: prefiltering/validator_inference.py:57-60)."""
    out = fe([np.asarray(r, dtype=np.float32) for r in pcm_f32], sampling_rate=16000, return_tensors="np")
    return np.asarray(out["input_features"], dtype=np.float32)


@torch.no_grad()
def hf_generate(model, feats: np.ndarray | torch.Tensor, max_length: int, return_timestamps: bool = False,
                language: str = "zh", task: str = "transcribe") -> np.ndarray:
    """The reference's generate call (ref: training/run_pseudo_labelling.py:864-876,917-918).
    transformers 5.5 returns generated tokens only (no forced prompt), right-padded."""
    feats = torch.as_tensor(feats).to(next(model.parameters()).dtype)
    ids = model.generate(feats, max_length=max_length, num_beams=1, return_timestamps=return_timestamps,
                         language=language, task=task)
    return ids.cpu().numpy().astype(np.int64)


@torch.no_grad()
def hf_encoder_states(model, feats: np.ndarray):
    """All encoder hidden states (pre-final-LN per layer, then the final LN output last)."""
    enc = model.model.encoder
    out = enc(torch.as_tensor(feats), output_hidden_states=True, return_dict=True)
    return [h.numpy() for h in out.hidden_states], out.last_hidden_state.numpy()
