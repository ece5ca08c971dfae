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

from taiwan_whisper_b200.configs import NON_SPEECH_TOKENS_MULTI, SHAPES, WhisperShape, token_ids  # noqa: E402,F401


from taiwan_whisper_b200.hf_compat import build_hf_model  # noqa: E402,F401  (re-exported for the tests)


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


@torch.no_grad()
def hf_teacher_forced_argmax(model, feats, prompt, forced, rules, device="cpu", dtype=torch.float32):
    """The reference implementation's per-position greedy choice when it is fed a given token history: one HF forward
    (`model(input_features, decoder_input_ids).logits`, modeling_whisper.py:1000-1100) over [prompt + forced[:-1]] at `dtype` on
    `device`, then the logits rules (oracle apply_rules = HF's processors) and argmax at every generated position.
    feats [n, n_mel, 3000] f32, forced [n, n_gen] int.  Returns int64 [n, n_gen].  This is how "HF's own bf16" is scored
    against the fp32 token stream on the same positions."""
    import copy

    from oracle import whisper_np
    forced = np.asarray(forced)
    n, n_gen = forced.shape
    P = len(prompt)
    m = copy.deepcopy(model).to(device=device, dtype=dtype).eval()
    dec_in = np.concatenate([np.tile(np.asarray(prompt)[None], (n, 1)), forced[:, :-1]], axis=1)
    out = np.zeros((n, n_gen), dtype=np.int64)
    for i in range(n):                      # row by row: bounded memory for [T, V] fp32 logits
        lg = m(input_features=torch.as_tensor(feats[i:i + 1]).to(device=device, dtype=dtype),
               decoder_input_ids=torch.as_tensor(dec_in[i:i + 1]).to(device)).logits[0, P - 1:].float().cpu().numpy()
        for s in range(n_gen):
            out[i, s] = int(np.argmax(whisper_np.apply_rules(lg[s], forced[i, :s].tolist(), rules)))
    del m
    return out
