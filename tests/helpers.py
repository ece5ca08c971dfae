"""Shared test helpers: HF weights -> numpy dict, default rules for the oracle."""
import numpy as np

from taiwan_whisper_b200.configs import NON_SPEECH_TOKENS_MULTI, token_ids


def weights_np(hf_model):
    return {k: v.detach().float().numpy() for k, v in hf_model.state_dict().items()}


def default_rules(vocab, timestamps=False, suppress=True):
    ids = token_ids(vocab)
    special = [ids.sot, ids.translate, ids.transcribe, ids.startofprev - 1, ids.startofprev, ids.nospeech]
    return dict(
        suppress=sorted(set(NON_SPEECH_TOKENS_MULTI[:-4] + special)) if suppress else [],
        begin_suppress=[220, ids.eos], eos=ids.eos,
        ts_begin=ids.timestamp_begin if timestamps else None,
        no_timestamps=ids.notimestamps, max_initial_ts=50,
    )


def prompt_ids(vocab, timestamps=False, language="zh"):
    ids = token_ids(vocab)
    p = [ids.sot, ids.lang_to_id[f"<|{language}|>"], ids.transcribe]
    if not timestamps:
        p.append(ids.notimestamps)
    return p
