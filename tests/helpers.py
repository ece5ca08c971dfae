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


class StubWhisperTokenizer:
    """Duck-typed tokenizer for HF's module-level `_decode_asr` (needs no vocabulary files): text of a token list is the
    comma-joined ids, so the chunks' token lists can be read back."""

    def __init__(self, ids):
        self.ids = ids
        self.all_special_ids = list(range(ids.eos, ids.timestamp_begin))

    def convert_tokens_to_ids(self, tok):
        return {"<|notimestamps|>": self.ids.notimestamps, "<|startofprev|>": self.ids.startofprev,
                "<|startoftranscript|>": self.ids.sot}[tok]

    def _strip_prompt(self, token_ids, prompt_token_id, decoder_start_token_id):
        if token_ids and token_ids[0] == prompt_token_id:
            return token_ids[token_ids.index(decoder_start_token_id):] if decoder_start_token_id in token_ids else []
        return token_ids

    def decode(self, ids):
        return "".join(f"{int(t)}," for t in ids)
