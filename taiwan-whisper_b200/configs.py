"""Model shape sheets and special-token ids for the Whisper teacher-inference path.

Shapes are the HF `WhisperConfig` values of the checkpoints the reference loads
(ref: training/run_pseudo_labelling.py:540-576, prefiltering/validator_inference.py:30,
training/create_student_model.py:139-150); token ids follow the multilingual vocabularies
(ref: utils/test_hg_whisper.py:55-56, utils/longform_eval.py:39-40,
prefiltering/validator_inference.py:36). Real checkpoints carry these in
generation_config.json; offline (no network, random-init weights) they are built here.
"""
from __future__ import annotations

from dataclasses import dataclass, field

N_SAMPLES = 480_000      # 30 s @ 16 kHz
N_FRAMES = 3000          # log-mel frames per window
N_CTX = 1500             # encoder positions
SAMPLING_RATE = 16_000

# transformers/models/whisper/configuration_whisper.py NON_SPEECH_TOKENS_MULTI
NON_SPEECH_TOKENS_MULTI = [
    1, 2, 7, 8, 9, 10, 14, 25, 26, 27, 28, 29, 31, 58, 59, 60, 61, 62, 63, 90, 91, 92, 93, 359, 503, 522, 542, 873,
    893, 902, 918, 922, 931, 1350, 1853, 1982, 2460, 2627, 3246, 3253, 3268, 3536, 3846, 3961, 4183, 4667, 6585, 6647,
    7273, 9061, 9383, 10428, 10929, 11938, 12033, 12331, 12562, 13793, 14157, 14635, 15265, 15618, 16553, 16604, 18362,
    18956, 20075, 21675, 22520, 26130, 26161, 26435, 28279, 29464, 31650, 32302, 32470, 36865, 42863, 47425, 49870,
    50254, 50258, 50360, 50361, 50362,
]

_LANGS = (
    "en zh de es ru ko fr ja pt tr pl ca nl ar sv it id hi fi vi he uk el ms cs ro da hu ta no th ur hr bg lt la mi "
    "ml cy sk te fa lv bn sr az sl kn et mk br eu is hy ne mn bs kk sq sw gl mr pa si km sn yo so af oc ka be tg sd "
    "gu am yi lo uz fo ht ps tk nn mt sa lb my bo tl mg as tt haw ln ha ba jw su"
).split()


@dataclass(frozen=True)
class WhisperShape:
    name: str
    n_mel: int
    d_model: int
    ffn: int
    heads: int
    enc_layers: int
    dec_layers: int
    vocab: int
    max_target: int = 448

    @property
    def head_dim(self) -> int:
        return self.d_model // self.heads


SHAPES = {
    # BASELINE.json configs[0]
    "tiny": WhisperShape("tiny", 80, 384, 1536, 6, 4, 4, 51865),
    "base": WhisperShape("base", 80, 512, 2048, 8, 6, 6, 51865),
    "small": WhisperShape("small", 80, 768, 3072, 12, 12, 12, 51865),
    # configs[4] validator
    "medium": WhisperShape("medium", 80, 1024, 4096, 16, 24, 24, 51865),
    # configs[1], configs[2]
    "large-v3": WhisperShape("large-v3", 128, 1280, 5120, 20, 32, 32, 51866),
    # configs[3]: large-v3 encoder + 2 decoder layers (ref: training/create_student_model.py:147-148)
    "distil-large-v3": WhisperShape("distil-large-v3", 128, 1280, 5120, 20, 32, 2, 51866),
    # test-only miniatures (same code paths, seconds on CPU)
    "micro80": WhisperShape("micro80", 80, 128, 256, 2, 2, 2, 51865),
    "micro128": WhisperShape("micro128", 128, 128, 512, 2, 2, 2, 51866),
    # test-only: the benched widths with few layers — large-v3 (configs[1]) and medium (configs[4]) at full d / heads / ffn /
    # vocabulary / mel bins, so that the parity tests run the kernels at the tile shapes the bench uses
    "lv3w": WhisperShape("lv3w", 128, 1280, 5120, 20, 2, 2, 51866),
    "medw": WhisperShape("medw", 80, 1024, 4096, 16, 2, 2, 51865),
}


@dataclass(frozen=True)
class TokenIds:
    vocab: int
    eos: int = 50257
    sot: int = 50258
    lang_to_id: dict = field(default_factory=dict)
    translate: int = 50358
    transcribe: int = 50359
    startofprev: int = 50361
    nospeech: int = 50362
    notimestamps: int = 50363

    @property
    def timestamp_begin(self) -> int:
        return self.notimestamps + 1

    @property
    def pad(self) -> int:
        return self.eos


def token_ids(vocab: int) -> TokenIds:
    """51865 = multilingual v1/v2 vocabulary; 51866 = large-v3 (adds <|yue|>, shifting the
    task/timestamp ids by one)."""
    if vocab == 51865:
        langs = _LANGS
        shift = 0
    elif vocab == 51866:
        langs = _LANGS + ["yue"]
        shift = 1
    else:
        raise ValueError(f"unsupported Whisper vocabulary size {vocab}")
    lang_to_id = {f"<|{l}|>": 50259 + i for i, l in enumerate(langs)}
    return TokenIds(
        vocab=vocab,
        lang_to_id=lang_to_id,
        translate=50358 + shift,
        transcribe=50359 + shift,
        startofprev=50361 + shift,
        nospeech=50362 + shift,
        notimestamps=50363 + shift,
    )
