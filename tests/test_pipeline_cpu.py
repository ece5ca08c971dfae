"""CPU: host logic of the first-pass transcription pipeline — segment assembly from timestamp tokens and
the CSV wire format of ref pseudo-labelling/initial_inference.py:47-54 as read by ref pseudo-labelling/prepare_dataset.py:37-53."""
import csv

import pytest

from taiwan_whisper_b200.configs import token_ids
from taiwan_whisper_b200.pipeline import Segment, save_transcription_to_csv, segments_from_tokens

IDS = token_ids(51866)
TSB, EOS = IDS.timestamp_begin, IDS.eos


def ts(sec):
    return TSB + int(round(sec / 0.02))


def test_segments_pairs_and_offsets():
    toks = [ts(0.0), 11, 12, ts(2.5), ts(2.5), 13, ts(4.0), IDS.eos, IDS.eos]
    seg = segments_from_tokens(toks, TSB, EOS, window_start_s=30.0, window_len_s=30.0)
    assert seg == [(30.0, 32.5, [11, 12]), (32.5, 34.0, [13])]


def test_segment_without_closing_timestamp_runs_to_window_end():
    seg = segments_from_tokens([ts(1.0), 5, 6, 7], TSB, EOS, 0.0, 12.0)
    assert seg == [(1.0, 12.0, [5, 6, 7])]
    # text before any timestamp starts at the window start; timestamps past the window are clipped to its length
    seg = segments_from_tokens([5, ts(29.0)], TSB, EOS, 60.0, 10.0)
    assert seg == [(60.0, 70.0, [5])]


def test_special_tokens_and_empty_windows():
    assert segments_from_tokens([], TSB, EOS, 0.0, 30.0) == []
    assert segments_from_tokens([ts(0.0), ts(0.0), EOS], TSB, EOS, 0.0, 30.0) == []
    seg = segments_from_tokens([ts(0.0), IDS.sot, 9, IDS.notimestamps, ts(1.0)], TSB, EOS, 0.0, 30.0)
    assert seg == [(0.0, 1.0, [9])]


def test_csv_wire_format(tmp_path):
    p = tmp_path / "a.csv"
    save_transcription_to_csv([Segment(0.251, 18.909, "你好, world", [1]), {"start": "19.00", "end": "20.50", "text": "x"}], str(p))
    lines = p.read_text(encoding="utf-8").splitlines()
    assert lines[0] == "start,end,text" and lines[1] == '0.25,18.91,"你好, world"'
    # the reference's reader (prepare_dataset.py:37-53): csv.reader, skip header, 3 fields, float(start), float(end), text.strip()
    with open(p, "r", encoding="utf-8") as f:
        r = csv.reader(f)
        next(r)
        rows = [(float(a), float(b), c.strip()) for a, b, c in r]
    assert rows == [(0.25, 18.91, "你好, world"), (19.0, 20.5, "x")]


def test_pipeline_argument_errors():
    from taiwan_whisper_b200.pipeline import B200BatchedInferencePipeline
    with pytest.raises(NotImplementedError):
        B200BatchedInferencePipeline(model=None, use_vad_model=True)
    with pytest.raises(ValueError):
        B200BatchedInferencePipeline(model=None, chunk_length=45)


def test_stitch_windows_matches_hf():
    """Long-form stitching of overlapping windows == HF `_find_longest_common_sequence` (what `tokenizer._decode_asr`, called at
    ref training/flax/distil_whisper/pipeline.py:353-375, uses in no-timestamp mode), on synthetic overlapping streams with
    substitution noise in the overlaps."""
    import numpy as np
    from transformers.models.whisper.tokenization_whisper import _find_longest_common_sequence
    from taiwan_whisper_b200.longform import stitch_windows
    rng = np.random.default_rng(5)
    for case in range(40):
        truth = rng.integers(0, 300, size=int(rng.integers(40, 160))).tolist()
        seqs, pos = [], 0
        while pos < len(truth):
            n = int(rng.integers(12, 40))
            w = truth[pos:pos + n]
            w = [t if rng.random() > 0.1 else int(rng.integers(0, 300)) for t in w]      # decoding noise
            seqs.append(w)
            if pos + n >= len(truth):
                break
            pos += n - int(rng.integers(0, 10))                                          # overlap of 0..9 tokens
        assert stitch_windows(seqs) == _find_longest_common_sequence(seqs), case
    assert stitch_windows([]) == [] and stitch_windows([[1, 2, 3]]) == [1, 2, 3]
    assert stitch_windows([[1, 2, 3, 4], [3, 4, 5]]) == [1, 2, 3, 4, 5]


def test_segments_properties_random_streams():
    """Size-independent properties on random token streams: every text token is kept exactly once and in order, times are
    ordered, inside the window, and segments do not overlap."""
    import numpy as np
    rng = np.random.default_rng(11)
    for case in range(200):
        n = int(rng.integers(0, 60))
        toks = []
        for _ in range(n):
            u = rng.random()
            if u < 0.25:
                toks.append(ts(float(rng.integers(0, 1500)) * 0.02))
            elif u < 0.30:
                toks.append(int(rng.integers(EOS, TSB)))            # special token
            else:
                toks.append(int(rng.integers(0, EOS)))
        w0, wl = 30.0 * case, float(rng.integers(1, 31))
        segs = segments_from_tokens(toks, TSB, EOS, w0, wl)
        assert [t for s in segs for t in s[2]] == [t for t in toks if t < EOS]
        for s0, s1, ids in segs:
            assert ids and w0 <= s0 <= w0 + wl + 1e-9 and w0 <= s1 <= w0 + wl + 1e-9


from tests.helpers import StubWhisperTokenizer


def _StubTokenizer():
    return StubWhisperTokenizer(IDS)


def _random_window_tokens(rng, t_lo, t_hi):
    """a plausible timestamp-mode window: <|t0|> text <|t1|><|t1|> text <|t2|> ... with times inside [t_lo, t_hi] seconds"""
    toks, t = [], t_lo
    while t < t_hi - 0.5 and len(toks) < 60:
        toks.append(ts(t))
        toks += [int(x) for x in rng.integers(0, 300, size=int(rng.integers(1, 6)))]
        t = min(t_hi, t + float(rng.integers(25, 400)) * 0.02)
        if rng.random() < 0.15:
            break                                   # window ends without a closing timestamp
        toks.append(ts(t))
    return toks


def test_stitch_windows_timestamps_matches_hf_decode_asr():
    """Timestamp-aware stitching == HF `_decode_asr(..., return_timestamps=True)` (called at ref training/flax/distil_whisper/
    pipeline.py:353-375) on random overlapping windows: strides in seconds, timestamps inside the strides, windows that end
    without a closing timestamp, special tokens, and the seek-loop case of several segments inside one window's ids."""
    import numpy as np
    import torch
    from transformers.models.whisper.tokenization_whisper import _decode_asr
    from taiwan_whisper_b200.longform import stitch_windows_timestamps
    rng = np.random.default_rng(17)
    tok = _StubTokenizer()
    n_checked = 0
    for case in range(150):
        n_win = int(rng.integers(1, 6))
        outs = []
        for w in range(n_win):
            sl = 0.0 if w == 0 else 5.0
            sr = 0.0 if w == n_win - 1 else 5.0
            chunk_len = 30.0 if w < n_win - 1 else float(rng.integers(8, 31))
            toks = _random_window_tokens(rng, 0.0, chunk_len)
            if rng.random() < 0.2:                  # long-form generate(): a second 30 s segment concatenated in the same ids
                toks += _random_window_tokens(rng, 0.0, 20.0)
            if rng.random() < 0.3:
                toks.insert(int(rng.integers(0, len(toks) + 1)), int(rng.integers(EOS, TSB)))       # a stray special token
            o = {"tokens": toks}
            if case % 5 != 4:                       # every fifth case: no stride entries at all
                o["stride"] = (chunk_len, sl, sr)
            outs.append(o)
        hf_in = [{**o, "tokens": torch.tensor([o["tokens"]])} for o in outs]
        text, opt = _decode_asr(tok, hf_in, return_timestamps=True, return_language=False, time_precision=0.02)
        mine = stitch_windows_timestamps(outs, TSB, tok.all_special_ids)
        assert [c["timestamp"] for c in mine] == [c["timestamp"] for c in opt["chunks"]], case
        assert [tok.decode(c["tokens"]) for c in mine] == [c["text"] for c in opt["chunks"]], case
        n_checked += len(mine)
    assert n_checked > 300
    # prompt stripping: a leading <|startofprev|> ... <|startoftranscript|> prefix is dropped, as HF's _strip_prompt does
    o = [{"tokens": [IDS.startofprev, 7, 8, IDS.sot, ts(0.0), 5, ts(1.0)]}]
    got = stitch_windows_timestamps(o, TSB, tok.all_special_ids, prompt_token_id=IDS.startofprev, decoder_start_token_id=IDS.sot)
    assert got == [{"timestamp": (0.0, 1.0), "tokens": [5]}]


def test_segments_drop_right_padding():
    """generate() right-pads a batch with pad_token_id; with pad != eos (50256 in large-v3's config.json) the padding is a
    text-range id: it must neither be appended as text nor open a trailing segment."""
    pad = 50256
    row = [ts(0.0), 5, 6, ts(2.0)] + [pad] * 7
    assert segments_from_tokens(row, TSB, EOS, 0.0, 30.0, pad=pad) == [(0.0, 2.0, [5, 6])]
    assert segments_from_tokens(row, TSB, EOS, 0.0, 30.0, length=4) == [(0.0, 2.0, [5, 6])]
    assert segments_from_tokens([pad] * 5, TSB, EOS, 0.0, 30.0, pad=pad) == []
    # without the hint the old behaviour is visible (documented hazard)
    assert segments_from_tokens(row, TSB, EOS, 0.0, 30.0)[-1][2] == [pad] * 7
