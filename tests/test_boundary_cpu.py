"""CPU (-m "not gpu"): the C-ABI library builds, loads and exports every symbol include/twb200.h
declares; host-side logic (prompt tokens, rules, batch packing, error behaviour) without a GPU."""
import os
import re
import types

import numpy as np
import pytest
import torch

from taiwan_whisper_b200 import lib as twlib
from taiwan_whisper_b200.build import build
from taiwan_whisper_b200.configs import SHAPES, token_ids

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def library():
    build()
    return twlib.load_library()


def test_header_symbols_exported(library):
    hdr = open(os.path.join(ROOT, "include", "twb200.h")).read()
    declared = re.findall(r"^TW_API [\w\s\*]+?\b(tw_\w+)\(", hdr, flags=re.M)
    assert len(declared) >= 14
    assert sorted(declared) == sorted(twlib.EXPORTS)
    for name in declared:
        assert hasattr(library, name), name
    assert library.tw_abi_version() == 1


def test_no_cpu_fallback_without_gpu(library):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(Exception) as e:
        twlib.Context(0)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_sass_is_blackwell_native():
    """the built library carries tcgen05 / TMA code (SASS mnemonics per B200_PROFILING.md)."""
    import shutil
    import subprocess
    build()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", twlib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnem in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnem in sass, mnem


def _fake_model(vocab=51866, multilingual=True):
    from taiwan_whisper_b200.host import B200WhisperForConditionalGeneration as M
    ids = token_ids(vocab)
    m = M.__new__(M)
    gc = types.SimpleNamespace(is_multilingual=multilingual, lang_to_id=dict(ids.lang_to_id),
                               task_to_id={"transcribe": ids.transcribe, "translate": ids.translate},
                               no_timestamps_token_id=ids.notimestamps, eos_token_id=ids.eos, pad_token_id=ids.pad,
                               suppress_tokens=[1, 2, 7], begin_suppress_tokens=[220, ids.eos], max_initial_timestamp_index=50)
    m.generation_config = gc
    m.config = types.SimpleNamespace(decoder_start_token_id=ids.sot)
    m.handle = None
    m.output_layout = "4.45"
    return m, ids


def test_init_tokens_match_reference_prompt():
    m, ids = _fake_model(51866)
    assert m._init_tokens("zh", "transcribe", False) == [50258, 50260, 50360, 50364]
    assert m._init_tokens("<|zh|>", "transcribe", True) == [50258, 50260, 50360]
    assert m._init_tokens("chinese", None, True) == [50258, 50260, 50360]
    m2, ids2 = _fake_model(51865)
    assert m2._init_tokens("zh", "transcribe", False) == [50258, 50260, 50359, 50363]
    with pytest.raises(ValueError):
        m._init_tokens("klingon", "transcribe", False)
    with pytest.raises(ValueError):
        m._init_tokens("zh", "summarise", False)
    m3, _ = _fake_model(51865, multilingual=False)
    with pytest.raises(ValueError):
        m3._init_tokens("zh", "transcribe", False)
    # language=None on a multilingual checkpoint would need HF's language detection: refuse, never guess <|en|>
    with pytest.raises(NotImplementedError):
        m._init_tokens(None, "transcribe", False)
    assert m3._init_tokens(None, None, False) == [50258, 50363 - 1 + 1]      # English-only: <|sot|><|notimestamps|>


def test_output_layout_argument():
    from taiwan_whisper_b200.host import B200WhisperForConditionalGeneration as M
    with pytest.raises(ValueError):
        M(hf_model=None, output_layout="4.46")


def test_rules_struct():
    m, ids = _fake_model(51866)
    r = m._rules(True)
    assert r["timestamp_begin"] == ids.timestamp_begin == 50365 and r["eos"] == 50257 and r["max_initial_ts"] == 50
    assert m._rules(False)["timestamp_begin"] is None
    s, keep = twlib.make_rules(r["suppress"], r["begin_suppress"], r["eos"], r["pad"], r["timestamp_begin"],
                               r["no_timestamps"], r["max_initial_ts"])
    assert s.n_suppress == 3 and s.n_begin_suppress == 2 and s.timestamp_begin == 50365
    assert [s.suppress[i] for i in range(3)] == [1, 2, 7]


def test_token_ids_tables():
    a, b = token_ids(51865), token_ids(51866)
    assert a.lang_to_id["<|zh|>"] == b.lang_to_id["<|zh|>"] == 50260      # ref: utils/test_hg_whisper.py:55-56
    assert len(a.lang_to_id) == 99 and len(b.lang_to_id) == 100 and b.lang_to_id["<|yue|>"] == 50358
    assert a.timestamp_begin == 50364 and b.timestamp_begin == 50365
    assert a.timestamp_begin + 1500 == 51864 and b.timestamp_begin + 1500 == 51865
    with pytest.raises(ValueError):
        token_ids(1000)


def test_feature_extractor_host_logic():
    from taiwan_whisper_b200.host import B200WhisperFeatureExtractor
    fe = B200WhisperFeatureExtractor(feature_size=128)
    assert fe.sampling_rate == 16000 and fe.model_input_names == ["input_features"] and fe.n_samples == 480000
    with pytest.raises(ValueError):
        fe(np.zeros(16000, np.float32), sampling_rate=8000)
    with pytest.raises(ValueError):
        B200WhisperFeatureExtractor(feature_size=64)
    with pytest.raises(NotImplementedError):
        B200WhisperFeatureExtractor(feature_size=80, hop_length=128)
    # batch packing (no kernel call): ragged rows, truncation at 30 s
    rows = [np.ones(10, np.float32), np.ones(500000, np.float32)]
    buf, nv = fe._to_device_batch(rows, "cpu")
    assert buf.shape == (2, 480000) and nv.tolist() == [10, 480000] and buf[0, 10:].abs().sum() == 0
    padded = fe.pad({"input_features": [np.zeros((128, 3000), np.float32)] * 3}, return_tensors="pt")
    assert padded["input_features"].shape == (3, 128, 3000)


def test_shapes_sheet():
    lv3 = SHAPES["large-v3"]
    assert (lv3.d_model, lv3.ffn, lv3.heads, lv3.enc_layers, lv3.dec_layers, lv3.n_mel, lv3.vocab) == \
        (1280, 5120, 20, 32, 32, 128, 51866)
    assert SHAPES["distil-large-v3"].dec_layers == 2     # ref: training/create_student_model.py:147-148
    for s in SHAPES.values():
        assert s.head_dim == 64


def test_retrieve_segments_matches_hf():
    """host seek-loop splitting == transformers' WhisperGenerationMixin._retrieve_segment on random windows."""
    from transformers.models.whisper.generation_whisper import WhisperGenerationMixin

    from taiwan_whisper_b200.host import retrieve_segments
    tsb = 50365
    rng = np.random.default_rng(5)
    cases = [[tsb, 5, 6, tsb + 10, tsb + 10, 7, tsb + 60], [tsb, 5, 6, tsb + 10, tsb + 10, 7, 8], [5, 6, 7],
             [tsb + 3], [tsb, tsb], [tsb, 9, tsb + 1500, tsb + 1500], [tsb, 9, tsb + 700, tsb + 700, tsb + 700, 3, tsb + 900]]
    for _ in range(200):
        n = int(rng.integers(1, 40))
        seq = np.where(rng.random(n) < 0.35, tsb + rng.integers(0, 1500, n), rng.integers(0, 50000, n))
        cases.append(seq.tolist())
    for seq in cases:
        for nframes in (3000, 1234):
            t = torch.tensor(seq)
            segs, off = WhisperGenerationMixin._retrieve_segment(
                seek_sequence=t, seek_outputs=[t], time_offset=torch.zeros(1, dtype=torch.float64), timestamp_begin=tsb,
                seek_num_frames=torch.tensor([nframes]), time_precision=0.02, time_precision_features=0.01, input_stride=2,
                prev_idx=0, idx=0, return_token_timestamps=False, decoder_input_ids=torch.zeros((1, 3), dtype=torch.long))
            mine, off2 = retrieve_segments(np.asarray(seq), tsb, nframes)
            assert int(off) == off2, seq
            assert [s["tokens"].tolist() for s in segs] == [m.tolist() for m in mine], seq


def test_chunk_plan_matches_reference_chunker():
    """window placement == ref training/flax/distil_whisper/pipeline.py:224-254 (restated inline here)."""
    from taiwan_whisper_b200.longform import chunk_plan

    def ref_plan(inputs_len, chunk_len, stride_left, stride_right):
        step = chunk_len - stride_left - stride_right
        starts = np.arange(0, inputs_len, step)
        out = []
        for st in starts:
            en = st + chunk_len
            sl = 0 if st == 0 else stride_left
            last = (en > inputs_len) if stride_right > 0 else (en >= inputs_len)
            out.append((min(en, inputs_len) - st, sl, 0 if last else stride_right))
        return starts, out

    for n in (1, 479_999, 480_000, 480_001, 1_000_000, 16000 * 600):
        for (sl, sr) in ((80000, 80000), (0, 0), (16000, 32000)):
            a, b = chunk_plan(n, 480000, sl, sr)
            c, d = ref_plan(n, 480000, sl, sr)
            assert np.array_equal(a, c) and b == d
    starts, strides = chunk_plan(16000 * 600)           # 10 minutes, default stride chunk/6 each side
    assert starts[1] - starts[0] == 320000 and strides[0][1] == 0 and strides[-1][2] == 0
    with pytest.raises(ValueError):
        chunk_plan(1000, 100, 60, 60)


def test_workspace_bytes_is_host_arithmetic():
    """tw_workspace_bytes needs no device: sizes a descriptor before loading."""
    from taiwan_whisper_b200 import lib as twlib
    from taiwan_whisper_b200.configs import SHAPES
    lib = twlib.load_library()
    sh = SHAPES["large-v3"]

    def nbytes(max_batch, dtype):
        desc = twlib.ModelDesc(sh.d_model, sh.ffn, sh.heads, sh.enc_layers, sh.dec_layers, sh.n_mel, sh.vocab, sh.max_target, dtype, max_batch)
        return int(lib.tw_workspace_bytes(desc))

    b64 = nbytes(64, twlib.TW_BF16)
    assert 25e9 < b64 < 40e9, b64                        # 3.1 GB weights + 15.7 GB cross-K/V + 4.7 GB cache pools + activations
    assert nbytes(128, twlib.TW_BF16) > b64 > nbytes(32, twlib.TW_BF16)
    assert nbytes(64, twlib.TW_F32) > 1.9 * b64 - 1e9
    assert lib.tw_workspace_bytes(None) == 0


def test_pipeline_balance_rule():
    """transcribe_batches(auto_sms=True): encoder-partition size from the stage times of the last group (host.py _balanced_sms)."""
    from taiwan_whisper_b200.host import B200WhisperForConditionalGeneration as M
    # inside the dead band (stage 1 between 0.75 and 1.02 of the decode): no change
    assert M._balanced_sms(24, 148, 2710.0, 3100.0) == 24
    assert M._balanced_sms(40, 148, 1600.0, 1700.0) == 40
    # stage 1 is the bottleneck: grow, in multiples of 8, never beyond half of the device
    assert M._balanced_sms(24, 148, 2700.0, 1500.0) == 48
    assert M._balanced_sms(24, 148, 9000.0, 1000.0) == 72
    # stage 1 idles: shrink, never below 16
    assert M._balanced_sms(48, 148, 800.0, 2500.0) == 24
    assert M._balanced_sms(24, 148, 1000.0, 3100.0) == 16
    assert M._balanced_sms(16, 148, 100.0, 3100.0) == 16
    # a stage that was not measured leaves the split alone
    assert M._balanced_sms(24, 148, -1.0, 3100.0) == 24
    # the rule converges: applying it to times that scale ~1 / SMs settles in one or two moves
    n, work, t_dec = 24, 24 * 2700.0, 1500.0
    for _ in range(4):
        n = M._balanced_sms(n, 148, work / n, t_dec)
    assert n == M._balanced_sms(n, 148, work / n, t_dec) and 0.75 <= (work / n) / t_dec <= 1.02
