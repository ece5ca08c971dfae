"""Generates the committed golden fixtures in tests/golden/ from the reference's own
implementation (HuggingFace transformers — the third-party dependency that executes the hot
path's arithmetic for ref: training/run_pseudo_labelling.py:739,917-918 and
prefiltering/validator_inference.py:57-60,78).  Run in the build container:

    python tests/golden/make_golden.py

Image versions at generation time are recorded in each fixture (transformers 5.5.0, torch 2.11;
the reference pins 4.45.2 / 2.5.0 — SURVEY.md §0.4).  The reference itself holds no golden
vectors for this path, so these are "outputs of the reference run here".
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import transformers  # noqa: E402
from transformers.generation.logits_process import (  # noqa: E402
    SuppressTokensAtBeginLogitsProcessor, SuppressTokensLogitsProcessor, WhisperTimeStampLogitsProcessor)

from oracle import hf_ref  # noqa: E402
from taiwan_whisper_b200.configs import SHAPES, token_ids  # noqa: E402
from taiwan_whisper_b200.synth import dequantise, edge_case_clips, synth_batch  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
VERS = np.array([transformers.__version__, torch.__version__, np.__version__])


def golden_logmel():
    clips = {f"synth{i}": c for i, c in enumerate(synth_batch(0, 2))}
    clips.update(edge_case_clips())
    for n_mel in (80, 128):
        fe = hf_ref.build_hf_feature_extractor(n_mel)
        rec = {"versions": VERS, "names": np.array(list(clips))}
        for name, pcm in clips.items():
            f = hf_ref.hf_features(fe, dequantise(pcm)[None])[0]
            rec[name + "_sub"] = f[:, ::16].copy()                       # strided sample
            rec[name + "_stats"] = np.array([f.astype(np.float64).sum(), (f.astype(np.float64) ** 2).sum(),
                                             f.min(), f.max()])
        np.savez_compressed(os.path.join(OUT, f"logmel_{n_mel}.npz"), **rec)


def golden_model(shape_name, max_length=48):
    sh = SHAPES[shape_name]
    model = hf_ref.build_hf_model(sh, seed=1234)
    fe = hf_ref.build_hf_feature_extractor(sh.n_mel)
    pcm = synth_batch(0, 2)
    feats = hf_ref.hf_features(fe, dequantise(pcm))
    hs, last = hf_ref.hf_encoder_states(model, feats)
    rec = {"versions": VERS, "max_length": np.array(max_length)}
    rec["stem_sub"] = hs[0][:, ::50, ::8].copy()                        # residual stream after conv stem + pos
    rec["layer1_sub"] = hs[1][:, ::50, ::8].copy()
    rec["enc_sub"] = last[:, ::50, ::8].copy()
    rec["enc_norm"] = np.array([np.linalg.norm(last[b].astype(np.float64)) for b in range(2)])
    rec["tokens_nots"] = hf_ref.hf_generate(model, feats, max_length, return_timestamps=False)
    rec["tokens_ts_seekloop"] = hf_ref.hf_generate(model, feats, max_length, return_timestamps=True)
    np.savez_compressed(os.path.join(OUT, f"model_{shape_name}.npz"), **rec)


def teacher_labels(vocab, n_rows=2, T=24, seed=77):
    """Label rows like the distillation collator makes them (ref knowledge-distillation/run_distillation.py:470-512):
    prompt + text tokens, right-padded with -100."""
    ids = token_ids(vocab)
    rng = np.random.default_rng(seed + vocab)
    lab = np.full((n_rows, T), -100, dtype=np.int64)
    for b in range(n_rows):
        n = T - 5 * b
        row = [ids.lang_to_id["<|zh|>"], ids.transcribe, ids.notimestamps] + rng.integers(0, 50000, n - 4).tolist() + [ids.eos]
        lab[b, :n] = row
    return lab


def golden_teacher(shape_name):
    """HF teacher forward: model(input_features, labels).logits — the call of the distillation step
    (ref knowledge-distillation/run_distillation.py:1543-1577)."""
    sh = SHAPES[shape_name]
    model = hf_ref.build_hf_model(sh, seed=1234)
    fe = hf_ref.build_hf_feature_extractor(sh.n_mel)
    feats = hf_ref.hf_features(fe, dequantise(synth_batch(0, 2)))
    labels = teacher_labels(sh.vocab)
    with torch.no_grad():
        out = model(input_features=torch.as_tensor(feats), labels=torch.as_tensor(labels))
    lg = out.logits.float().numpy()
    from transformers.models.whisper.modeling_whisper import shift_tokens_right
    dec_ids = shift_tokens_right(torch.as_tensor(labels), model.config.pad_token_id, model.config.decoder_start_token_id).numpy()
    rec = {"versions": VERS, "labels": labels, "decoder_input_ids": dec_ids,
           "logits_sub": lg[:, :, ::97].copy(), "argmax": lg.argmax(-1),
           "row_norm": np.linalg.norm(lg.astype(np.float64), axis=-1)}
    np.savez_compressed(os.path.join(OUT, f"teacher_{shape_name}.npz"), **rec)


def rules_case_logits(vocab, ci, variant, tsb):
    """Logits of a rules case are regenerated from (vocab, case, variant) — not stored."""
    rng = np.random.default_rng(vocab * 1000 + ci * 10 + variant)
    logits = rng.standard_normal(vocab).astype(np.float32) * np.float32(0.3)
    if variant == 1:
        logits[tsb:] += np.float32(1.5)          # timestamp mass dominates
    if variant == 2:
        logits[tsb:] -= np.float32(6.0)          # text dominates
    return logits


def golden_rules():
    """HF logits processors on crafted histories — pins oracle.whisper_np.apply_rules."""
    rec = {"versions": VERS}
    for vocab in (51865, 51866):
        ids = token_ids(vocab)
        model = hf_ref.build_hf_model(SHAPES["micro80" if vocab == 51865 else "micro128"])
        gc = model.generation_config
        tsb = ids.timestamp_begin
        begin_index = 3
        hists = [
            [], [tsb + 3], [tsb + 3, 1000], [tsb + 3, 1000, 2000], [tsb + 3, 1000, tsb + 40],
            [tsb + 3, 1000, tsb + 40, tsb + 40], [tsb + 3, 1000, tsb + 40, tsb + 40, 77],
            [tsb, tsb], [tsb + 10, 5, 6, 7, tsb + 1500], [tsb + 10, 5, 6, 7, tsb + 1500, tsb + 1500], [400], [400, 401],
        ]
        procs_ts = [SuppressTokensAtBeginLogitsProcessor(gc.begin_suppress_tokens, begin_index),
                    SuppressTokensLogitsProcessor(gc.suppress_tokens),
                    WhisperTimeStampLogitsProcessor(gc, begin_index=begin_index)]
        procs_nots = procs_ts[:2]
        cases = []
        for ci, h in enumerate(hists):
            for variant in range(3):
                logits = rules_case_logits(vocab, ci, variant, tsb)
                prefix = [ids.sot, ids.lang_to_id["<|zh|>"], ids.transcribe]
                inp = torch.tensor([prefix + h])
                for mode, procs in (("ts", procs_ts), ("nots", procs_nots)):
                    s = torch.from_numpy(logits[None].copy())
                    for p in procs:
                        s = p(inp, s)
                    s = s[0].numpy()
                    cases.append((mode, ci, variant, logits, np.packbits(np.isneginf(s)), int(np.argmax(s))))
        rec[f"v{vocab}_mode"] = np.array([c[0] for c in cases])
        rec[f"v{vocab}_hist"] = np.array([c[1] for c in cases])
        rec[f"v{vocab}_hists"] = np.array([",".join(map(str, h)) for h in hists])
        rec[f"v{vocab}_variant"] = np.array([c[2] for c in cases])
        rec[f"v{vocab}_mask"] = np.stack([c[4] for c in cases])
        rec[f"v{vocab}_argmax"] = np.array([c[5] for c in cases])
    np.savez_compressed(os.path.join(OUT, "rules.npz"), **rec)


if __name__ == "__main__":
    if "teacher" in sys.argv[1:]:          # only the fixtures added later (the others are unchanged)
        golden_teacher("tiny")
        golden_teacher("micro128")
    else:
        golden_logmel()
        golden_model("tiny")
        golden_model("micro128")
        golden_rules()
        golden_teacher("tiny")
        golden_teacher("micro128")
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
