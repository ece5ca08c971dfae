"""TEST INFRASTRUCTURE ONLY — the parity checker for the B200 Whisper teacher-inference path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it, and only as the checker
or as the timed CPU reference — never as (part of) the thing measured or shipped.

Contents
  logmel_np.py   numpy restatement of HF WhisperFeatureExtractor (the arithmetic the reference
                 calls at ref: training/run_pseudo_labelling.py:739 and
                 prefiltering/validator_inference.py:57-60; in-tree twin
                 ref: training/flax/distil_whisper/pipeline.py:40-58)
  whisper_np.py  numpy restatement of the Whisper encoder / KV-cached greedy decoder and the
                 logits rules (in-tree spec ref: training/flax/distil_whisper/modeling_flax_whisper.py;
                 executable twin: transformers/models/whisper/modeling_whisper.py)
  hf_ref.py      builds the reference's *own* implementation (HuggingFace transformers, the
                 third-party dependency the reference pins at ref: environment.yml:175) with
                 random-init weights and an offline generation config, and runs it on CPU

Pinning: the reference ships no tests, golden vectors or fixtures for this path,
so the restatement is pinned against outputs of the importable HF implementation itself:
``tests/golden/make_golden.py`` generated the committed fixtures in ``tests/golden/`` and
``tests/test_oracle.py`` checks the restatement against them (and live against HF, which is in
the image on both boxes).  Version skew: the image has transformers 5.5.0, the reference pins
4.45.2.
"""
