"""taiwan-whisper_b200 — B200-native (sm_100a) batched Whisper teacher inference.

Only the hot path of forbes110/taiwan-whisper lives here: log-mel front end, Whisper encoder and
KV-cached greedy decoding behind the reference's own call sites
(`feature_extractor(...)`, `WhisperForConditionalGeneration.generate(...)`).
Heavy imports (torch, the CUDA library) happen lazily in the submodules.
"""
__version__ = "0.1.0"
