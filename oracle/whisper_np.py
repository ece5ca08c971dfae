"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement of the Whisper encoder, the KV-cached decoder, the logits rules and the
greedy loop.  Weights are a dict keyed by HF state_dict names (numpy arrays, [out,in] Linear
layout).  Each function cites the arithmetic it follows:
  in-reference spec   ref: training/flax/distil_whisper/modeling_flax_whisper.py
  executable twin     transformers/models/whisper/modeling_whisper.py ("HF:" below)
  rules / greedy      transformers/generation/logits_process.py, generation/utils.py,
                      models/whisper/generation_whisper.py
"""
from __future__ import annotations

import math

import numpy as np

try:  # erf without scipy dependency surprises
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)

NEG_INF = -np.inf


def gelu(x):
    """exact GELU (HF: activation_function="gelu"; ref flax :977,979 approximate=False)."""
    return (0.5 * x * (1.0 + _erf(x / np.sqrt(2.0)))).astype(x.dtype)


def layer_norm(x, w, b, eps=1e-5):
    """biased variance, eps 1e-5 (HF: nn.LayerNorm; ref flax :472)."""
    x64 = x.astype(np.float64)
    mu = x64.mean(-1, keepdims=True)
    var = ((x64 - mu) ** 2).mean(-1, keepdims=True)
    return (((x64 - mu) / np.sqrt(var + eps)) * w + b).astype(x.dtype)


def linear(x, w, b=None):
    y = x @ w.T
    if b is not None:
        y = y + b
    return y


def sinusoids(length: int, channels: int) -> np.ndarray:
    """HF: modeling_whisper.py:55-64 (stored as encoder.embed_positions.weight)."""
    inc = math.log(10000.0) / (channels // 2 - 1)
    inv = np.exp(-inc * np.arange(channels // 2, dtype=np.float64))
    t = np.arange(length, dtype=np.float64)[:, None] * inv[None, :]
    return np.concatenate([np.sin(t), np.cos(t)], axis=1).astype(np.float32)


def conv1d_k3(x, w, b, stride):
    """x [C_in, T], w [C_out, C_in, 3], padding 1 (HF :619-620)."""
    c_in, t = x.shape
    xp = np.pad(x, ((0, 0), (1, 1)))
    t_out = (t + 2 - 3) // stride + 1
    cols = np.stack([xp[:, k:k + stride * t_out:stride] for k in range(3)], axis=1)   # [C_in,3,T_out]
    return (w.reshape(w.shape[0], -1) @ cols.reshape(c_in * 3, t_out)) + b[:, None]


def _split_heads(x, heads):
    t, d = x.shape
    return x.reshape(t, heads, d // heads).transpose(1, 0, 2)          # [H,T,hd]


def attention(q_in, kv_in, W, prefix, heads, causal_offset=None, k_cache=None, v_cache=None):
    """HF: WhisperAttention.forward :284-358 — q scaled by hd^-0.5 before the product, k_proj has
    no bias, softmax in fp32, no extra scaling."""
    d = q_in.shape[-1]
    hd = d // heads
    q = linear(q_in, W[prefix + "q_proj.weight"], W[prefix + "q_proj.bias"]) * np.float32(hd ** -0.5)
    if k_cache is None:
        k = linear(kv_in, W[prefix + "k_proj.weight"])
        v = linear(kv_in, W[prefix + "v_proj.weight"], W[prefix + "v_proj.bias"])
    else:
        k, v = k_cache, v_cache
    qh, kh, vh = _split_heads(q, heads), _split_heads(k, heads), _split_heads(v, heads)
    s = qh @ kh.transpose(0, 2, 1)                                     # [H,Tq,Tk]
    if causal_offset is not None:
        tq, tk = s.shape[1:]
        mask = np.arange(tk)[None, :] > (np.arange(tq)[:, None] + causal_offset)
        s = np.where(mask[None], NEG_INF, s)
    s = s - s.max(-1, keepdims=True)
    p = np.exp(s)
    p = p / p.sum(-1, keepdims=True)
    o = (p.astype(q.dtype) @ vh).transpose(1, 0, 2).reshape(q_in.shape[0], d)
    return linear(o, W[prefix + "out_proj.weight"], W[prefix + "out_proj.bias"]), k, v


def encoder_forward(W, mel, heads, n_layers, taps=None):
    """mel [n_mel,3000] f32 -> enc_out [1500,d].  HF: WhisperEncoder.forward :593-647, layer :380-414.
    `taps` (list) receives the residual stream after the stem and after every layer."""
    p = "model.encoder."
    h = gelu(conv1d_k3(mel, W[p + "conv1.weight"], W[p + "conv1.bias"], 1))
    h = gelu(conv1d_k3(h, W[p + "conv2.weight"], W[p + "conv2.bias"], 2))
    h = h.T + W[p + "embed_positions.weight"]
    if taps is not None:
        taps.append(h.copy())
    for l in range(n_layers):
        lp = f"{p}layers.{l}."
        a = layer_norm(h, W[lp + "self_attn_layer_norm.weight"], W[lp + "self_attn_layer_norm.bias"])
        o, _, _ = attention(a, a, W, lp + "self_attn.", heads)
        h = h + o
        m = layer_norm(h, W[lp + "final_layer_norm.weight"], W[lp + "final_layer_norm.bias"])
        h = h + linear(gelu(linear(m, W[lp + "fc1.weight"], W[lp + "fc1.bias"])), W[lp + "fc2.weight"], W[lp + "fc2.bias"])
        if taps is not None:
            taps.append(h.copy())
    return layer_norm(h, W[p + "layer_norm.weight"], W[p + "layer_norm.bias"])


def cross_kv(W, enc_out, n_layers):
    """computed once per window and cached (HF :314-336)."""
    out = []
    for l in range(n_layers):
        pre = f"model.decoder.layers.{l}.encoder_attn."
        out.append((linear(enc_out, W[pre + "k_proj.weight"]),
                    linear(enc_out, W[pre + "v_proj.weight"], W[pre + "v_proj.bias"])))
    return out


class DecoderState:
    def __init__(self, n_layers):
        self.k = [None] * n_layers
        self.v = [None] * n_layers
        self.pos = 0


def decoder_forward(W, tokens, state: DecoderState, xkv, heads, n_layers, all_positions=False):
    """tokens [Tq] int -> logits of the last position [V] f32 (or of every position [Tq, V] with all_positions: the
    teacher-forced full-sequence pass of the distillation step, ref knowledge-distillation/run_distillation.py:1543-1577).
    HF: WhisperDecoder.forward :691-798, layer :449-507, tied proj_out :1081, logits[:, -1].float()
    (generation/utils.py:2762)."""
    p = "model.decoder."
    tq = len(tokens)
    h = W[p + "embed_tokens.weight"][tokens] + W[p + "embed_positions.weight"][state.pos:state.pos + tq]
    for l in range(n_layers):
        lp = f"{p}layers.{l}."
        a = layer_norm(h, W[lp + "self_attn_layer_norm.weight"], W[lp + "self_attn_layer_norm.bias"])
        k_new = linear(a, W[lp + "self_attn.k_proj.weight"])
        v_new = linear(a, W[lp + "self_attn.v_proj.weight"], W[lp + "self_attn.v_proj.bias"])
        state.k[l] = k_new if state.k[l] is None else np.concatenate([state.k[l], k_new])
        state.v[l] = v_new if state.v[l] is None else np.concatenate([state.v[l], v_new])
        o, _, _ = attention(a, None, W, lp + "self_attn.", heads, causal_offset=state.pos,
                            k_cache=state.k[l], v_cache=state.v[l])
        h = h + o
        c = layer_norm(h, W[lp + "encoder_attn_layer_norm.weight"], W[lp + "encoder_attn_layer_norm.bias"])
        o, _, _ = attention(c, None, W, lp + "encoder_attn.", heads, k_cache=xkv[l][0], v_cache=xkv[l][1])
        h = h + o
        m = layer_norm(h, W[lp + "final_layer_norm.weight"], W[lp + "final_layer_norm.bias"])
        h = h + linear(gelu(linear(m, W[lp + "fc1.weight"], W[lp + "fc1.bias"])), W[lp + "fc2.weight"], W[lp + "fc2.bias"])
    state.pos += tq
    if all_positions:
        h = layer_norm(h, W[p + "layer_norm.weight"], W[p + "layer_norm.bias"])
        return (h @ W[p + "embed_tokens.weight"].T).astype(np.float32)
    h = layer_norm(h[-1:], W[p + "layer_norm.weight"], W[p + "layer_norm.bias"])
    return (h @ W[p + "embed_tokens.weight"].T)[0].astype(np.float32)


def teacher_logits(W, enc_out, decoder_input_ids, heads, n_layers):
    """Logits [T, V] of every position of one row of decoder_input_ids [T] (causal mask, no cache):
    WhisperForConditionalGeneration.forward(...).logits of HF (modeling_whisper.py:1000-1100)."""
    xkv = cross_kv(W, enc_out, n_layers)
    return decoder_forward(W, np.asarray(decoder_input_ids), DecoderState(n_layers), xkv, heads, n_layers, all_positions=True)


def shift_tokens_right(labels, pad_token_id, decoder_start_token_id):
    """HF shift_tokens_right (modeling_whisper.py:67-81)."""
    labels = np.asarray(labels)
    out = np.zeros_like(labels)
    out[:, 1:] = labels[:, :-1]
    out[:, 0] = decoder_start_token_id
    out[out == -100] = pad_token_id
    return out


def log_softmax(x):
    m = x.max()
    if not np.isfinite(m):
        return x - m
    z = x - m
    return z - np.log(np.exp(z).sum())


def apply_rules(logits, generated, rules):
    """generated: tokens sampled so far (after the forced prompt).  Order = HF
    _retrieve_logit_processors (generation_whisper.py:1774-1812): begin-suppress, suppress, timestamps.
    rules: dict(suppress, begin_suppress, eos, ts_begin (or None), no_timestamps, max_initial_ts)."""
    s = logits.astype(np.float32).copy()
    if len(generated) == 0 and rules.get("begin_suppress"):
        s[np.asarray(rules["begin_suppress"])] = NEG_INF            # logits_process.py:1855-1863
    if rules.get("suppress"):
        s[np.asarray(rules["suppress"])] = NEG_INF                  # :1896-1903
    tsb = rules.get("ts_begin")
    if tsb is not None:                                             # :1996-2044
        eos = rules["eos"]
        s[rules["no_timestamps"]] = NEG_INF
        seq = list(generated)
        last_ts = len(seq) >= 1 and seq[-1] >= tsb
        pen_ts = len(seq) < 2 or seq[-2] >= tsb
        if last_ts:
            if pen_ts:
                s[tsb:] = NEG_INF
            else:
                s[:eos] = NEG_INF
        ts = [t for t in seq if t >= tsb]
        if ts:
            last = ts[-1] if (last_ts and not pen_ts) else ts[-1] + 1
            s[tsb:last] = NEG_INF
        if len(seq) == 0:
            s[:tsb] = NEG_INF
            if rules.get("max_initial_ts") is not None:
                s[tsb + rules["max_initial_ts"] + 1:] = NEG_INF
        lp = log_softmax(s.astype(np.float32))
        ts_lp = lp[tsb:]
        mts = ts_lp.max()
        ts_logprob = mts + np.log(np.exp(ts_lp - mts).sum()) if np.isfinite(mts) else -np.inf
        if ts_logprob > lp[:tsb].max():
            s[:tsb] = NEG_INF
    return s


def greedy_decode(W, enc_out, prompt, rules, max_length, heads, n_layers, forced=None, logits_tap=None):
    """Greedy loop of generation/utils.py:2658-2805 for one row: stop at eos or when the total
    length (prompt included) reaches max_length.  Returns generated tokens *excluding* eos
    (generation_whisper.py:1063-1086 strips it).  `forced`: teacher-forcing tokens fed back
    instead of the argmax (diagnostic)."""
    xkv = cross_kv(W, enc_out, n_layers)
    st = DecoderState(n_layers)
    out = []
    fed = []
    cur = list(prompt)
    while len(prompt) + len(out) < max_length:
        logits = decoder_forward(W, np.asarray(cur), st, xkv, heads, n_layers)
        s = apply_rules(logits, fed, rules)
        if logits_tap is not None:
            logits_tap.append(s.copy())
        tok = int(np.argmax(s))
        out.append(tok)
        nxt = tok if forced is None else int(forced[len(out) - 1])
        if nxt == rules["eos"]:
            if tok == rules["eos"]:
                out.pop()
            break
        fed.append(nxt)
        cur = [nxt]
    return out
