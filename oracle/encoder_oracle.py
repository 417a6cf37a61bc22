"""CPU oracle of the recompute encoder — TEST INFRASTRUCTURE ONLY (imported by tests/, smoke() and
bench.py's checker legs; never by the product).

Restates the reference's CandleEmbedder::embed_texts_raw (src/core/embedding/candle_provider.rs:
353-507) for already-tokenised input in fp32 numpy: BERT forward -> masked mean pooling with
clamp(sum_mask, 1e-9) (:438-474) -> L2 normalisation with clamp(norm, 1e-12) (:477-494).  The BERT
forward is third-party in the reference (candle-transformers 0.9.1 `bert`, Cargo.lock:1113; not
vendored): this follows the published architecture (post-LayerNorm encoder, erf GELU, additive
attention mask).  PINNING: the reference's own encoder tests check only config / preset strings
(candle_provider.rs:514-592), so it holds no golden embedding; this restatement is pinned against
Hugging Face transformers' BertModel — the implementation candle's bert.rs mirrors — plus the reference's
pooling recipe, through tests/golden/encoder_golden.npz (tests/golden/make_encoder_golden.py; agreement
2e-6 on unit vectors).  Not pinned against candle itself: no Rust toolchain here.
"""
import numpy as np
from scipy.special import erf


def _ln(x, w, b, eps):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * w + b


def bert_embed(params, cfg, token_ids, lengths):
    """params: dict of Hugging Face BERT names -> f32 arrays; cfg: object with hidden_size, num_layers,
    num_heads, layer_norm_eps, normalize.  token_ids [B,S] int, lengths [B] -> [B,H] f32."""
    t = np.asarray(token_ids)
    B, S = t.shape
    H, nh, eps = cfg.hidden_size, cfg.num_heads, np.float32(cfg.layer_norm_eps)
    hd = H // nh
    mask = (np.arange(S)[None, :] < np.asarray(lengths)[:, None]).astype(np.float32)  # [B,S]
    x = (params["embeddings.word_embeddings.weight"][t] + params["embeddings.position_embeddings.weight"][None, :S]
         + params["embeddings.token_type_embeddings.weight"][0][None, None])
    x = _ln(x.astype(np.float32), params["embeddings.LayerNorm.weight"], params["embeddings.LayerNorm.bias"], eps)
    neg = np.where(mask[:, None, None, :] > 0, np.float32(0), np.float32(-np.inf))  # additive mask over keys
    for l in range(cfg.num_layers):
        p = f"encoder.layer.{l}."
        g = lambda n: params[p + n]
        q = x @ g("attention.self.query.weight").T + g("attention.self.query.bias")
        k = x @ g("attention.self.key.weight").T + g("attention.self.key.bias")
        v = x @ g("attention.self.value.weight").T + g("attention.self.value.bias")
        sp = lambda a: a.reshape(B, S, nh, hd).transpose(0, 2, 1, 3)
        sc = sp(q) @ sp(k).transpose(0, 1, 3, 2) / np.sqrt(np.float32(hd)) + neg
        with np.errstate(invalid="ignore"):
            sc = sc - sc.max(-1, keepdims=True)
            pr = np.exp(sc)
            pr = pr / pr.sum(-1, keepdims=True)
        pr = np.nan_to_num(pr)  # rows of a fully masked sequence
        ctx = (pr @ sp(v)).transpose(0, 2, 1, 3).reshape(B, S, H)
        y = ctx @ g("attention.output.dense.weight").T + g("attention.output.dense.bias") + x
        x = _ln(y, g("attention.output.LayerNorm.weight"), g("attention.output.LayerNorm.bias"), eps)
        h = x @ g("intermediate.dense.weight").T + g("intermediate.dense.bias")
        h = 0.5 * h * (1.0 + erf(h / np.sqrt(2.0)))
        z = h.astype(np.float32) @ g("output.dense.weight").T + g("output.dense.bias") + x
        x = _ln(z, g("output.LayerNorm.weight"), g("output.LayerNorm.bias"), eps)
    s = (x * mask[:, :, None]).sum(1)
    den = np.clip(mask.sum(1, keepdims=True), 1e-9, None)
    m = (s / den).astype(np.float32)
    if cfg.normalize:
        m = m / np.clip(np.sqrt((m * m).sum(1, keepdims=True)), 1e-12, None)
    return m.astype(np.float32)
