"""Generates tests/golden/encoder_golden.npz with Hugging Face `transformers`' BertModel (torch, CPU, fp32):
the published BERT implementation that candle-transformers' `bert.rs` (the reference's model forward,
Cargo.lock candle-transformers 0.9.1) mirrors, followed by the reference's own pooling recipe
(src/core/embedding/candle_provider.rs:438-494: mean over the sequence weighted by the attention mask with
clamp(sum mask, 1e-9), then L2 normalisation with clamp(norm, 1e-12)).  The weights are the deterministic
streams of tests/golden/encoder_params.py, so only inputs and outputs are stored.  Pins oracle/encoder_oracle.py
(tests/test_encoder.py::test_encoder_oracle_vs_transformers_golden) and, through it, the CUDA encoder.
Run here:  python tests/golden/make_encoder_golden.py"""
import os
import sys

import numpy as np
import torch
from transformers import BertConfig, BertModel

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from encoder_params import SHAPE, make_params  # noqa: E402


def main():
    cfg = BertConfig(vocab_size=SHAPE["vocab_size"], hidden_size=SHAPE["hidden_size"], num_hidden_layers=SHAPE["num_layers"],
                     num_attention_heads=SHAPE["num_heads"], intermediate_size=SHAPE["intermediate_size"],
                     max_position_embeddings=SHAPE["max_position"], type_vocab_size=2, hidden_act="gelu", hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0, layer_norm_eps=1e-12)
    model = BertModel(cfg, add_pooling_layer=False).eval()
    params = make_params(SHAPE, seed=0)
    sd = model.state_dict()
    assert set(params) == {k for k in sd if not k.endswith("position_ids")}, "parameter names differ from transformers'"
    model.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=False)
    rng = np.random.RandomState(3)
    B, S = 24, 20
    assert S <= SHAPE["max_position"]
    tok = rng.randint(1, SHAPE["vocab_size"], size=(B, S)).astype(np.int64)
    ln = rng.randint(1, S + 1, size=B).astype(np.int64)
    ln[0], ln[1] = S, 1
    mask = (np.arange(S)[None, :] < ln[:, None]).astype(np.int64)
    tok = tok * mask  # padding id 0, as embed_texts_raw pads (candle_provider.rs:391-401)
    with torch.no_grad():
        h = model(input_ids=torch.from_numpy(tok), token_type_ids=torch.zeros((B, S), dtype=torch.long),
                  attention_mask=torch.from_numpy(mask)).last_hidden_state
        m = torch.from_numpy(mask).to(torch.float32)
        summed = (h * m[:, :, None]).sum(1)                      # candle_provider.rs:438-466
        mean = summed / m.sum(1, keepdim=True).clamp(min=1e-9)   # :467-474
        emb = mean / mean.norm(dim=1, keepdim=True).clamp(min=1e-12)  # :477-494
    out = {}
    out.update(token_ids=tok.astype(np.int32), lengths=ln.astype(np.int32), pooled=mean.numpy().astype(np.float32),
               embeddings=emb.numpy().astype(np.float32))
    path = os.path.join(HERE, "encoder_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; transformers", __import__("transformers").__version__)


if __name__ == "__main__":
    main()
