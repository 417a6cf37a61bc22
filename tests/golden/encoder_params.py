"""Deterministic BERT weights shared by tests/golden/make_encoder_golden.py (which loads them into Hugging Face
transformers' BertModel) and the tests (which load them into the oracle and the CUDA encoder): numpy
RandomState streams keyed by parameter name, so the golden file only has to hold inputs and outputs."""
import zlib

import numpy as np

SHAPE = dict(vocab_size=120, hidden_size=128, num_layers=2, num_heads=2, intermediate_size=256, max_position=24)


def parameter_shapes(shape=SHAPE):
    H, I = shape["hidden_size"], shape["intermediate_size"]
    out = {"embeddings.word_embeddings.weight": (shape["vocab_size"], H),
           "embeddings.position_embeddings.weight": (shape["max_position"], H),
           "embeddings.token_type_embeddings.weight": (2, H),
           "embeddings.LayerNorm.weight": (H,), "embeddings.LayerNorm.bias": (H,)}
    for l in range(shape["num_layers"]):
        p = f"encoder.layer.{l}."
        for n in ("query", "key", "value"):
            out[p + f"attention.self.{n}.weight"] = (H, H)
            out[p + f"attention.self.{n}.bias"] = (H,)
        out[p + "attention.output.dense.weight"] = (H, H)
        out[p + "attention.output.dense.bias"] = (H,)
        out[p + "attention.output.LayerNorm.weight"] = (H,)
        out[p + "attention.output.LayerNorm.bias"] = (H,)
        out[p + "intermediate.dense.weight"] = (I, H)
        out[p + "intermediate.dense.bias"] = (I,)
        out[p + "output.dense.weight"] = (H, I)
        out[p + "output.dense.bias"] = (H,)
        out[p + "output.LayerNorm.weight"] = (H,)
        out[p + "output.LayerNorm.bias"] = (H,)
    return out


def make_params(shape=SHAPE, seed=0):
    params = {}
    for name, shp in parameter_shapes(shape).items():
        rng = np.random.RandomState((zlib.crc32(name.encode()) + seed) & 0x7fffffff)
        if name.endswith("LayerNorm.weight"):
            a = 1.0 + 0.1 * rng.randn(*shp)
        elif name.endswith(".bias"):
            a = 0.1 * rng.randn(*shp)
        else:
            a = 0.08 * rng.randn(*shp)  # wide enough that attention and GELU are exercised
        params[name] = a.astype(np.float32)
    return params
