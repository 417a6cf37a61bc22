"""BERT WordPiece tokenisation (the `tokenizers` crate call of candle_provider.rs:366-369) against golden
vectors produced by the `tokenizers` package itself (tests/golden/make_tokenizer_golden.py): ids,
type ids and attention masks for plain / truncated / padded configurations, cased and uncased
normalisers, added special tokens, both post-processor spellings.  Host-side code: no GPU."""
import copy
import json
import os

import numpy as np
import pytest

from islands_b200 import InvalidConfig
from islands_b200.tokenizer import BertWordPieceTokenizer

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "tokenizer_golden.json")


def _golden():
    with open(GOLDEN, "r", encoding="utf-8") as f:
        return json.load(f)


def _spec(g, case):
    spec = copy.deepcopy(case["tokenizer_json"])
    spec["model"]["vocab"] = g["vocab"]
    return spec


@pytest.mark.parametrize("case_index", range(10))
def test_matches_tokenizers_crate(case_index):
    g = _golden()
    case = g["cases"][case_index]
    tok = BertWordPieceTokenizer.from_str(json.dumps(_spec(g, case)))
    for run in case["runs"]:
        encs = tok.encode_batch(g["texts"], add_special_tokens=run["add_special_tokens"])
        for i, e in enumerate(encs):
            where = (case["name"], run["add_special_tokens"], i, g["texts"][i][:40])
            assert e.ids == run["ids"][i], where
            n = run["attended"][i]  # the generator checked: masks are prefixes of ones, type ids all zero
            assert e.type_ids == [0] * len(e.ids), where
            assert e.attention_mask == [1] * n + [0] * (len(e.ids) - n), where


def test_embed_texts_raw_padding_layout():
    """candle_provider.rs:372-402: rows are padded with 0 / 0 / 0 to the longest encoding of the batch."""
    g = _golden()
    tok = BertWordPieceTokenizer.from_str(json.dumps(_spec(g, g["cases"][0])))
    ids, types, mask = tok.encode_batch_padded(["hello world", "hello", ""])
    assert ids.dtype == np.int32 and ids.shape == (3, 4) and types.shape == mask.shape == ids.shape
    v = g["vocab"]
    assert ids.tolist() == [[v["[CLS]"], v["hello"], v["world"], v["[SEP]"]], [v["[CLS]"], v["hello"], v["[SEP]"], 0],
                            [v["[CLS]"], v["[SEP]"], 0, 0]]
    assert mask.tolist() == [[1, 1, 1, 1], [1, 1, 1, 0], [1, 1, 0, 0]]
    assert not types.any()
    e = tok.encode_batch_padded([])
    assert e[0].shape == (0, 0)


def test_unsupported_components_fail_loudly(tmp_path):
    g = _golden()
    spec = _spec(g, g["cases"][0])
    for mutate in (lambda s: s["model"].update(type="BPE"), lambda s: s.update(pre_tokenizer={"type": "ByteLevel"}),
                   lambda s: s.update(normalizer={"type": "NFKC"}), lambda s: s.update(post_processor={"type": "RobertaProcessing"}),
                   lambda s: s["model"].update(unk_token="[NOPE]")):
        bad = copy.deepcopy(spec)
        mutate(bad)
        with pytest.raises(InvalidConfig):
            BertWordPieceTokenizer.from_str(json.dumps(bad))
    with pytest.raises(InvalidConfig):
        BertWordPieceTokenizer.from_str("{not json")
    path = tmp_path / "tokenizer.json"
    path.write_text(json.dumps(spec), encoding="utf-8")
    assert BertWordPieceTokenizer.from_file(path).encode("hello").ids == [g["vocab"]["[CLS]"], g["vocab"]["hello"], g["vocab"]["[SEP]"]]
