"""BERT WordPiece tokenisation for the recompute encoder's text entry point
(`CandleEmbedder::embed_texts_raw`, src/core/embedding/candle_provider.rs:353-437).

The reference loads a Hugging Face `tokenizer.json` with the `tokenizers` crate (0.22, Cargo.lock)
and calls `encode_batch(texts, true)`; that crate is third-party and not part of /root/reference, so
this module restates its published pipeline for the BERT family of tokenizer.json files:

    added tokens -> BertNormalizer -> BertPreTokenizer -> WordPiece -> TemplateProcessing / BertProcessing
    -> truncation -> padding

and is pinned by golden vectors produced with the `tokenizers` package itself (same Rust core) in
tests/golden/make_tokenizer_golden.py.  Host-side code: token ids are the only thing that goes to the
device (`Encoder.embed`).  Components outside the BERT family raise `InvalidConfig` instead of guessing.
"""
import json
import unicodedata
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .core import InvalidConfig


@dataclass
class Encoding:
    """The three arrays `embed_texts_raw` reads from a `tokenizers::Encoding` (candle_provider.rs:385-389)."""
    ids: List[int]
    type_ids: List[int]
    attention_mask: List[int]
    tokens: List[str]

    def get_ids(self):
        return self.ids

    def get_type_ids(self):
        return self.type_ids

    def get_attention_mask(self):
        return self.attention_mask


def _is_whitespace(c: str) -> bool:
    # normalizers/bert.rs: '\t' | '\n' | '\r' or char::is_whitespace (Unicode White_Space)
    if c in "\t\n\r":
        return True
    return c in _WHITE_SPACE


# Unicode White_Space property (what Rust's char::is_whitespace tests)
_WHITE_SPACE = frozenset(
    [chr(c) for c in (0x09, 0x0A, 0x0B, 0x0C, 0x0D, 0x20, 0x85, 0xA0, 0x1680, 0x2028, 0x2029, 0x202F, 0x205F, 0x3000)] +
    [chr(c) for c in range(0x2000, 0x200B)])


def _is_control(c: str) -> bool:
    # normalizers/bert.rs: tab / newline / carriage return are not control; otherwise category "Other"
    if c in "\t\n\r":
        return False
    return unicodedata.category(c).startswith("C")


def _is_chinese_char(cp: int) -> bool:
    # normalizers/bert.rs is_chinese_char
    return (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2A6DF or 0x2A700 <= cp <= 0x2B73F or
            0x2B740 <= cp <= 0x2B81F or 0x2B920 <= cp <= 0x2CEAF or 0xF900 <= cp <= 0xFAFF or 0x2F800 <= cp <= 0x2FA1F)


def _is_bert_punc(c: str) -> bool:
    # pre_tokenizers/bert.rs: char::is_ascii_punctuation || unicode category P*
    cp = ord(c)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
        return True
    return unicodedata.category(c).startswith("P")


class BertWordPieceTokenizer:
    """`tokenizers::Tokenizer` for BERT-family tokenizer.json files.  `from_file` / `from_str` mirror
    `Tokenizer::from_file` (candle_provider.rs:274); `encode_batch(texts, add_special_tokens)` mirrors the
    call at candle_provider.rs:366-369."""

    def __init__(self, spec: dict):
        model = spec.get("model") or {}
        if model.get("type", "WordPiece") != "WordPiece" or "vocab" not in model:
            raise InvalidConfig("tokenizer.json: only the WordPiece model is supported")
        self.vocab = dict(model["vocab"])
        self.unk_token = model.get("unk_token", "[UNK]")
        self.prefix = model.get("continuing_subword_prefix", "##")
        self.max_chars = int(model.get("max_input_chars_per_word", 100))
        if self.unk_token not in self.vocab:
            raise InvalidConfig("tokenizer.json: unk_token is not in the vocabulary")
        self.id_to_token = {i: t for t, i in self.vocab.items()}

        norm = spec.get("normalizer")
        if norm is None:
            self.norm = None
        elif norm.get("type") == "BertNormalizer":
            lowercase = bool(norm.get("lowercase", True))
            strip = norm.get("strip_accents")
            self.norm = dict(clean_text=bool(norm.get("clean_text", True)),
                             handle_chinese_chars=bool(norm.get("handle_chinese_chars", True)),
                             strip_accents=lowercase if strip is None else bool(strip), lowercase=lowercase)
        else:
            raise InvalidConfig(f"tokenizer.json: normalizer {norm.get('type')} is not supported")
        pre = spec.get("pre_tokenizer")
        if pre is None or pre.get("type") != "BertPreTokenizer":
            raise InvalidConfig("tokenizer.json: only BertPreTokenizer is supported")

        self.added = []
        for t in spec.get("added_tokens") or []:
            if t.get("single_word") or t.get("lstrip") or t.get("rstrip") or t.get("normalized"):
                if t.get("special"):
                    raise InvalidConfig("tokenizer.json: added tokens with single_word / lstrip / rstrip / normalized are not supported")
                continue
            self.added.append((t["content"], int(t["id"])))
            self.id_to_token.setdefault(int(t["id"]), t["content"])
        self.added.sort(key=lambda t: -len(t[0]))

        self.single = [("seq", "A", 0)]  # no post-processor: the sequence alone
        post = spec.get("post_processor")
        if post is not None:
            if post.get("type") == "TemplateProcessing":
                self.single = []
                for piece in post["single"]:
                    if "SpecialToken" in piece:
                        sp = post["special_tokens"][piece["SpecialToken"]["id"]]
                        for tid, tok in zip(sp["ids"], sp["tokens"]):
                            self.single.append(("special", (int(tid), tok), int(piece["SpecialToken"]["type_id"])))
                    else:
                        if piece["Sequence"]["id"] != "A":
                            raise InvalidConfig("tokenizer.json: single template refers to sequence B")
                        self.single.append(("seq", "A", int(piece["Sequence"]["type_id"])))
            elif post.get("type") == "BertProcessing":
                sep, cls = post["sep"], post["cls"]
                self.single = [("special", (int(cls[1]), cls[0]), 0), ("seq", "A", 0), ("special", (int(sep[1]), sep[0]), 0)]
            else:
                raise InvalidConfig(f"tokenizer.json: post_processor {post.get('type')} is not supported")
        self.n_added_single = sum(1 for p in self.single if p[0] == "special")

        self.truncation = None
        tr = spec.get("truncation")
        if tr is not None:
            if tr.get("direction", "Right") != "Right" or int(tr.get("stride", 0)) != 0:
                raise InvalidConfig("tokenizer.json: only right truncation with stride 0 is supported")
            self.truncation = int(tr["max_length"])
        self.padding = None
        pad = spec.get("padding")
        if pad is not None:
            if pad.get("direction", "Right") != "Right":
                raise InvalidConfig("tokenizer.json: only right padding is supported")
            strat = pad.get("strategy", "BatchLongest")
            fixed = None if strat == "BatchLongest" else int(strat["Fixed"])
            self.padding = dict(fixed=fixed, multiple=pad.get("pad_to_multiple_of"), pad_id=int(pad.get("pad_id", 0)),
                                pad_type_id=int(pad.get("pad_type_id", 0)), pad_token=pad.get("pad_token", "[PAD]"))

    # -- construction ---------------------------------------------------------------------------
    @classmethod
    def from_str(cls, text: str) -> "BertWordPieceTokenizer":
        try:
            return cls(json.loads(text))
        except (ValueError, KeyError, TypeError) as e:
            raise InvalidConfig(f"tokenizer.json: {e}") from None

    @classmethod
    def from_file(cls, path) -> "BertWordPieceTokenizer":
        with open(path, "r", encoding="utf-8") as f:
            return cls.from_str(f.read())

    # -- pipeline stages ------------------------------------------------------------------------
    def _normalize(self, text: str) -> str:
        n = self.norm
        if n is None:
            return text
        if n["clean_text"]:
            text = "".join(" " if _is_whitespace(c) else c for c in text
                           if not (ord(c) == 0 or ord(c) == 0xFFFD or _is_control(c)))
        if n["handle_chinese_chars"]:
            text = "".join(f" {c} " if _is_chinese_char(ord(c)) else c for c in text)
        if n["strip_accents"]:
            text = "".join(c for c in unicodedata.normalize("NFD", text) if unicodedata.category(c) != "Mn")
        if n["lowercase"]:
            text = text.lower()
        return text

    @staticmethod
    def _pre_tokenize(text: str) -> List[str]:
        words, cur = [], []
        for c in text:
            if c in _WHITE_SPACE:  # split on whitespace, removed
                if cur:
                    words.append("".join(cur))
                    cur = []
            elif _is_bert_punc(c):  # punctuation is isolated
                if cur:
                    words.append("".join(cur))
                    cur = []
                words.append(c)
            else:
                cur.append(c)
        if cur:
            words.append("".join(cur))
        return words

    def _wordpiece(self, word: str) -> List[int]:
        # models/wordpiece: greedy longest match first; a word that cannot be covered is one [UNK]
        if len(word) > self.max_chars:
            return [self.vocab[self.unk_token]]
        out, start = [], 0
        while start < len(word):
            end, found = len(word), None
            while start < end:
                sub = word[start:end]
                if start > 0:
                    sub = self.prefix + sub
                if sub in self.vocab:
                    found = self.vocab[sub]
                    break
                end -= 1
            if found is None:
                return [self.vocab[self.unk_token]]
            out.append(found)
            start = end
        return out

    def _split_added(self, text: str):
        """Leftmost-longest extraction of added (special) tokens from the raw text."""
        if not self.added:
            return [(text, None)]
        parts, i, last = [], 0, 0
        while i < len(text):
            hit = None
            for content, tid in self.added:
                if text.startswith(content, i):
                    hit = (content, tid)
                    break
            if hit:
                if last < i:
                    parts.append((text[last:i], None))
                parts.append((hit[0], hit[1]))
                i += len(hit[0])
                last = i
            else:
                i += 1
        if last < len(text):
            parts.append((text[last:], None))
        return parts

    def _encode_sequence(self, text: str) -> List[int]:
        ids = []
        for piece, tid in self._split_added(text):
            if tid is not None:
                ids.append(tid)
                continue
            for w in self._pre_tokenize(self._normalize(piece)):
                ids.extend(self._wordpiece(w))
        return ids

    # -- public API -----------------------------------------------------------------------------
    def encode(self, text: str, add_special_tokens: bool = True) -> Encoding:
        return self.encode_batch([text], add_special_tokens)[0]

    def encode_batch(self, texts: Sequence[str], add_special_tokens: bool = True) -> List[Encoding]:
        encs = []
        for text in texts:
            seq = self._encode_sequence(text)
            if self.truncation is not None:
                room = self.truncation - (self.n_added_single if add_special_tokens else 0)
                seq = seq[:max(room, 0)]
            ids, types = [], []
            if add_special_tokens:
                for kind, val, type_id in self.single:
                    if kind == "special":
                        ids.append(val[0])
                        types.append(type_id)
                    else:
                        ids.extend(seq)
                        types.extend([type_id] * len(seq))
            else:
                ids, types = list(seq), [0] * len(seq)
            encs.append(Encoding(ids, types, [1] * len(ids), []))
        if self.padding is not None and encs:
            p = self.padding
            target = p["fixed"] if p["fixed"] is not None else max(len(e.ids) for e in encs)
            if p["multiple"]:
                m = int(p["multiple"])
                if target % m:
                    target += m - target % m
            for e in encs:
                short = target - len(e.ids)
                if short > 0:
                    e.ids += [p["pad_id"]] * short
                    e.type_ids += [p["pad_type_id"]] * short
                    e.attention_mask += [0] * short
        for e in encs:
            e.tokens = [self.id_to_token.get(i, self.unk_token) for i in e.ids]
        return encs

    def encode_batch_padded(self, texts: Sequence[str]):
        """The tensors `embed_texts_raw` builds (candle_provider.rs:372-402): every row padded with zeros
        to the longest encoding of the batch.  Returns (input_ids, token_type_ids, attention_mask) as
        int32 [B][max_len] plus the per-row count of attended tokens."""
        encs = self.encode_batch(texts, True)
        max_len = max((len(e.ids) for e in encs), default=0)
        b = len(encs)
        ids = np.zeros((b, max_len), np.int32)
        types = np.zeros((b, max_len), np.int32)
        mask = np.zeros((b, max_len), np.int32)
        for r, e in enumerate(encs):
            ids[r, :len(e.ids)] = e.ids
            types[r, :len(e.ids)] = e.type_ids
            mask[r, :len(e.ids)] = e.attention_mask
        return ids, types, mask
