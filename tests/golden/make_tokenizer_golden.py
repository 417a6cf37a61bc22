"""Generates tests/golden/tokenizer_golden.json with the `tokenizers` package (0.22.x: the Python
binding of the same Rust crate the reference links, Cargo.lock `tokenizers`), i.e. with the reference's
own tokenisation code.  The vocabulary is synthetic (no network): specials, ASCII letters / digits /
punctuation as whole and continuation pieces, a few hundred frequent English and code words and
subwords, accented / CJK / Cyrillic samples.  Run here (the package is not needed at test time):
    python tests/golden/make_tokenizer_golden.py
"""
import json
import os
import random

from tokenizers import Tokenizer, normalizers, pre_tokenizers, processors
from tokenizers.models import WordPiece

HERE = os.path.dirname(os.path.abspath(__file__))

WORDS = """the of and to in is that for it as was with be by on not he this are or his from at which but have an had they you were
their one all we can her has there been if more when will would who so no out up into than them only its time some could these two may
then do first any my now such like our over man me even most made after also did many before must through back years where much your way
well down should because each just those people how too little state good very make world still own see men work long get here between
both life being under never day same another know while last might us great old year off come since against go came right used take three
fn let mut pub struct impl trait enum match return self use mod crate async await vec string option result some none ok err index search
graph node vector query distance cosine embed embedding token tokenizer model batch layer hidden attention mask pool norm weight bias
function class def import from print value key list dict int float bool true false null void static const for while if else elif
hello world un aff able ing ed er est ly tion ment ness ous ive s es re pre dis mis non anti de over under""".split()
SUFFIXES = ["s", "es", "ed", "ing", "er", "est", "ly", "tion", "ment", "ness", "able", "aff", "ous", "ive", "al", "ity", "ize", "ors", "or"]
EXTRA = ["café", "cafe", "naive", "uber", "strasse", "елка", "привет", "мир", "世", "界", "你", "好", "日", "本", "語", "ß", "ø", "æ", "œ", "i̇"]


def build_vocab():
    toks = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    chars = [chr(c) for c in range(33, 127)]
    toks += chars + ["##" + c for c in chars if c.isalnum()]
    toks += [w.lower() for w in WORDS] + ["##" + s for s in SUFFIXES] + EXTRA + ["##" + e for e in ("é", "е", "и", "界")]
    seen, vocab = set(), {}
    for t in toks:
        if t not in seen:
            seen.add(t)
            vocab[t] = len(vocab)
    return vocab


TEXTS = [
    "Hello, world!", "hello", "", "   ", "The quick brown fox jumps over the lazy dog.",
    "unaffable searching indexes; embeddings' normalization?", "fn main() { let mut v: Vec<u32> = Vec::new(); v.push(42); }",
    "Café CAFÉ naïve Über straße ÀÉÎÕÜ ñ ç", "İstanbul ǅ ẞ ﬁ ﬂ", "你好世界 日本語のテキスト mixed 世界hello界",
    "Привет, мир! ёлка Ёлка", "tabs\tand\nnewlines\r\nand\x00nul�repl​zw nbsp em　ideographic",
    "[CLS] literal [SEP] specials [MASK] inside [PAD] text [UNK]", "a" * 101 + " " + "b" * 100, "x" * 99 + "é",
    "é combining vs é precomposed; ȫ stacked", "math: 3.14159 * 2 = 6.28318; 1e-5 <= x && y >= 10_000",
    "snake_case camelCase PascalCase kebab-case SCREAMING_CASE", "emoji 😀 and symbols ∑ ∞ € £ ¥ © ® ™ ° ± × ÷ — – … « » “ ” ‘ ’ ¿ ¡",
    "url https://example.com/path?q=1&r=2#frag and email a.b@c.de", "١٢٣ arabic digits שלום hebrew مرحبا",
    "control \x01\x02\x7f\x80\x9f chars ‮ rtl ﻿ bom ­ shy", "thai สวัสดี devanagari नमस्ते korean 안녕하세요",
]


def make(vocab, truncation=None, padding=None, bert_processing=False, lowercase=True, strip_accents=None, with_added=True):
    tok = Tokenizer(WordPiece(vocab, unk_token="[UNK]", max_input_chars_per_word=100))
    tok.normalizer = normalizers.BertNormalizer(clean_text=True, handle_chinese_chars=True, strip_accents=strip_accents, lowercase=lowercase)
    tok.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
    if bert_processing:
        tok.post_processor = processors.BertProcessing(("[SEP]", vocab["[SEP]"]), ("[CLS]", vocab["[CLS]"]))
    else:
        tok.post_processor = processors.TemplateProcessing(
            single="[CLS] $A [SEP]", pair="[CLS] $A [SEP] $B:1 [SEP]:1",
            special_tokens=[("[CLS]", vocab["[CLS]"]), ("[SEP]", vocab["[SEP]"])])
    if with_added:
        tok.add_special_tokens(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"])
    if truncation:
        tok.enable_truncation(max_length=truncation)
    if padding == "longest":
        tok.enable_padding(pad_id=vocab["[PAD]"], pad_token="[PAD]")
    elif padding:
        tok.enable_padding(pad_id=vocab["[PAD]"], pad_token="[PAD]", length=padding[0], pad_to_multiple_of=padding[1])
    return tok


def main():
    vocab = build_vocab()
    rng = random.Random(5)
    pool = list(vocab.keys())[5:] + [" ", " ", " ", "  ", "Zq", "é", "世"]
    fuzz = ["".join(rng.choice(pool).replace("##", "") + rng.choice(["", " ", " ", ""]) for _ in range(rng.randint(1, 40))) for _ in range(60)]
    texts = TEXTS + fuzz
    cases = []
    for name, kw in [("plain", {}), ("trunc16", dict(truncation=16)), ("pad_longest", dict(padding="longest")),
                     ("trunc12_fixed24", dict(truncation=12, padding=(24, None))), ("pad_multiple8", dict(padding=(None, 8))),
                     ("bert_processing", dict(bert_processing=True)), ("cased", dict(lowercase=False)),
                     ("cased_strip", dict(lowercase=False, strip_accents=True)), ("lower_keep_accents", dict(strip_accents=False)),
                     ("no_added_tokens", dict(with_added=False))]:
        tok = make(vocab, **kw)
        out = []
        for special in (True, False):
            encs = tok.encode_batch(texts, add_special_tokens=special)
            for e in encs:  # right padding only and a single sequence: the masks are prefixes, the type ids zero
                n = sum(e.attention_mask)
                assert e.attention_mask == [1] * n + [0] * (len(e.ids) - n) and not any(e.type_ids)
            out.append(dict(add_special_tokens=special, ids=[e.ids for e in encs], attended=[sum(e.attention_mask) for e in encs]))
        spec = json.loads(tok.to_str())
        assert spec["model"].pop("vocab") == vocab  # stored once below
        cases.append(dict(name=name, tokenizer_json=spec, runs=out))
    with open(os.path.join(HERE, "tokenizer_golden.json"), "w", encoding="utf-8") as f:
        json.dump(dict(texts=texts, vocab=vocab, cases=cases), f, ensure_ascii=True, separators=(",", ":"))
    print("texts", len(texts), "cases", len(cases), "bytes", os.path.getsize(os.path.join(HERE, "tokenizer_golden.json")))


if __name__ == "__main__":
    main()
