"""Search with on-demand recompute (SURVEY §8 a20 + the EmbeddingProvider seam, leann.rs:82-99, :947-950):
ADC traversal -> bf16 tcgen05 encoder over the distinct survivors -> exact rerank.  The index is
built over the fp32 ORACLE embeddings of a token table; the recompute search never reads them.
Bars (north_star): traversal counters identical to the stored-vector search (same kernel, same
inputs), recall@10 within 0.002 of the fp32 stored-vector search, distances within bf16 tolerance."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _token_table(rng, n, nq, S, vocab, clusters):
    base = rng.randint(1, vocab, size=(clusters, S))

    def draw(count):
        t = base[rng.randint(0, clusters, size=count)].copy()
        flips = rng.rand(count, S) < 0.15
        t[flips] = rng.randint(1, vocab, size=int(flips.sum()))
        ln = rng.randint(S // 2, S + 1, size=count)
        for i in range(count):
            t[i, ln[i]:] = 0
        return t.astype(np.int32), ln.astype(np.int32)

    return draw(n), draw(nq)


def test_recompute_search_matches_stored_vector_search(gpu_lib):
    from islands_b200 import (Encoder, EncoderConfig, LeannConfig, LeannIndex, PQConfig, ProductQuantizer)
    from oracle.encoder_oracle import bert_embed

    rng = np.random.RandomState(11)
    n, nq, S, k, ef = 3000, 200, 16, 10, 64
    cfg_e = EncoderConfig(vocab_size=2000, hidden_size=128, num_layers=2, num_heads=2, intermediate_size=512, max_position=32)
    enc = Encoder(cfg_e).init_random(seed=3, stddev=0.08)
    (tok, ln), (qtok, qln) = _token_table(rng, n, nq, S, 2000, 150)
    sd = enc.state_dict()
    vectors = bert_embed(sd, cfg_e, tok, ln)          # fp32 oracle embeddings: what the index is built over
    queries = bert_embed(sd, cfg_e, qtok, qln)
    cfg = LeannConfig(m=12, m0=24, ef_construction=64)
    index = LeannIndex(cfg)
    index.build(vectors, n, seed=5, batch=64)
    pq = ProductQuantizer(128, PQConfig(16, 64, 10, 1))
    pq.train(vectors)
    index.attach_pq(pq, pq.encode(vectors))

    ids_a, dist_a, cnt_a, st_a = index.search_adc_rerank_batch(queries, k, ef, stats=True)   # stored fp32 vectors
    index.set_recompute(enc, tok, ln)
    ids_b, dist_b, cnt_b, st_b = index.search_adc_recompute_batch(queries, k, ef, stats=True)  # bf16 recompute
    info = index.last_recompute()
    assert 0 < info["unique_nodes"] <= min(n, nq * ef)
    assert info["encoder_ms"] > 0 and info["traverse_ms"] > 0 and info["rerank_ms"] > 0

    # the traversal is the same kernel on the same inputs: counters are identical
    for f in ("n_hop", "n_edge", "n_adc", "n_rerank"):
        assert np.array_equal(getattr(st_a, f), getattr(st_b, f)), f
    assert np.array_equal(cnt_a, cnt_b)
    # recall@10 against the exact fp32 ground truth: within 0.002 (north_star bar for bf16 recompute)
    vn = vectors / np.linalg.norm(vectors, axis=1, keepdims=True)
    qn = queries / np.linalg.norm(queries, axis=1, keepdims=True)
    gt = np.argsort(-(qn @ vn.T), axis=1, kind="stable")[:, :k]
    rec = lambda ids: np.mean([len(set(ids[i].tolist()) & set(gt[i].tolist())) / k for i in range(nq)])
    ra, rb = rec(ids_a), rec(ids_b)
    assert ra > 0.5
    assert abs(ra - rb) <= 0.002, (ra, rb)
    # distances: cosine of unit vectors with bf16 encoder noise
    same = ids_a == ids_b
    assert same.mean() > 0.9, same.mean()  # rank swaps among near-equal distances are the bf16 noise
    assert np.abs(dist_a[same] - dist_b[same]).max() < 2e-2

    # the stored vectors are not needed any more
    index.drop_vectors()
    ids_c, dist_c, _ = index.search_adc_recompute_batch(queries, k, ef)
    assert np.array_equal(ids_b, ids_c) and np.array_equal(dist_b.view(np.uint32), dist_c.view(np.uint32))


def test_recompute_errors(gpu_lib):
    from islands_b200 import (DimensionMismatch, Encoder, EncoderConfig, InvalidArgument, LeannConfig, LeannIndex, PQError)

    rng = np.random.RandomState(0)
    v = rng.randn(300, 128).astype(np.float32)
    index = LeannIndex(LeannConfig(m=8, m0=16, ef_construction=32))
    index.build(v, 300, seed=1, batch=16)
    with pytest.raises(PQError):
        index.search_adc_recompute_batch(v[:2], 5, 16)
    enc64 = Encoder(EncoderConfig(vocab_size=100, hidden_size=64, num_layers=1, num_heads=1, intermediate_size=128,
                                  max_position=16)).init_random()
    with pytest.raises(DimensionMismatch):
        index.set_recompute(enc64, np.ones((300, 8), np.int32), np.full(300, 8, np.int32))
    with pytest.raises(InvalidArgument):
        index.drop_vectors()  # nothing attached: refusing beats silently losing the only copy
