"""Handle-level behaviour of the C ABI that the reference's API implies:
  * any number of exact distance ties (leann.rs:924-928: the candidate heap is unbounded) — corpora full of
    duplicate embeddings must search like the oracle, not fail;
  * `&self` search: several threads may search one handle at once (src/core/mod.rs re-exports are Send + Sync);
  * device inputs produced asynchronously on the caller's stream are waited for (isl_set_caller_stream);
  * after isl_index_drop_vectors only the recompute search remains, every other entry point says so;
  * CsrGraph::set_neighbors (leann.rs:256-293)."""
import threading

import numpy as np
import pytest

from conftest import oracle_graph, uniform

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("ef", [16, 48, 200])
def test_more_duplicates_than_ef_plus_64(gpu_lib, orc, metric, ef):
    """One vector stored 700 times plus 500 others: every search that reaches the copies sees hundreds of exact
    ties with the worst result distance (the old 64-entry tie list failed the whole batch here)."""
    from islands_b200 import LeannConfig, LeannIndex

    rng = np.random.RandomState(31)
    n, d = 1200, 24
    v = uniform(rng, n, d)
    v[500:] = v[7]            # 700 copies of row 7 (+ the original)
    v[100:140] = 0.0          # zero vectors: cosine distance 1.0 to everything (distance.rs:82-84)
    cfg = LeannConfig(metric=metric, m=8, m0=16, ef_construction=32)
    levels = orc.draw_levels(3, n, cfg.ml, cfg.max_layers)
    off, nbrs, entry, _ = orc.leann_build(cfg._s, v, levels, batch=16, threads=8)
    q = np.concatenate([v[7:8], v[100:101], v[:20], uniform(np.random.RandomState(2), 40, d)])
    idx = LeannIndex.from_csr(cfg, v, off, nbrs, levels, entry)
    for k in (10, ef):
        ids, dist, cnt, st = idx.search_batch(q, k, ef, stats=True)
        o_ids, o_dist, o_cnt, o_st = orc.leann_search(cfg._s, v, off, nbrs, entry, q, k, ef, threads=8, stats=True)
        assert np.array_equal(cnt, o_cnt)
        assert np.array_equal(ids, o_ids)
        assert np.array_equal(dist.view(np.uint32), o_dist.view(np.uint32))
        for f in ("n_hop", "n_edge", "n_dist"):
            assert np.array_equal(getattr(st, f), o_st[f]), f
    # the GPU construction over the same data goes through the same tie handling: CSR bit-exact with the oracle twin
    built = LeannIndex(cfg)
    built.build(v, n, levels=levels, batch=16)
    g = built.graph
    assert np.array_equal(g.node_offsets, off) and np.array_equal(g.neighbors, nbrs) and g.entry_point == entry
    idx.free()
    built.free()


def test_concurrent_searches_on_one_handle(gpu_lib, orc):
    """Eight threads, one handle, different ef / k per thread: every result equals the single-threaded oracle."""
    from islands_b200 import LeannIndex

    cfg, v, levels, off, nbrs, entry = oracle_graph(orc, 4000, 96, seed=11, metric=0)
    idx = LeannIndex.from_csr(cfg, v, off, nbrs, levels, entry)
    q = uniform(np.random.RandomState(5), 300, 96)
    plans = [(10, 32), (10, 64), (5, 200), (20, 20), (10, 3000), (1, 16), (10, 128), (50, 50)]
    expect = [orc.leann_search(cfg._s, v, off, nbrs, entry, q, k, ef, threads=8) for k, ef in plans]
    errors = []

    def work(t):
        try:
            k, ef = plans[t]
            for _ in range(6):
                ids, dist, cnt = idx.search_batch(q, k, ef)
                assert np.array_equal(ids, expect[t][0]) and np.array_equal(cnt, expect[t][2])
                assert np.array_equal(dist.view(np.uint32), expect[t][1].view(np.uint32))
        except BaseException as ex:  # noqa: BLE001 - reported by the main thread
            errors.append((t, repr(ex)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(len(plans))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    idx.free()


def test_dev_inputs_are_ordered_after_the_callers_stream(gpu_lib, orc):
    """Queries written by a slow kernel chain on a side stream: the `_dev` search must wait for them on the device."""
    import torch

    from islands_b200 import LeannIndex
    from islands_b200.core import set_caller_stream

    cfg, v, levels, off, nbrs, entry = oracle_graph(orc, 4000, 96, seed=11, metric=0)
    idx = LeannIndex.from_csr(cfg, v, off, nbrs, levels, entry)
    dev = torch.device("cuda", 0)
    nq, d, k, ef = 256, 96, 10, 64
    q = uniform(np.random.RandomState(6), nq, d)
    o_ids, _, _ = orc.leann_search(cfg._s, v, off, nbrs, entry, q, k, ef, threads=8)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    dst = torch.empty((nq, k), dtype=torch.float32, device=dev)
    cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
    src = torch.from_numpy(q).to(dev)
    big = torch.randn((6144, 6144), device=dev)
    for stream in (None, torch.cuda.Stream(device=dev)):
        tq = torch.zeros((nq, d), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        with torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.default_stream(dev)):
            for _ in range(8):
                big = big @ big * 1e-3  # tens of milliseconds of work ahead of the copy
            tq.copy_(src, non_blocking=True)
            set_caller_stream(stream.cuda_stream if stream is not None else 0)
            idx.search_batch_dev(tq.data_ptr(), nq, d, k, ef, ids.data_ptr(), dst.data_ptr(), cnt.data_ptr())
        assert np.array_equal(ids.cpu().numpy().view(np.uint64), o_ids)
    set_caller_stream(0)
    idx.free()


def test_dropped_vectors_are_refused_everywhere_but_recompute(gpu_lib):
    from islands_b200 import Encoder, EncoderConfig, InvalidArgument, LeannConfig, LeannIndex, PQConfig, ProductQuantizer

    rng = np.random.RandomState(0)
    n, S = 400, 8
    enc = Encoder(EncoderConfig(vocab_size=500, hidden_size=128, num_layers=1, num_heads=2, intermediate_size=256,
                                max_position=16)).init_random(seed=1, stddev=0.05)
    tok = rng.randint(1, 500, size=(n, S)).astype(np.int32)
    ln = np.full(n, S, np.int32)
    v = enc.embed(tok, ln)
    index = LeannIndex(LeannConfig(m=8, m0=16, ef_construction=32))
    index.build(v, n, seed=1, batch=16)
    pq = ProductQuantizer(128, PQConfig(16, 32, 5, 1))
    pq.train(v)
    index.attach_pq(pq, pq.encode(v))
    index.set_recompute(enc, tok, ln)
    before = index.search_adc_recompute_batch(v[:8], 5, 32)
    index.drop_vectors()
    for call in (lambda: index.search_batch(v[:8], 5, 32),
                 lambda: index.search_two_level_batch(v[:8], 5, 32, 0.5),
                 lambda: index.search_adc_rerank_batch(v[:8], 5, 32),
                 lambda: index.set_recompute(None, None, None)):
        with pytest.raises(InvalidArgument):
            call()
    after = index.search_adc_recompute_batch(v[:8], 5, 32)  # the context is intact and the one remaining search works
    assert np.array_equal(before[0], after[0]) and np.array_equal(before[1].view(np.uint32), after[1].view(np.uint32))
    index.free()


def test_set_neighbors_matches_the_reference_semantics(gpu_lib, orc):
    from islands_b200 import LeannIndex, NodeNotFound

    cfg, v, levels, off, nbrs, entry = oracle_graph(orc, 1000, 32, seed=11, metric=0)
    idx = LeannIndex.from_csr(cfg, v, off, nbrs, levels, entry)
    g = idx.graph
    # same length: overwrite in place; different length: offsets and neighbours rebuilt (leann.rs:267-290)
    same = np.asarray(g.get_neighbors(5))[::-1].copy()
    shorter = np.asarray(g.get_neighbors(9))[:3].copy()
    longer = np.concatenate([np.asarray(g.get_neighbors(11)), np.array([1, 2, 3], np.uint64)])
    for node, new in ((5, same), (9, shorter), (11, longer), (10_000, same)):  # the last one is ignored (leann.rs:258-260)
        g.set_neighbors(node, new)
        idx.set_neighbors(node, new)
    h = idx.graph
    assert np.array_equal(h.node_offsets, g.node_offsets) and np.array_equal(h.neighbors, g.neighbors)
    assert np.array_equal(h.degree_counts, g.degree_counts)
    q = uniform(np.random.RandomState(3), 64, 32)
    ids, dist, cnt = idx.search_batch(q, 10, 64)
    o_ids, o_dist, o_cnt = orc.leann_search(cfg._s, v, g.node_offsets, g.neighbors, entry, q, 10, 64)
    assert np.array_equal(ids, o_ids) and np.array_equal(dist.view(np.uint32), o_dist.view(np.uint32))
    with pytest.raises(NodeNotFound):
        idx.set_neighbors(3, [999_999])
    idx.free()
