"""A second, independent reading of the reference's algorithms, against the oracle.

The C++ oracle (oracle/oracle.cpp) is a restatement of the Rust; the GPU kernels are checked against it bit for bit.
What pins the oracle itself is the handful of known answers the reference's tests hold plus *faithfulness of the
restatement*.  This file adds a second restatement written separately from the first — plain Python over `heapq`,
`set`, lists and scalar `numpy.float32` arithmetic, following the Rust line by line — and requires the two to agree
bit for bit:

* the search loop `leann.rs:899-988` with Global / Local frontier pruning (`:991-1016`) and the four distance folds
  (`distance.rs:71-122`): ids, distance bits and the reference's own work counter (`embeddings_computed`, :920, :950);
* graph construction `leann.rs:560-658, 661-749` with the hub-preserving selection `:761-833` — the part the reference's
  tests never exercise: whole CSR arrays, entry point, top level;
* `HnswGraph` insert / search (`hnsw.rs:214-504`) with the `prune_connections` quirk: every list of every layer;
* `ProductQuantizer` encode / decode / tables / table and asymmetric distance (`pq.rs:86-106, 221-348`);
* `train` + `kmeans` (`pq.rs:175-218, 362-463`): seeding, Lloyd iterations, empty-cluster re-seeding, on the reference's
  random stream — trained codebooks bit for bit;
* the batched construction ("round model") as DESIGN.md 3.3 words it — this repository's own widening of the sequential
  loop — for several round sizes, `batch = 1` coinciding with the sequential reading;
* the HNSW insert in rounds as DESIGN.md 3.5 words it;
* the two-level search as DESIGN.md 3.4 pins down the specification's Algorithm 2 (no reference code);
* "PQ ADC traversal + exact rerank" as include/islands_b200.h defines it (not a reference algorithm): checks that the
  oracle's twin implements the written definition, bfloat16 table rule included.

Small cases only (pure-Python loops); tie-heavy data where the reference's order is defined.  CPU.
"""
import heapq

import numpy as np
import pytest

from conftest import oracle_graph, uniform
from islands_b200 import LeannConfig

F = np.float32


def _fold(terms):
    acc = F(0.0)
    for t in terms:  # `sum()` / `+=` over an iterator: left to right, one rounding per add
        acc = F(acc + t)
    return acc


def distance(metric, a, b):
    """distance.rs:71-122, scalar f32, every multiply and add rounded on its own."""
    if metric == 0:  # cosine_distance
        dot = na = nb = F(0.0)
        for x, y in zip(a, b):
            dot = F(dot + F(x * y))
            na = F(na + F(x * x))
            nb = F(nb + F(y * y))
        norm = F(np.sqrt(F(na * nb)))
        return F(1.0) if norm == 0 else F(F(1.0) - F(dot / norm))
    if metric == 1:  # euclidean_distance = sqrt(euclidean_distance_squared)
        return F(np.sqrt(_fold(F(F(x - y) * F(x - y)) for x, y in zip(a, b))))
    if metric == 2:  # dot_product_distance
        return F(-_fold(F(x * y) for x, y in zip(a, b)))
    return _fold(F(abs(F(x - y))) for x, y in zip(a, b))  # manhattan_distance


def prune(cfg, candidates, results_len, ef):
    """apply_pruning_strategy (leann.rs:991-1016), Global and Local; the arithmetic is f32 as in the Rust."""
    ratio = F(cfg.prune_ratio)
    if ratio == 0 or not candidates:
        return list(candidates)
    if cfg.pruning_strategy == 0:  # Global
        fill = F(F(results_len) / F(ef))
        adjusted = int(np.ceil(F(F(len(candidates)) * F(F(1.0) - F(fill * ratio)))))
        return candidates[:max(adjusted, 1)]
    keep = int(np.ceil(F(F(len(candidates)) * F(F(1.0) - ratio))))  # Local
    return candidates[:max(keep, 1)]


MASK64 = (1 << 64) - 1


def _mix(z):  # splitmix64's finaliser, as written in include/islands_b200.h
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & MASK64
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


class SeededDraws:
    """The counter stream include/islands_b200.h defines for PruningStrategy::Proportional (the reference draws from
    thread_rng, leann.rs:1043): draw c of query q = 24 bits of mix(mix(seed + G * (q + 1)) + G * (c + 1))."""
    G = 0x9E3779B97F4A7C15

    def __init__(self, seed, query_index):
        self.base, self.c = _mix((seed + self.G * (query_index + 1)) & MASK64), 0

    def next_f32(self):
        h = _mix((self.base + self.G * (self.c + 1)) & MASK64)
        self.c += 1
        return F(F(h >> 40) * F(2.0 ** -24))


def prune_proportional(cfg, candidates, degree_of, draws):
    """apply_pruning_strategy, Proportional arm (leann.rs:1017-1053)."""
    if F(cfg.prune_ratio) == 0 or not candidates:
        return list(candidates)
    keep = max(int(np.ceil(F(F(len(candidates)) * F(F(1.0) - F(cfg.prune_ratio))))), 1)
    total = sum(degree_of(i) for i in candidates)
    if total == 0:
        return candidates[:keep]
    selected = []
    for i in candidates:
        prob = F(F(degree_of(i)) / F(total))
        if draws.next_f32() < F(prob * F(keep)):
            selected.append(i)
            if len(selected) >= keep:
                break
    return selected or [candidates[0]]


def search(cfg, vectors, offsets, nbrs, entry, query, k, ef, query_index=0):
    """search_with_params + search_layer_recompute (leann.rs:868-988).  The final order of exact distance ties is
    unspecified in the reference (a distance-only stable sort over heap order); like the oracle this uses (dist, id)."""
    ef = max(ef, k)
    metric = cfg.metric
    visited = {entry}
    d0 = distance(metric, query, vectors[entry])
    candidates = [(d0, entry)]                 # min-heap of (dist, id): Reverse<(OrderedFloat, u64)>
    results = [(-d0, -entry)]                  # max-heap of (dist, id) through negation
    computed = 1
    draws = SeededDraws(int(cfg.prune_seed), query_index)
    degree_of = lambda i: int(offsets[i + 1]) - int(offsets[i])  # CsrGraph::degree_counts
    while candidates:
        dist, node = heapq.heappop(candidates)
        if len(results) >= ef and dist > -results[0][0]:
            break
        unvisited = []
        for nb in nbrs[int(offsets[node]):int(offsets[node + 1])]:
            nb = int(nb)
            if nb not in visited:              # `filter(|&n| visited.insert(n))`: marked before pruning
                visited.add(nb)
                unvisited.append(nb)
        if not unvisited:
            continue
        if cfg.pruning_strategy == 2:
            to_compute = prune_proportional(cfg, unvisited, degree_of, draws)
        else:
            to_compute = prune(cfg, unvisited, len(results), ef)
        computed += len(to_compute)
        for nb in to_compute:
            nd = distance(metric, query, vectors[nb])
            if len(results) < ef or nd < -results[0][0]:
                heapq.heappush(candidates, (nd, nb))
                heapq.heappush(results, (-nd, -nb))
                if len(results) > ef:
                    heapq.heappop(results)
    out = sorted((-d, -i) for d, i in results)
    return [(i, d) for d, i in out[:k]], computed


def _compare(orc, cfg, v, off, nbrs, entry, queries, k, ef):
    ids, dist, cnt, st = orc.leann_search(cfg._s, v, off, nbrs, entry, queries, k, ef, stats=True)
    for qi, q in enumerate(queries):
        mine, computed = search(cfg, v, off, nbrs, entry, q, k, ef, query_index=qi)
        assert cnt[qi] == len(mine)
        assert [int(i) for i in ids[qi, :cnt[qi]]] == [i for i, _ in mine], qi
        assert [d.view(np.uint32) for d in dist[qi, :cnt[qi]]] == [F(d).view(np.uint32) for _, d in mine], qi
        assert int(st["n_dist"][qi]) == computed, qi  # the reference's own counter: embeddings_computed


@pytest.mark.parametrize("metric", [0, 1, 2, 3])
def test_search_loop_second_reading(orc, metric):
    cfg, v, _, off, nbrs, entry = oracle_graph(orc, 400, 12, seed=21, metric=metric, m=6, m0=12, ef_construction=32)
    q = uniform(np.random.RandomState(22 + metric), 12, 12)
    _compare(orc, cfg, v, off, nbrs, entry, q, 10, 24)
    _compare(orc, cfg, v, off, nbrs, entry, q[:4], 5, 3)     # ef < k is raised to k (leann.rs:890)
    _compare(orc, cfg, v, off, nbrs, entry, q[:3], 30, 500)  # ef > n: everything reachable is scored


@pytest.mark.parametrize("strategy,ratio", [(0, 0.3), (0, 0.9), (1, 0.5), (1, 0.8)])
def test_frontier_pruning_second_reading(orc, strategy, ratio):
    cfg, v, _, off, nbrs, entry = oracle_graph(orc, 400, 12, seed=21, metric=0, m=6, m0=12, ef_construction=32)
    pruned = LeannConfig(m=6, m0=12, ef_construction=32, prune_ratio=ratio, pruning_strategy=strategy)
    q = uniform(np.random.RandomState(5), 10, 12)
    _compare(orc, pruned, v, off, nbrs, entry, q, 10, 32)


@pytest.mark.parametrize("ratio,seed", [(0.3, 0), (0.6, 12345), (0.9, 2 ** 63 + 7)])
def test_proportional_pruning_second_reading(orc, ratio, seed):
    """The Proportional arm with the draws of the stream the header defines (query row and draw counter as keys)."""
    cfg, v, _, off, nbrs, entry = oracle_graph(orc, 400, 12, seed=21, metric=0, m=6, m0=12, ef_construction=32)
    pruned = LeannConfig(m=6, m0=12, ef_construction=32, prune_ratio=ratio, pruning_strategy=2, prune_seed=seed)
    q = uniform(np.random.RandomState(6), 10, 12)
    _compare(orc, pruned, v, off, nbrs, entry, q, 10, 32)
    u = [SeededDraws(seed, 3).next_f32() for _ in range(1)] + [SeededDraws(seed, 4).next_f32()]
    assert all(0 <= x < 1 for x in u) and u[0] != u[1]  # different query rows draw different streams


def test_exact_ties_second_reading(orc):
    """A third of the vectors are exact copies: distances tie bit for bit, pops and evictions are decided by the id."""
    cfg, v, _, off, nbrs, entry = oracle_graph(orc, 300, 8, seed=31, metric=0, dup=100, m=6, m0=12, ef_construction=32)
    q = np.concatenate([v[:6], uniform(np.random.RandomState(32), 6, 8)])
    _compare(orc, cfg, v, off, nbrs, entry, q, 16, 16)
    _compare(orc, cfg, v, off, nbrs, entry, q, 10, 40)


def test_distance_folds_second_reading(orc):
    rng = np.random.RandomState(41)
    for d in (1, 3, 16, 97):
        a, b = uniform(rng, 1, d)[0], uniform(rng, 1, d)[0]
        for metric in range(4):
            assert distance(metric, a, b).view(np.uint32) == F(orc.distance(metric, a, b)).view(np.uint32), (metric, d)
    z = np.zeros(5, F)
    assert distance(0, z, uniform(rng, 1, 5)[0]) == 1.0 == orc.distance(0, z, z)  # zero vector: cosine distance 1.0


# ---- construction: leann.rs:560-658 (insert loop, reverse edges, prune), :661-749 (insert search), :761-833 (hub-preserving
# selection — the part the reference's own tests never exercise) ---------------------------------------------

def insert_search(cfg, emb, adjacency, query, entry, ef):
    """search_layer_with_adjacency (leann.rs:692-749): the search loop without frontier pruning, over the temporary lists."""
    metric = cfg.metric
    visited = {entry}
    d0 = distance(metric, query, emb[entry])
    candidates, results = [(d0, entry)], [(-d0, -entry)]
    while candidates:
        dist, node = heapq.heappop(candidates)
        if len(results) >= ef and dist > -results[0][0]:
            break
        for nb in adjacency[node]:
            if nb in visited:
                continue
            visited.add(nb)
            nd = distance(metric, query, emb[nb])
            if len(results) < ef or nd < -results[0][0]:
                heapq.heappush(candidates, (nd, nb))
                heapq.heappush(results, (-nd, -nb))
                if len(results) > ef:
                    heapq.heappop(results)
    return [(i, d) for d, i in sorted((-d, -i) for d, i in results)]  # ascending; exact ties by id (the oracle's rule)


def hub_preserving_selection(cfg, candidates, adjacency, max_conn):
    """prune_with_degree_preservation_temp (leann.rs:761-833)."""
    if len(candidates) <= max_conn:
        return list(candidates)
    degrees = sorted((len(adjacency[i]) for i, _ in candidates), reverse=True)
    hub_count = int(np.ceil(F(F(len(degrees)) * F(cfg.hub_percentile))))
    threshold = degrees[hub_count - 1] if 0 < hub_count < len(degrees) else None  # None: usize::MAX, no hubs
    hubs, regular = [], []
    for i, d in candidates:
        deg = len(adjacency[i])
        if threshold is not None and deg >= threshold:
            hubs.append((i, d, deg))
        else:
            regular.append((i, d))
    hubs.sort(key=lambda t: -t[2])      # stable: equal degrees keep the candidates' (distance) order
    regular.sort(key=lambda t: t[1])    # stable, by distance
    hub_slots = max(max_conn // 4, 1)
    selected = [(i, d) for i, d, _ in hubs[:hub_slots]]
    for i, d in regular:
        if len(selected) >= max_conn:
            break
        if all(s != i for s, _ in selected):
            selected.append((i, d))
    for i, d, _ in hubs[hub_slots:]:
        if len(selected) >= max_conn:
            break
        if all(s != i for s, _ in selected):
            selected.append((i, d))
    return selected


def build(cfg, emb, levels):
    """LeannIndex::build (leann.rs:560-631) with the levels as an input (the reference draws them from thread_rng)."""
    n = len(emb)
    adjacency, entry, max_level = [], None, 0
    for node in range(n):
        if not adjacency:
            neighbors = []
        else:  # find_neighbors_for_insert_temp (leann.rs:661-689)
            cand = insert_search(cfg, emb, adjacency, emb[node], 0 if entry is None else entry, cfg.ef_construction)
            cand = hub_preserving_selection(cfg, cand, adjacency, cfg.m0) if cfg.high_degree_pruning else cand[:cfg.m0]
            neighbors = [i for i, _ in cand]
        adjacency.append(list(neighbors))
        for nb in neighbors:
            if node not in adjacency[nb]:
                adjacency[nb].append(node)
                if len(adjacency[nb]) > cfg.m0:  # prune_neighbors_temp (leann.rs:634-658): stable sort by distance, keep m0
                    scored = [(i, distance(cfg.metric, emb[nb], emb[i])) for i in adjacency[nb]]
                    scored.sort(key=lambda t: t[1])
                    adjacency[nb] = [i for i, _ in scored[:cfg.m0]]
        if entry is None or levels[node] > max_level:
            entry, max_level = node, int(levels[node])
    offsets = np.concatenate([[0], np.cumsum([len(a) for a in adjacency])]).astype(np.uint64)
    return offsets, np.array([i for a in adjacency for i in a], np.uint64), entry, max_level


@pytest.mark.parametrize("metric,hub,dup", [(0, True, 0), (1, True, 0), (0, False, 0), (3, True, 0), (0, True, 60), (2, True, 0)])
def test_construction_second_reading(orc, metric, hub, dup):
    """Small m0 and efC so that every branch runs many times: candidate lists longer than m0 (hub selection, with the
    degree-saturation degeneracy of SURVEY H6), reverse-edge overflow and pruning, entry-point moves, duplicate vectors."""
    n, d = 220, 8
    v = uniform(np.random.RandomState(50 + metric), n, d)
    if dup:
        v[n - dup:] = v[:dup]
    cfg = LeannConfig(metric=metric, m=4, m0=8, ef_construction=24, high_degree_pruning=int(hub), hub_percentile=0.1)
    levels = orc.draw_levels(9, n, cfg.ml, cfg.max_layers)
    assert levels.max() > 0  # the entry point moves at least once
    off, nbrs, entry, max_level = orc.leann_build(cfg._s, v, levels)
    m_off, m_nbrs, m_entry, m_max = build(cfg, v, levels)
    assert np.array_equal(off, m_off) and np.array_equal(nbrs, m_nbrs)
    assert (int(entry), int(max_level)) == (m_entry, m_max)
    deg = np.diff(off.astype(np.int64))
    assert deg.max() == cfg.m0 and (deg == cfg.m0).sum() > n // 4  # saturated lists: the pruning paths really ran


# ---- HnswGraph: hnsw.rs:214-329 (insert / insert_node), :332-402 (search_layer), :405-446 (prune_connections),
# :458-504 (search).  `Candidate` orders by distance alone (hnsw.rs:136-141), so the reference leaves exact ties to the
# heap's internals; these cases are tie-free (random data), where the reading is unambiguous. -----------------------------

class Hnsw:
    def __init__(self, cfg):
        self.cfg, self.nodes, self.entry, self.max_level = cfg, {}, None, 0  # nodes: id -> (vector, [list per layer])

    def _dist(self, q, i):
        return distance(self.cfg.metric, q, self.nodes[i][0])

    def _greedy(self, q, current, layers):
        cur_d = self._dist(q, current)
        for layer in layers:
            while True:
                changed = False
                conns = self.nodes[current][1]
                if layer < len(conns):
                    for nb in list(conns[layer]):  # the list of the node the sweep STARTED from, even after `current` moves
                        d = self._dist(q, nb)
                        if d < cur_d:
                            current, cur_d, changed = nb, d, True
                if not changed:
                    break
        return current

    def search_layer(self, q, entry, ef, layer):
        visited = {entry}
        d0 = self._dist(q, entry)
        candidates, results = [(d0, entry)], [(-d0, -entry)]
        while candidates:
            d, node = heapq.heappop(candidates)
            if d > -results[0][0] and len(results) >= ef:
                break
            conns = self.nodes[node][1]
            if layer < len(conns):
                for nb in conns[layer]:
                    if nb in visited:
                        continue
                    visited.add(nb)
                    nd = self._dist(q, nb)
                    if len(results) < ef or nd < -results[0][0]:
                        heapq.heappush(candidates, (nd, nb))
                        heapq.heappush(results, (-nd, -nb))
                        if len(results) > ef:
                            heapq.heappop(results)
        return [(i, d) for d, i in sorted((-d, -i) for d, i in results)]

    def insert(self, vector, level):
        node = len(self.nodes)
        conns = [[] for _ in range(level + 1)]
        if self.entry is None:
            self.entry, self.max_level = node, level
            self.nodes[node] = (vector, conns)
            return node
        current = self._greedy(vector, self.entry, range(self.max_level, level, -1))
        for layer in range(level, -1, -1):
            found = self.search_layer(vector, current, self.cfg.ef_construction, layer)
            m = self.cfg.m0 if layer == 0 else self.cfg.m
            selected = [i for i, _ in found[:m]]
            conns[layer] = list(selected)
            for nb in selected:
                nb_conns = self.nodes[nb][1]
                if layer < len(nb_conns):
                    nb_conns[layer].append(node)
                    if len(nb_conns[layer]) > m:  # prune_connections: ids without a stored node are skipped, and the node
                        scored = [(i, distance(self.cfg.metric, self.nodes[nb][0], self.nodes[i][0]))  # being inserted is
                                  for i in nb_conns[layer] if i in self.nodes]                          # not stored yet
                        scored.sort(key=lambda t: t[1])
                        nb_conns[layer] = [i for i, _ in scored[:m]]
            if selected:
                current = selected[0]
        if level > self.max_level:
            self.max_level, self.entry = level, node
        self.nodes[node] = (vector, conns)
        return node

    def search(self, q, k, ef):
        if not self.nodes:
            return []
        current = self._greedy(q, self.entry, range(self.max_level, 0, -1))
        return self.search_layer(q, current, max(ef, k), 0)[:k]


@pytest.mark.parametrize("metric", [0, 1])
def test_hnsw_second_reading(orc, metric):
    from islands_b200 import HnswConfig

    n, d = 260, 8
    v = uniform(np.random.RandomState(70 + metric), n, d)
    cfg = HnswConfig(m=4, m0=8, ef_construction=20, metric=metric, ml=0.9)  # ml 0.9: several layers at this size
    levels = orc.draw_levels(13, n, cfg.ml, cfg.max_layers)
    assert levels.max() >= 2
    theirs, mine = orc.Hnsw(cfg._s, d), Hnsw(cfg)
    for i in range(n):
        assert theirs.insert(v[i], int(levels[i])) == mine.insert(v[i], int(levels[i])) == i
    assert (theirs.entry_point(), theirs.max_level()) == (mine.entry, mine.max_level)
    for i in range(n):
        assert theirs.node_level(i) == len(mine.nodes[i][1]) - 1
        for layer, conns in enumerate(mine.nodes[i][1]):
            assert theirs.neighbors(i, layer).tolist() == conns, (i, layer)
        assert theirs.neighbors(i, len(mine.nodes[i][1])) is None
    # the prune path ran: many layer-0 lists reached m0 (from then on a later node never enters them — the quirk)
    assert sum(1 for i in range(n) if len(mine.nodes[i][1][0]) == cfg.m0) > n // 8
    q = uniform(np.random.RandomState(71), 20, d)
    ids, dist, cnt = theirs.search(q, 5, 30)
    for qi in range(20):
        got = mine.search(q[qi], 5, 30)
        assert cnt[qi] == len(got) and ids[qi, :cnt[qi]].tolist() == [i for i, _ in got]
        assert [x.view(np.uint32) for x in dist[qi, :cnt[qi]]] == [F(x).view(np.uint32) for _, x in got]


# ---- ProductQuantizer: pq.rs:86-106 (find_nearest), :221-271 (encode / decode), :275-348 (asymmetric distance, tables,
# table_distance) -------------------------------------------------------------------------------------------------------

def find_nearest(codebook, sub, metric):
    best, best_d = 0, F(np.finfo(np.float32).max)  # f32::MAX
    for i, centroid in enumerate(codebook):
        d = distance(metric, sub, centroid)
        if d < best_d:  # strict: the first of equal centroids wins
            best, best_d = i, d
    return best


def squared_sub(a, b):
    return _fold(F(F(x - y) * F(x - y)) for x, y in zip(a, b))  # `(a - b).powi(2)` summed left to right


@pytest.mark.parametrize("metric", [1, 0, 3])
def test_pq_second_reading(orc, metric):
    rng = np.random.RandomState(80 + metric)
    m, ksub, dsub = 4, 9, 3
    cb = uniform(rng, m * ksub, dsub).reshape(m, ksub, dsub)
    cb[1, 5] = cb[1, 2]  # duplicate centroid: the strict `<` keeps the lower index
    v = uniform(rng, 30, m * dsub)
    v[7, dsub:2 * dsub] = cb[1, 2]
    codes = orc.pq_encode(metric, cb, v)
    mine = np.array([[find_nearest(cb[s], row[s * dsub:(s + 1) * dsub], metric) for s in range(m)] for row in v])
    assert np.array_equal(codes, mine) and codes[7, 1] == 2
    assert np.array_equal(orc.pq_decode(cb, codes), np.stack([np.concatenate([cb[s, c] for s, c in enumerate(row)]) for row in mine]))
    q = uniform(rng, 1, m * dsub)[0]
    tables = orc.pq_build_tables(cb, q)
    my_tables = np.array([[squared_sub(q[s * dsub:(s + 1) * dsub], c) for c in cb[s]] for s in range(m)], F)
    assert np.array_equal(tables.view(np.uint32), my_tables.view(np.uint32))
    for row in mine[:10]:
        total = F(0.0)
        for s, c in enumerate(row):  # asymmetric_distance: `total_dist += sub_dist`, then sqrt
            total = F(total + squared_sub(q[s * dsub:(s + 1) * dsub], cb[s, c]))
        adc = F(np.sqrt(total))
        looked_up = F(np.sqrt(_fold(my_tables[s, c] for s, c in enumerate(row))))
        one = row.astype(np.uint16)[None, :]
        assert orc.pq_asymmetric_distance(cb, q, one)[0].view(np.uint32) == adc.view(np.uint32)
        assert orc.pq_table_distance(tables, one)[0].view(np.uint32) == looked_up.view(np.uint32)
        assert adc.view(np.uint32) == looked_up.view(np.uint32)  # same terms, same order: the two routes agree bit for bit


# ---- ProductQuantizer::train + kmeans: pq.rs:175-218, 362-463.  The random stream is the reference's generator (rand 0.8.5
# StdRng, restated in the oracle and pinned by tests/test_std_rng.py); what is read a second time here is everything that
# CONSUMES it: the distance-weighted (not squared) seeding, the cumulative-sum pick, Lloyd's iterations in f32, the
# re-seeding of empty clusters, one generator running through all subquantizers. ---------------------------------------

class ScriptedRng:
    """Draws from `orc_std_rng_draw` — a stateless scripted interface — by replaying the growing script."""

    def __init__(self, orc, seed, bound):
        import ctypes as C

        self.C, self.seed, self.bound, self.kinds = C, seed, bound, []
        self.fn = orc.lib().orc_std_rng_draw
        self.fn.restype = None
        self.fn.argtypes = [C.c_uint64, C.POINTER(C.c_uint8), C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]

    def _next(self, kind):
        self.kinds.append(kind)
        kinds = np.array(self.kinds, np.uint8)
        out = np.zeros(len(kinds), np.uint64)
        self.fn(self.seed, kinds.ctypes.data_as(self.C.POINTER(self.C.c_uint8)), len(kinds), self.bound,
                out.ctypes.data_as(self.C.POINTER(self.C.c_uint64)))
        return int(out[-1])

    def gen_usize(self):
        return self._next(1)

    def gen_f32(self):
        return np.array([self._next(2)], np.uint32).view(np.float32)[0]

    def choose(self):
        return self._next(3)


def kmeans(vectors, k, iterations, metric, rng):
    k = min(k, len(vectors))
    centroids = [vectors[rng.gen_usize() % len(vectors)].copy()]
    f32_max = F(np.finfo(np.float32).max)
    while len(centroids) < k:
        dists = []
        for v in vectors:
            best = f32_max
            for c in centroids:
                best = min(best, distance(metric, v, c))  # fold(f32::MAX, f32::min)
            dists.append(best)
        total = _fold(dists)
        if total > 0:
            dists = [F(x / total) for x in dists]
        threshold, cumsum, selected = rng.gen_f32(), F(0.0), 0
        for i, x in enumerate(dists):
            cumsum = F(cumsum + x)
            if cumsum >= threshold:
                selected = i
                break
        centroids.append(vectors[selected].copy())
    for _ in range(iterations):
        assign = [find_nearest(centroids, v, metric) for v in vectors]  # same strict `<` scan from f32::MAX
        sums = [np.zeros(len(vectors[0]), F) for _ in range(k)]
        counts = [0] * k
        for v, c in zip(vectors, assign):
            counts[c] += 1
            sums[c] = (sums[c] + v).astype(F)  # elementwise f32 adds in vector order
        for c in range(k):
            if counts[c] > 0:
                sums[c] = (sums[c] / F(counts[c])).astype(F)
            else:
                sums[c] = vectors[rng.choose()].copy()
        centroids = sums
    return centroids


@pytest.mark.parametrize("metric,seed", [(1, 42), (0, 7)])
def test_kmeans_second_reading(orc, metric, seed):
    rng = np.random.RandomState(90 + metric)
    m, dsub, n, ksub, iterations = 3, 2, 14, 6, 4
    v = uniform(rng, n, m * dsub)
    if seed == 42:  # only three distinct rows for six centroids: seeding runs out of distance mass (total == 0, the pick
        v = v[rng.randint(0, 3, size=n)]  # falls through to row 0), duplicate centroids stay empty and are re-seeded
    theirs = orc.pq_train(metric, v, m, ksub, iterations, seed)
    stream = ScriptedRng(orc, seed, n)
    mine = np.stack([np.stack(kmeans([row[s * dsub:(s + 1) * dsub] for row in v], ksub, iterations, metric, stream))
                     for s in range(m)])
    assert theirs.shape == mine.shape == (m, ksub, dsub)
    assert np.array_equal(theirs.view(np.uint32), mine.view(np.uint32))
    assert (stream.kinds.count(3) > 0) == (seed == 42)  # empty clusters: the `choose` path, taken in the degenerate case
    assert stream.kinds.count(1) == m and stream.kinds.count(2) == m * (ksub - 1)


# ---- "PQ ADC traversal + exact rerank" — not a reference algorithm (leann.rs:54-56 only names it): the mode is DEFINED in
# include/islands_b200.h.  Read a second time from that definition: the search loop above with the table distance over
# bfloat16-rounded table entries as the key, then exact distances for the ef survivors, sorted by (distance, id).  This
# checks that the oracle's twin implements what the header says, nothing more (parity of the mode stays unpinned). ------

def bf16_round(x):
    bits = int(F(x).view(np.uint32))
    if (bits & 0x7FFFFFFF) > 0x7F800000:
        return np.array([(bits | 0x00400000) & 0xFFFF0000], np.uint32).view(F)[0]  # NaN -> quiet NaN
    bits = (bits + 0x7FFF + ((bits >> 16) & 1)) & 0xFFFF0000  # round to nearest even on the bit pattern
    return np.array([bits], np.uint32).view(F)[0]


def adc_rerank(cfg, vectors, offsets, nbrs, entry, cb, codes, query, k, ef):
    m, ksub, dsub = cb.shape
    table = [[bf16_round(squared_sub(query[s * dsub:(s + 1) * dsub], c)) for c in cb[s]] for s in range(m)]

    def adc(node):  # table_distance (pq.rs:341-348) over the rounded entries
        return F(np.sqrt(_fold(table[s][int(codes[node, s])] for s in range(m))))

    ef = max(ef, k)
    visited = {entry}
    d0 = adc(entry)
    candidates, results = [(d0, entry)], [(-d0, -entry)]
    while candidates:
        d, node = heapq.heappop(candidates)
        if len(results) >= ef and d > -results[0][0]:
            break
        for nb in nbrs[int(offsets[node]):int(offsets[node + 1])]:
            nb = int(nb)
            if nb in visited:
                continue
            visited.add(nb)
            nd = adc(nb)
            if len(results) < ef or nd < -results[0][0]:
                heapq.heappush(candidates, (nd, nb))
                heapq.heappush(results, (-nd, -nb))
                if len(results) > ef:
                    heapq.heappop(results)
    exact = sorted((distance(cfg.metric, query, vectors[-i]), -i) for _, i in results)
    return [(i, d) for d, i in exact[:k]], len(visited)


def test_adc_traversal_rerank_second_reading(orc):
    cfg, v, _, off, nbrs, entry = oracle_graph(orc, 400, 12, seed=21, metric=0, m=6, m0=12, ef_construction=32)
    cb = orc.pq_train(1, v, 4, 16, 5, 3)
    codes = orc.pq_encode(1, cb, v)
    q = uniform(np.random.RandomState(95), 10, 12)
    ids, dist, cnt, st = orc.leann_search_adc_rerank(cfg._s, v, off, nbrs, entry, cb, codes, q, 10, 40, stats=True)
    for qi in range(len(q)):
        mine, scored = adc_rerank(cfg, v, off, nbrs, entry, cb, codes, q[qi], 10, 40)
        assert cnt[qi] == len(mine) and ids[qi, :cnt[qi]].tolist() == [i for i, _ in mine], qi
        assert [x.view(np.uint32) for x in dist[qi, :cnt[qi]]] == [F(x).view(np.uint32) for _, x in mine], qi
        assert int(st["n_adc"][qi]) == scored and int(st["n_rerank"][qi]) == min(40, scored)
    vals = np.array([1.0, 1.00390625, 1.005859375, 3.3895314e38, -2.5e-41, 0.1], F)
    assert np.array_equal(orc.adc_table_round(vals).view(np.uint32), np.array([bf16_round(x) for x in vals], F).view(np.uint32))


# ---- the batched construction ("round model") — this repository's own definition of how the reference's
# sequential loop is widened for the GPU; `batch = 1` is the reference loop itself.  Read a second time from the text:
# rounds of min(batch, max(1, inserted / 2)) nodes; every node of a round searches the graph as it stood before the
# round (entry point included) and selects its neighbours; the reverse edges of the round are applied per target in
# ascending source id with the reference's rule (append, prune to the m0 closest on overflow). ---------------------------

def build_rounds(cfg, emb, levels, batch):
    n = len(emb)
    adjacency, entry, max_level, inserted = [], None, 0, 0
    while inserted < n:
        size = min(batch, max(1, inserted // 2), n - inserted)
        round_nodes = range(inserted, inserted + size)
        frozen = [list(a) for a in adjacency]  # the graph before the round
        chosen = {}
        for node in round_nodes:
            if not frozen:
                chosen[node] = []
                continue
            cand = insert_search(cfg, emb, frozen, emb[node], 0 if entry is None else entry, cfg.ef_construction)
            cand = hub_preserving_selection(cfg, cand, frozen, cfg.m0) if cfg.high_degree_pruning else cand[:cfg.m0]
            chosen[node] = [i for i, _ in cand]
        for node in round_nodes:
            adjacency.append(list(chosen[node]))
        incoming = sorted((target, node) for node in round_nodes for target in chosen[node])
        for target, node in incoming:  # per target, ascending source id
            if node not in adjacency[target]:
                adjacency[target].append(node)
                if len(adjacency[target]) > cfg.m0:
                    scored = [(i, distance(cfg.metric, emb[target], emb[i])) for i in adjacency[target]]
                    scored.sort(key=lambda t: t[1])
                    adjacency[target] = [i for i, _ in scored[:cfg.m0]]
        for node in round_nodes:
            if entry is None or levels[node] > max_level:
                entry, max_level = node, int(levels[node])
        inserted += size
    offsets = np.concatenate([[0], np.cumsum([len(a) for a in adjacency])]).astype(np.uint64)
    return offsets, np.array([i for a in adjacency for i in a], np.uint64), entry, max_level


@pytest.mark.parametrize("batch,hub", [(1, True), (8, True), (64, True), (16, False)])
def test_round_model_second_reading(orc, batch, hub):
    n, d = 200, 8
    v = uniform(np.random.RandomState(61), n, d)
    cfg = LeannConfig(metric=0, m=4, m0=8, ef_construction=24, high_degree_pruning=int(hub), hub_percentile=0.1)
    levels = orc.draw_levels(9, n, cfg.ml, cfg.max_layers)
    off, nbrs, entry, max_level = orc.leann_build(cfg._s, v, levels, batch=batch)
    m_off, m_nbrs, m_entry, m_max = build_rounds(cfg, v, levels, batch)
    assert np.array_equal(off, m_off) and np.array_equal(nbrs, m_nbrs) and (int(entry), int(max_level)) == (m_entry, m_max)
    if batch == 1:  # one node per round IS the sequential loop
        s_off, s_nbrs, s_entry, s_max = build(cfg, v, levels)
        assert np.array_equal(off, s_off) and np.array_equal(nbrs, s_nbrs) and (s_entry, s_max) == (m_entry, m_max)


# ---- two-level search (docs/leann-specification.md:223-269 as DESIGN.md 3.4 pins it down; no reference code): the
# oracle's twin against a reading of that paragraph. ----------------------------------------------------------------------

def two_level(cfg, vectors, offsets, nbrs, entry, cb, codes, query, k, ef, a):
    m, ksub, dsub = cb.shape
    table = [[squared_sub(query[s * dsub:(s + 1) * dsub], c) for c in cb[s]] for s in range(m)]  # f32 tables

    def adc(node):
        return F(np.sqrt(_fold(table[s][int(codes[node, s])] for s in range(m))))

    ef = max(ef, k)
    visited = {entry}
    d0 = distance(cfg.metric, query, vectors[entry])
    exact_queue, results, approx_queue = [(d0, entry)], [(-d0, -entry)], []
    n_adc = n_rerank = 0
    while exact_queue:
        d, node = heapq.heappop(exact_queue)
        if len(results) >= ef and d > -results[0][0]:
            break
        for nb in nbrs[int(offsets[node]):int(offsets[node + 1])]:
            nb = int(nb)
            if nb not in visited:
                visited.add(nb)
                heapq.heappush(approx_queue, (adc(nb), nb))
                n_adc += 1
        if not approx_queue:
            continue
        promote = min(max(int(np.ceil(F(F(len(approx_queue)) * F(a)))), 1), len(approx_queue))
        for _ in range(promote):
            _, nb = heapq.heappop(approx_queue)
            nd = distance(cfg.metric, query, vectors[nb])
            n_rerank += 1
            if len(results) < ef or nd < -results[0][0]:
                heapq.heappush(exact_queue, (nd, nb))
                heapq.heappush(results, (-nd, -nb))
                if len(results) > ef:
                    heapq.heappop(results)
    return [(i, d) for d, i in sorted((-d, -i) for d, i in results)[:k]], n_adc, n_rerank


@pytest.mark.parametrize("a", [0.1, 0.5, 1.0])
def test_two_level_second_reading(orc, a):
    cfg, v, _, off, nbrs, entry = oracle_graph(orc, 400, 12, seed=21, metric=0, m=6, m0=12, ef_construction=32)
    cb = orc.pq_train(1, v, 4, 16, 5, 3)
    codes = orc.pq_encode(1, cb, v)
    q = uniform(np.random.RandomState(96), 8, 12)
    ids, dist, cnt, st = orc.leann_search_two_level(cfg._s, v, off, nbrs, entry, cb, codes, q, 10, 40, a, stats=True)
    for qi in range(len(q)):
        mine, n_adc, n_rerank = two_level(cfg, v, off, nbrs, entry, cb, codes, q[qi], 10, 40, a)
        assert cnt[qi] == len(mine) and ids[qi, :cnt[qi]].tolist() == [i for i, _ in mine], qi
        assert [x.view(np.uint32) for x in dist[qi, :cnt[qi]]] == [F(x).view(np.uint32) for _, x in mine], qi
        assert (int(st["n_adc"][qi]), int(st["n_rerank"][qi]), int(st["n_dist"][qi])) == (n_adc, n_rerank, n_rerank + 1)
    if a == 1.0:  # everything scored is promoted: the exact search over the same graph, node for node
        e_ids, e_dist, _ = orc.leann_search(cfg._s, v, off, nbrs, entry, q, 10, 40)
        assert np.array_equal(ids, e_ids) and np.array_equal(dist.view(np.uint32), e_dist.view(np.uint32))


# ---- HnswGraph in rounds: the read-only half of insert_node for all nodes of a round against the
# pre-round graph, then the mutating half node by node in id order. ------------------------------------------------------

def hnsw_insert_rounds(g, vectors, levels, batch):
    n, inserted = len(vectors), len(g.nodes)
    while inserted < n:
        size = 1 if not g.nodes else min(batch, max(1, len(g.nodes) // 2), n - inserted)
        nodes = range(inserted, inserted + size)
        plans = {}
        for node in nodes:  # read-only half, against the graph as it stands before the round
            level = int(levels[node])
            plan = [[] for _ in range(level + 1)]
            if g.entry is not None:
                current = g._greedy(vectors[node], g.entry, range(g.max_level, level, -1))
                for layer in range(level, -1, -1):
                    found = g.search_layer(vectors[node], current, g.cfg.ef_construction, layer)
                    plan[layer] = [i for i, _ in found[:g.cfg.m0 if layer == 0 else g.cfg.m]]
                    if plan[layer]:
                        current = plan[layer][0]
            plans[node] = plan
        for node in nodes:  # mutating half, node by node
            for layer in range(len(plans[node]) - 1, -1, -1):
                m = g.cfg.m0 if layer == 0 else g.cfg.m
                for target in plans[node][layer]:
                    conns = g.nodes[target][1]
                    if layer < len(conns):
                        conns[layer].append(node)
                        if len(conns[layer]) > m:
                            scored = [(i, distance(g.cfg.metric, g.nodes[target][0], g.nodes[i][0]))
                                      for i in conns[layer] if i in g.nodes]
                            scored.sort(key=lambda t: t[1])
                            conns[layer] = [i for i, _ in scored[:m]]
            level = int(levels[node])
            if g.entry is None:
                g.entry, g.max_level = node, level
            elif level > g.max_level:
                g.max_level, g.entry = level, node
            g.nodes[node] = (vectors[node], plans[node])
        inserted += size


@pytest.mark.parametrize("batch", [1, 8, 32])
def test_hnsw_round_model_second_reading(orc, batch):
    from islands_b200 import HnswConfig

    n, d = 260, 8
    v = uniform(np.random.RandomState(70), n, d)
    cfg = HnswConfig(m=4, m0=8, ef_construction=20, metric=0, ml=0.9)
    levels = orc.draw_levels(13, n, cfg.ml, cfg.max_layers)
    theirs, mine = orc.Hnsw(cfg._s, d), Hnsw(cfg)
    theirs.insert_batch(v, levels, batch=batch)
    hnsw_insert_rounds(mine, v, levels, batch)
    assert (theirs.entry_point(), theirs.max_level(), len(theirs)) == (mine.entry, mine.max_level, n)
    for i in range(n):
        for layer, conns in enumerate(mine.nodes[i][1]):
            assert theirs.neighbors(i, layer).tolist() == conns, (i, layer)
    if batch == 1:  # one node per round is the sequential insert
        seq = Hnsw(cfg)
        for i in range(n):
            seq.insert(v[i], int(levels[i]))
        assert all(seq.nodes[i][1] == mine.nodes[i][1] for i in range(n))


def test_search_on_arbitrary_graphs_second_reading(orc):
    """Graphs no construction would produce — repeated ids inside a list, self loops, empty lists, unreachable nodes, an
    entry point with no edges — and coarse-grid vectors (many exact distance ties): the loop's semantics alone decide
    (`visited.insert` keeps the first of repeated ids; ties fall to the id), for every metric and both deterministic
    pruning strategies."""
    rng = np.random.RandomState(123)
    for trial in range(40):
        n, d = int(rng.randint(1, 40)), int(rng.randint(1, 6))
        v = rng.randint(-2, 3, size=(n, d)).astype(np.float32)  # coarse grid, zero vectors included
        lists = [rng.randint(0, n, size=int(rng.randint(0, 9))).tolist() for _ in range(n)]
        off = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.uint64)
        nbrs = np.array([i for x in lists for i in x], np.uint64)
        entry = int(rng.randint(0, n))
        cfg = LeannConfig(metric=int(rng.randint(0, 4)), pruning_strategy=int(rng.randint(0, 2)),
                          prune_ratio=float(rng.choice([0.0, 0.0, 0.4, 0.9])))
        q = rng.randint(-2, 3, size=(3, d)).astype(np.float32)
        k = int(rng.randint(1, 12))
        _compare(orc, cfg, v, off, nbrs, entry, q, k, int(rng.randint(1, 20)))
