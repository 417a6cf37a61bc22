"""Host-side mirror of the reference's search interface (src/core/search.rs): SearchConfig,
SearchResult, Searcher, MultiIndexSearcher.  Thin wrappers: graph search and the island merge run on
the GPU (isl_hnsw_search / isl_index_search / isl_merge_topk); this module only shapes results.
"""
import ctypes as C

import numpy as np

from . import _ffi
from .core import _check, _f32, _ptr, merge_topk
from ._ffi import f32p


class SearchConfig:
    """SearchConfig (search.rs:9-46)."""

    def __init__(self, top_k=10, ef=100, include_vectors=False, include_metadata=True, min_similarity=None):
        self.top_k = top_k
        self.ef = ef
        self.include_vectors = include_vectors
        self.include_metadata = include_metadata
        self.min_similarity = min_similarity

    @classmethod
    def fast(cls, k):
        return cls(top_k=k, ef=k * 2)  # search.rs:31-37

    @classmethod
    def accurate(cls, k):
        return cls(top_k=k, ef=k * 10)  # search.rs:40-46


class SearchResult:
    """SearchResult (search.rs:50-103): score is the distance (lower is better)."""

    def __init__(self, id, score, vector=None, metadata=None, text=None):
        self.id = int(id)
        self.score = np.float32(score)
        self.vector = vector
        self.metadata = metadata
        self.text = text

    def with_vector(self, vector):
        self.vector = vector
        return self

    def with_metadata(self, metadata):
        self.metadata = metadata
        return self

    def with_text(self, text):
        self.text = str(text)
        return self

    def to_similarity(self):
        return np.float32(1.0) / (np.float32(1.0) + self.score)  # search.rs:99-102

    def __repr__(self):
        return f"SearchResult(id={self.id}, score={float(self.score):.6f})"


def _get_vector(graph, id):
    """get_node(id).vector of an HnswGraph, or the resident embedding of a LeannIndex."""
    lib = _ffi.load()
    dim = graph.dimension()
    out = np.empty(dim, np.float32)
    fn = lib.isl_hnsw_get_vector if hasattr(graph, "export_layer") else lib.isl_index_get_vector
    _check(fn(graph._h, int(id), _ptr(out, f32p)))
    return out


def _graph_search(graph, queries, k, ef):
    """(ids, dist, count) of HnswGraph::search / LeannIndex::search_with_params for a batch."""
    return graph.search_batch(queries, k, ef)


class Searcher:
    """Searcher (search.rs:106-182) over an HnswGraph (or LeannIndex) handle."""

    def __init__(self, graph, config=None):
        self.graph = graph
        self.config = config or SearchConfig()

    @classmethod
    def with_config(cls, graph, config):
        return cls(graph, config)

    def top_k(self, k):
        self.config.top_k = k
        return self

    def ef(self, ef):
        self.config.ef = ef
        return self

    def include_vectors(self):
        self.config.include_vectors = True
        return self

    def min_similarity(self, threshold):
        self.config.min_similarity = threshold
        return self

    def _shape(self, ids, dist, cnt):
        out = []
        for i in range(int(cnt)):
            r = SearchResult(ids[i], dist[i])
            if self.config.include_vectors:  # search.rs:160-164
                r = r.with_vector(_get_vector(self.graph, ids[i]))
            out.append(r)
        if self.config.min_similarity is not None:  # search.rs:171-173
            ms = np.float32(self.config.min_similarity)
            out = [r for r in out if r.to_similarity() >= ms]
        return out

    def search(self, query):
        return self.search_batch(_f32(query).reshape(1, -1))[0]

    def search_batch(self, queries):
        """search.rs:179-181 maps search over the queries; here the whole batch is one GPU call."""
        q = _f32(queries)
        if q.size == 0:
            return []
        ids, dist, cnt = _graph_search(self.graph, q, self.config.top_k, self.config.ef)
        return [self._shape(ids[i], dist[i], cnt[i]) for i in range(q.shape[0])]


class MultiIndexSearcher:
    """MultiIndexSearcher (search.rs:185-249): search every island, merge by score ascending with a
    stable sort (ties keep island order, then rank), truncate to top_k."""

    def __init__(self):
        self.graphs = []
        self.config = SearchConfig()

    def add_index(self, name, graph):
        self.graphs.append((str(name), graph))

    def with_config(self, config):
        self.config = config
        return self

    def search(self, query):
        return self.search_batch(_f32(query).reshape(1, -1))[0]

    def search_batch(self, queries):
        q = _f32(queries)
        k = self.config.top_k
        if not self.graphs or q.size == 0:
            return [[] for _ in range(q.shape[0] if q.ndim == 2 else 0)]
        nq = q.shape[0]
        all_ids = np.full((len(self.graphs), nq, k), _ffi.ISL_INVALID_ID, np.uint64)
        all_dst = np.full((len(self.graphs), nq, k), np.inf, np.float32)
        for g, (_, graph) in enumerate(self.graphs):
            if len(graph) == 0:
                continue
            ids, dist, cnt = _graph_search(graph, q, k, self.config.ef)
            # island index in the high bits: the (dist, id) merge rule then equals the reference's stable
            # sort over the island-ordered concatenation (search.rs:231)
            valid = np.arange(k)[None, :] < cnt[:, None]
            all_ids[g] = np.where(valid, (np.uint64(g) << np.uint64(40)) | ids, np.uint64(_ffi.ISL_INVALID_ID))
            all_dst[g] = np.where(valid, dist, np.float32(np.inf))
        m_ids, m_dst, m_cnt = merge_topk(all_ids, all_dst, k)  # isl_merge_topk (GPU)
        out = []
        for i in range(nq):
            row = []
            for j in range(int(m_cnt[i])):
                g = int(m_ids[i, j] >> np.uint64(40))
                nid = int(m_ids[i, j] & np.uint64((1 << 40) - 1))
                r = SearchResult(nid, m_dst[i, j])
                if self.config.include_vectors:
                    r = r.with_vector(_get_vector(self.graphs[g][1], nid))
                row.append((self.graphs[g][0], r))
            out.append(row)
        return out

    def num_indexes(self):
        return len(self.graphs)

    def total_vectors(self):
        return sum(len(g) for _, g in self.graphs)
