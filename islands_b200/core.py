"""Host-side mirror of the reference's `src/core` API for the search hot path.

Same names, argument meaning and error behaviour as the Rust (paths relative to the reference):
  DistanceMetric / Distance           src/core/distance.rs:7-67
  CoreError variants                  src/core/error.rs:9-62
  LeannConfig, LeannIndex, CsrGraph   src/core/leann.rs:193-302, 322-461, 493-1067
  InMemoryEmbeddingProvider           src/core/leann.rs:104-159
  PQConfig, ProductQuantizer          src/core/pq.rs:13-65, 116-359
  MultiIndexSearcher-style merge      src/core/search.rs:211-237

Everything here is argument marshalling over the C ABI (include/islands_b200.h); all compute
runs in libislands_b200.so on the GPU.  numpy arrays in, numpy arrays out.
"""
import ctypes as C
import math

import numpy as np

from . import _ffi
from ._ffi import (EncoderConfigStruct, HnswConfigStruct, LeannConfigStruct, PQConfigStruct, SearchStatsStruct,
                   f32p, u16p, u32p, u64p)


# ---- errors (src/core/error.rs:9-62) ---------------------------------------------------------
class CoreError(Exception):
    status = -1


class DimensionMismatch(CoreError):
    status = 1


class EmptyCollection(CoreError):
    status = 2


class InvalidConfig(CoreError):
    status = 3


class IndexNotBuilt(CoreError):
    status = 4


class NodeNotFound(CoreError):
    status = 5


class PQError(CoreError):
    status = 6


class SerializationError(CoreError):
    status = 7


class CudaError(CoreError):
    status = 8


class InvalidArgument(CoreError):
    status = 9


_ERRORS = {c.status: c for c in (DimensionMismatch, EmptyCollection, InvalidConfig, IndexNotBuilt,
                                 NodeNotFound, PQError, SerializationError, CudaError, InvalidArgument)}


def _check(status):
    if status != _ffi.ISL_OK:
        raise _ERRORS.get(status, CoreError)(_ffi.last_error())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


# ---- distance.rs ------------------------------------------------------------------------------
class DistanceMetric:
    """DistanceMetric (distance.rs:9-19) + the Distance trait methods (distance.rs:22-67)."""

    Cosine = 0
    Euclidean = 1
    DotProduct = 2
    Manhattan = 3
    _names = {0: "cosine", 1: "euclidean", 2: "dotproduct", 3: "manhattan"}

    def __init__(self, value=0):
        self.value = int(value)

    def __eq__(self, other):
        return int(getattr(other, "value", other)) == self.value

    def __hash__(self):
        return hash(self.value)

    def __repr__(self):
        return f"DistanceMetric.{self._names[self.value]}"

    def debug_name(self):
        """`{:?}` of the variant (distance.rs:9-19): Cosine, Euclidean, DotProduct, Manhattan."""
        return ("Cosine", "Euclidean", "DotProduct", "Manhattan")[self.value]

    def calculate(self, a, b):
        a, b = _f32(a), _f32(b)
        out = C.c_float()
        _check(_ffi.load().isl_distance_calculate(self.value, _ptr(a, f32p), a.size, _ptr(b, f32p), b.size,
                                                  C.byref(out)))
        return out.value

    def calculate_squared(self, a, b):
        a, b = _f32(a), _f32(b)
        out = C.c_float()
        _check(_ffi.load().isl_distance_calculate_squared(self.value, _ptr(a, f32p), a.size, _ptr(b, f32p),
                                                          b.size, C.byref(out)))
        return out.value

    def batch_calculate(self, query, vectors):
        query = _f32(query)
        vectors = _f32(vectors)
        if vectors.size == 0:
            return np.zeros(0, np.float32)
        vectors = vectors.reshape(-1, vectors.shape[-1])
        if vectors.shape[1] != query.size:
            raise DimensionMismatch(f"dimension mismatch: expected {query.size}, got {vectors.shape[1]}")
        out = np.empty(vectors.shape[0], np.float32)
        _check(_ffi.load().isl_distance_batch(self.value, _ptr(query, f32p), _ptr(vectors, f32p),
                                              vectors.shape[0], query.size, _ptr(out, f32p)))
        return out


def normalize_vector(v):
    """normalize_vector / normalized (distance.rs:125-139); returns a new array."""
    a = _f32(v).copy()
    rows = a.reshape(1, -1) if a.ndim == 1 else a
    _check(_ffi.load().isl_normalize_rows(_ptr(rows, f32p), rows.shape[0], rows.shape[1]))
    return a


normalized = normalize_vector


# ---- leann.rs ---------------------------------------------------------------------------------
class PruningStrategy:
    Global = 0
    Local = 1
    Proportional = 2


class LeannConfig:
    """LeannConfig (leann.rs:322-461)."""

    _fields = [f[0] for f in LeannConfigStruct._fields_]

    def __init__(self, **kw):
        s = LeannConfigStruct()
        _check(_ffi.load().isl_leann_config_default(C.byref(s)))
        self._s = s
        for k, v in kw.items():
            setattr(self, k, v)

    def __getattr__(self, name):
        if name in LeannConfig._fields:
            return getattr(self._s, name)
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in LeannConfig._fields:
            setattr(self._s, name, getattr(value, "value", value))
        else:
            object.__setattr__(self, name, value)

    @classmethod
    def paper_default(cls):
        return cls()

    @classmethod
    def default(cls):
        return cls()

    @classmethod
    def fast(cls):
        c = cls()
        _check(_ffi.load().isl_leann_config_fast(C.byref(c._s)))
        return c

    @classmethod
    def accurate(cls):
        c = cls()
        _check(_ffi.load().isl_leann_config_accurate(C.byref(c._s)))
        return c

    def validate(self):
        _check(_ffi.load().isl_leann_config_validate(C.byref(self._s)))


class HnswConfig:
    """HnswConfig (hnsw.rs:15-86)."""

    _fields = [f[0] for f in HnswConfigStruct._fields_]

    def __init__(self, **kw):
        s = HnswConfigStruct()
        _check(_ffi.load().isl_hnsw_config_default(C.byref(s)))
        self._s = s
        for k, v in kw.items():
            setattr(self._s, k, getattr(v, "value", v))

    def __getattr__(self, name):
        if name in HnswConfig._fields:
            return getattr(self._s, name)
        raise AttributeError(name)

    @classmethod
    def new(cls, _config=None):
        """HnswConfig::new (hnsw.rs:30-35) drops its argument and returns the defaults; kept."""
        return cls()

    @classmethod
    def fast(cls):
        return cls(m=12, m0=24, ef_construction=100)  # hnsw.rs:52-59

    @classmethod
    def accurate(cls):
        return cls(m=32, m0=64, ef_construction=400)  # hnsw.rs:62-69

    def validate(self):
        _check(_ffi.load().isl_hnsw_config_validate(C.byref(self._s)))


class HnswNode:
    """HnswNode (hnsw.rs:90-125): id, level, per-layer connections (vector stays on the device)."""

    def __init__(self, id, level, connections):
        self.id = id
        self.level = level
        self.connections = connections

    def neighbors_at(self, layer):
        return self.connections[layer] if layer < len(self.connections) else None


class HnswGraph:
    """HnswGraph (hnsw.rs:151-531) with vectors and all layers resident in HBM."""

    def __init__(self, config=None):
        self.config = config or HnswConfig()
        self.config.validate()  # HnswGraph::new (hnsw.rs:167-178)
        h = C.c_void_p()
        _check(_ffi.load().isl_hnsw_new(C.byref(self.config._s), C.byref(h)))
        self._h = h

    def free(self):
        if getattr(self, "_h", None) is not None:
            _ffi.load().isl_hnsw_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def __len__(self):
        return int(_ffi.load().isl_hnsw_len(self._h))

    def is_empty(self):
        return len(self) == 0

    def dimension(self):
        return int(_ffi.load().isl_hnsw_dimension(self._h)) or None

    @property
    def entry_point(self):
        e = int(_ffi.load().isl_hnsw_entry_point(self._h))
        return None if e < 0 else e

    @property
    def max_level(self):
        return int(_ffi.load().isl_hnsw_max_level(self._h))

    def insert(self, vector, level=None, seed=0):
        """HnswGraph::insert (hnsw.rs:214-250) -> id.  `level` replaces the thread_rng draw."""
        v = _f32(vector).reshape(1, -1)
        return self.insert_batch(v, None if level is None else [level], seed=seed, batch=1)

    def insert_batch(self, vectors, levels=None, seed=0, batch=1):
        """len(vectors) inserts; batch > 1 inserts up to `batch` nodes per graph snapshot. -> first id"""
        v = _f32(vectors)
        v = v.reshape(1, -1) if v.ndim == 1 else v
        lv = np.ascontiguousarray(levels, np.uint64) if levels is not None else None
        first = C.c_uint64()
        _check(_ffi.load().isl_hnsw_insert_batch(self._h, _ptr(v, f32p), v.shape[0], v.shape[1], _ptr(lv, u64p),
                                                 seed, batch, C.byref(first)))
        return first.value

    def insert_batch_dev(self, d_vectors_ptr, count, dim, levels=None, seed=0, batch=1):
        lv = np.ascontiguousarray(levels, np.uint64) if levels is not None else None
        first = C.c_uint64()
        _check(_ffi.load().isl_hnsw_insert_batch_dev(self._h, C.c_void_p(d_vectors_ptr), count, dim, _ptr(lv, u64p),
                                                     seed, batch, C.byref(first)))
        return first.value

    def get_node(self, id):
        """HnswGraph::get_node (hnsw.rs:201-203)."""
        lib = _ffi.load()
        lvl = C.c_uint64()
        if lib.isl_hnsw_node_level(self._h, id, C.byref(lvl)) != 0:
            return None
        conns = []
        cap = int(max(self.config.m0, self.config.m))
        for layer in range(lvl.value + 1):
            buf = np.empty(cap, np.uint64)
            cnt = C.c_uint64()
            _check(lib.isl_hnsw_get_neighbors(self._h, id, layer, _ptr(buf, u64p), cap, C.byref(cnt)))
            conns.append(buf[:cnt.value].copy())
        return HnswNode(id, lvl.value, conns)

    def export_layer(self, layer):
        """-> (degrees [n] int64, -1 where the node lacks the layer; neighbors [n, M] u64 padded)."""
        n = len(self)
        cc = int(self.config.m0 if layer == 0 else self.config.m)
        deg = np.empty(n, np.int64)
        nb = np.empty((n, cc), np.uint64)
        _check(_ffi.load().isl_hnsw_export_layer(self._h, layer, deg.ctypes.data_as(C.POINTER(C.c_int64)), _ptr(nb, u64p)))
        return deg, nb

    def search(self, query, k, ef):
        """HnswGraph::search (hnsw.rs:458-504): one query -> [(id, dist)] ascending."""
        ids, dist, cnt = self.search_batch(_f32(query).reshape(1, -1), k, ef)
        return [(int(ids[0, i]), float(dist[0, i])) for i in range(int(cnt[0]))]

    def search_batch(self, queries, k, ef):
        q = _f32(queries)
        q = q.reshape(1, -1) if q.ndim == 1 else q
        nq, qd = q.shape
        ids = np.empty((nq, k), np.uint64)
        dist = np.empty((nq, k), np.float32)
        cnt = np.empty(nq, np.uint32)
        _check(_ffi.load().isl_hnsw_search(self._h, _ptr(q, f32p), nq, qd, k, int(ef), _ptr(ids, u64p),
                                           _ptr(dist, f32p), _ptr(cnt, u32p)))
        return ids, dist, cnt

    def search_batch_dev(self, d_queries_ptr, nq, dim, k, ef, d_ids_ptr, d_dist_ptr, d_count_ptr=None):
        _check(_ffi.load().isl_hnsw_search_dev(self._h, C.c_void_p(d_queries_ptr), nq, dim, k, int(ef),
                                               C.c_void_p(d_ids_ptr), C.c_void_p(d_dist_ptr),
                                               C.c_void_p(d_count_ptr) if d_count_ptr else None))

    def last_search_timing(self):
        ms = C.c_float()
        _check(_ffi.load().isl_hnsw_last_search_timing(self._h, C.byref(ms)))
        return ms.value


class InMemoryEmbeddingProvider:
    """InMemoryEmbeddingProvider (leann.rs:104-159): id -> stored embedding."""

    def __init__(self, embeddings):
        e = _f32(embeddings)
        if e.size == 0:
            raise EmptyCollection("empty collection")
        self.embeddings = e.reshape(-1, e.shape[-1])

    @classmethod
    def with_dimension(cls, dimension):
        """An empty provider of a fixed dimension (leann.rs:136-141), filled with `add`."""
        self = cls.__new__(cls)
        self.embeddings = np.zeros((0, int(dimension)), np.float32)
        return self

    def add(self, embedding):
        """leann.rs:123-133: append one embedding, returns its id."""
        e = _f32(embedding).reshape(-1)
        if e.size != self.embeddings.shape[1]:
            raise DimensionMismatch(f"dimension mismatch: expected {self.embeddings.shape[1]}, got {e.size}")
        self.embeddings = np.concatenate([self.embeddings, e[None, :]])
        return self.embeddings.shape[0] - 1

    def dimension(self):
        return self.embeddings.shape[1]

    def compute_embedding(self, id):
        if id >= self.embeddings.shape[0]:
            raise NodeNotFound(f"node {id} not found")
        return self.embeddings[id].copy()

    def compute_embeddings_batch(self, ids):
        return np.stack([self.compute_embedding(i) for i in ids])


class CsrGraph:
    """CsrGraph (leann.rs:193-302): plain arrays in the reference's layout."""

    def __init__(self):
        self.node_offsets = np.zeros(1, np.uint64)
        self.neighbors = np.zeros(0, np.uint64)
        self.levels = np.zeros(0, np.uint64)
        self.entry_point = None
        self.max_level = 0
        self.num_nodes = 0
        self.degree_counts = np.zeros(0, np.uint64)

    def get_neighbors(self, node_id):
        if node_id >= self.num_nodes:
            return None
        s, e = int(self.node_offsets[node_id]), int(self.node_offsets[node_id + 1])
        return self.neighbors[s:e]

    def add_node(self, neighbors, level):
        nid = self.num_nodes
        self.num_nodes += 1
        self.levels = np.append(self.levels, np.uint64(level))
        self.degree_counts = np.append(self.degree_counts, np.uint64(len(neighbors)))
        self.neighbors = np.concatenate([self.neighbors, np.asarray(neighbors, np.uint64)])
        self.node_offsets = np.append(self.node_offsets, np.uint64(self.neighbors.size))
        if self.entry_point is None or level > self.max_level:
            self.entry_point = nid
            self.max_level = level
        return nid

    def set_neighbors(self, node_id, new_neighbors):
        """CsrGraph::set_neighbors (leann.rs:256-293): same length overwrites in place, any other length
        rebuilds offsets and neighbours; an unknown node is ignored."""
        if node_id >= self.num_nodes:
            return
        nb = np.asarray(new_neighbors, np.uint64)
        s, e = int(self.node_offsets[node_id]), int(self.node_offsets[node_id + 1])
        if nb.size == e - s:
            self.neighbors[s:e] = nb
        else:
            self.neighbors = np.concatenate([self.neighbors[:s], nb, self.neighbors[e:]])
            delta = nb.size - (e - s)
            off = self.node_offsets.astype(np.int64)
            off[node_id + 1:] += delta
            self.node_offsets = off.astype(np.uint64)
        self.degree_counts[node_id] = nb.size

    def storage_bytes(self):
        return 8 * (self.node_offsets.size + self.neighbors.size + self.levels.size + self.degree_counts.size)


class ShardComm:
    """isl_shard: one rank's membership in a sharded index (NCCL communicator inside the library)."""

    UNIQUE_ID_BYTES = 128

    @staticmethod
    def unique_id():
        """ncclGetUniqueId: call on one rank, hand the 128 bytes to the others by any host channel."""
        buf = (C.c_uint8 * ShardComm.UNIQUE_ID_BYTES)()
        _check(_ffi.load().isl_shard_unique_id(buf, ShardComm.UNIQUE_ID_BYTES))
        return bytes(buf)

    def __init__(self, rank, world, unique_id):
        uid = (C.c_uint8 * ShardComm.UNIQUE_ID_BYTES).from_buffer_copy(bytes(unique_id))
        h = C.c_void_p()
        _check(_ffi.load().isl_shard_init(int(rank), int(world), uid, C.byref(h)))
        self._h = h
        self.rank, self.world = int(rank), int(world)

    def enable_peer_exchange(self, max_records):
        """Exchange by peer stores over NVLink instead of ncclAllGather (collective; ranks of one node)."""
        _check(_ffi.load().isl_shard_enable_peer_exchange(self._h, int(max_records)))

    def last_timing(self):
        """(search_ms, exchange_ms, merge_ms) of the last sharded search on this rank (CUDA events)."""
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        _check(_ffi.load().isl_shard_last_timing(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def free(self):
        if getattr(self, "_h", None) is not None:
            _ffi.load().isl_shard_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def merge_packed_dev(d_records_ptr, parts, nq, k, d_ids_ptr, d_dist_ptr, d_count_ptr=None):
    """Merge of `parts` record lists [parts][nq][k] (isl_shard_record) on the device."""
    _check(_ffi.load().isl_merge_packed_dev(C.c_void_p(d_records_ptr), parts, nq, k, C.c_void_p(d_ids_ptr),
                                            C.c_void_p(d_dist_ptr), C.c_void_p(d_count_ptr) if d_count_ptr else None))


def set_caller_stream(cuda_stream_ptr):
    """isl_set_caller_stream: the CUDA stream (raw cudaStream_t value, 0 / None = legacy default stream) on
    which this thread produces the device buffers it hands to `_dev` entry points."""
    _check(_ffi.load().isl_set_caller_stream(C.c_void_p(cuda_stream_ptr or 0)))


class SearchStats:
    def __init__(self, arr):
        self.n_hop = arr["n_hop"]
        self.n_edge = arr["n_edge"]
        self.n_dist = arr["n_dist"]
        self.n_adc = arr["n_adc"]
        self.n_rerank = arr["n_rerank"]


_STATS_DTYPE = np.dtype([("n_hop", "<u8"), ("n_edge", "<u8"), ("n_dist", "<u8"), ("n_adc", "<u8"), ("n_rerank", "<u8")])


class LeannIndex:
    """LeannIndex (leann.rs:493-1067) with the graph and the provider's embeddings resident in HBM."""

    def __init__(self, config=None):
        self.config = config or LeannConfig()
        self.config.validate()  # LeannIndex::new (leann.rs:504-511)
        self._h = None
        self._pq = None

    # -- construction -------------------------------------------------------------------------
    @classmethod
    def from_csr(cls, config, vectors, node_offsets, neighbors, levels=None, entry_point=None):
        self = cls(config)
        v = _f32(vectors)
        n = v.shape[0] if v.ndim == 2 else 0
        dim = v.shape[1] if v.ndim == 2 else 0
        off = np.ascontiguousarray(node_offsets, np.uint64)
        nb = np.ascontiguousarray(neighbors, np.uint64)
        lv = np.ascontiguousarray(levels, np.uint64) if levels is not None else None
        h = C.c_void_p()
        _check(_ffi.load().isl_index_from_csr(C.byref(self.config._s), dim, n, _ptr(off, u64p), _ptr(nb, u64p),
                                              _ptr(lv, u64p), -1 if entry_point is None else int(entry_point),
                                              _ptr(v, f32p), C.byref(h)))
        self._h = h
        return self

    def build(self, provider, num_vectors, levels=None, seed=0, batch=1):
        """LeannIndex::build (leann.rs:560-631).  `provider` is an InMemoryEmbeddingProvider (or an
        [n, d] array); `levels` replaces the reference's thread_rng draw."""
        self.free()
        emb = provider.embeddings if hasattr(provider, "embeddings") else _f32(provider)
        if num_vectors == 0:
            h = C.c_void_p()
            _check(_ffi.load().isl_index_from_csr(C.byref(self.config._s), 0, 0, None, None, None, -1, None,
                                                  C.byref(h)))
            self._h = h
            return
        if num_vectors > emb.shape[0]:
            raise NodeNotFound(f"node {emb.shape[0]} not found")
        v = _f32(emb[:num_vectors])
        lv = np.ascontiguousarray(levels, np.uint64) if levels is not None else None
        h = C.c_void_p()
        _check(_ffi.load().isl_index_build(C.byref(self.config._s), v.shape[1], num_vectors, _ptr(v, f32p),
                                           _ptr(lv, u64p), seed, batch, C.byref(h)))
        self._h = h

    def build_dev(self, d_vectors_ptr, num_vectors, dim, levels=None, seed=0, batch=1):
        """LeannIndex::build from embeddings already resident on the current CUDA device
        (row-major [num_vectors, dim] f32 at the raw device address `d_vectors_ptr`)."""
        self.free()
        lv = np.ascontiguousarray(levels, np.uint64) if levels is not None else None
        h = C.c_void_p()
        _check(_ffi.load().isl_index_build_dev(C.byref(self.config._s), dim, num_vectors, C.c_void_p(d_vectors_ptr),
                                               _ptr(lv, u64p), seed, batch, C.byref(h)))
        self._h = h

    def search_batch_dev(self, d_queries_ptr, nq, dim, k, ef, d_ids_ptr, d_dist_ptr, d_count_ptr=None,
                         d_stats_ptr=None):
        """Batched search with queries and outputs on the device (raw addresses): no host copies."""
        _check(_ffi.load().isl_index_search_dev(self._h, C.c_void_p(d_queries_ptr), nq, dim, k, int(ef),
                                                C.c_void_p(d_ids_ptr), C.c_void_p(d_dist_ptr),
                                                C.c_void_p(d_count_ptr) if d_count_ptr else None,
                                                C.c_void_p(d_stats_ptr) if d_stats_ptr else None))

    def last_build_stats(self):
        """isl_index_last_build_stats: traversal counters / edges / CUDA-event times of the construction."""
        st = _ffi.BuildStatsStruct()
        _check(_ffi.load().isl_index_last_build_stats(self._h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in st._fields_}

    def set_neighbors(self, node_id, new_neighbors):
        """graph.set_neighbors (leann.rs:256-293) on the resident graph (device copy refreshed)."""
        nb = np.ascontiguousarray(new_neighbors, np.uint64)
        _check(_ffi.load().isl_index_set_neighbors(self._h, int(node_id), _ptr(nb, u64p) if nb.size else None, nb.size))

    # -- sharded search (service.rs:777-801, search.rs:211-237; include/islands_b200.h "sharded search") --
    def search_sharded(self, shard, id_base, queries, k, ef=None):
        """Collective over the ranks of `shard` (a ShardComm): every rank passes the same queries, searches
        its own shard and receives the same merged top-k with global ids (local id + id_base)."""
        q = _f32(queries)
        q = q.reshape(1, -1) if q.ndim == 1 else q
        nq, qd = q.shape
        ef = int(self.config.ef_search) if ef is None else int(ef)
        ids = np.empty((nq, k), np.uint64)
        dist = np.empty((nq, k), np.float32)
        cnt = np.empty(nq, np.uint32)
        _check(_ffi.load().isl_index_search_sharded(self._h, shard._h, int(id_base), _ptr(q, f32p), nq, qd, k, ef,
                                                    _ptr(ids, u64p), _ptr(dist, f32p), _ptr(cnt, u32p)))
        return ids, dist, cnt

    def search_sharded_adc(self, shard, id_base, queries, k, ef, recompute=False):
        """Sharded "PQ ADC traversal + exact rerank" (recompute=False) or its recompute form (True): collective."""
        q = _f32(queries)
        q = q.reshape(1, -1) if q.ndim == 1 else q
        nq, qd = q.shape
        ids = np.empty((nq, k), np.uint64)
        dist = np.empty((nq, k), np.float32)
        cnt = np.empty(nq, np.uint32)
        _check(_ffi.load().isl_index_search_sharded_adc(self._h, shard._h, int(id_base), 2 if recompute else 1, _ptr(q, f32p), nq,
                                                        qd, k, int(ef), _ptr(ids, u64p), _ptr(dist, f32p), _ptr(cnt, u32p)))
        return ids, dist, cnt

    def search_sharded_dev(self, shard, id_base, d_queries_ptr, nq, dim, k, ef, d_ids_ptr, d_dist_ptr, d_count_ptr=None):
        """The same with queries and outputs on the device (raw addresses)."""
        _check(_ffi.load().isl_index_search_sharded_dev(self._h, shard._h, int(id_base), C.c_void_p(d_queries_ptr), nq, dim,
                                                        k, int(ef), C.c_void_p(d_ids_ptr), C.c_void_p(d_dist_ptr),
                                                        C.c_void_p(d_count_ptr) if d_count_ptr else None))

    def search_packed_dev(self, id_base, d_queries_ptr, nq, dim, k, ef, d_records_ptr):
        """This shard's half of a sharded search: [nq][k] isl_shard_record on the device, no exchange."""
        _check(_ffi.load().isl_index_search_packed_dev(self._h, int(id_base), C.c_void_p(d_queries_ptr), nq, dim, k,
                                                       int(ef), C.c_void_p(d_records_ptr)))

    def free(self):
        if self._h is not None:
            _ffi.load().isl_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # -- accessors ----------------------------------------------------------------------------
    def __len__(self):
        return int(_ffi.load().isl_index_len(self._h)) if self._h else 0

    def is_empty(self):
        return len(self) == 0

    def dimension(self):
        d = int(_ffi.load().isl_index_dimension(self._h)) if self._h else 0
        return d or None

    def storage_bytes(self):
        return int(_ffi.load().isl_index_storage_bytes(self._h)) if self._h else 8

    def is_recompute(self):
        return bool(self.config.is_recompute)

    def is_compact(self):
        return bool(self.config.is_compact)

    @property
    def graph(self):
        lib = _ffi.load()
        g = CsrGraph()
        if not self._h:
            return g
        n = len(self)
        e = int(lib.isl_index_num_edges(self._h))
        g.num_nodes = n
        g.node_offsets = np.zeros(n + 1, np.uint64)
        g.neighbors = np.zeros(e, np.uint64)
        g.levels = np.zeros(n, np.uint64)
        g.degree_counts = np.zeros(n, np.uint64)
        _check(lib.isl_index_export_csr(self._h, _ptr(g.node_offsets, u64p), _ptr(g.neighbors, u64p) if e else None,
                                        _ptr(g.levels, u64p) if n else None,
                                        _ptr(g.degree_counts, u64p) if n else None))
        ep = int(lib.isl_index_entry_point(self._h))
        g.entry_point = None if ep < 0 else ep
        g.max_level = int(lib.isl_index_max_level(self._h))
        return g

    # -- search -------------------------------------------------------------------------------
    def search(self, query, k, provider=None):
        """LeannIndex::search (leann.rs:858-865): one query -> [(id, dist)] sorted ascending."""
        return self.search_with_params(query, k, int(self.config.ef_search), provider)

    def search_with_params(self, query, k, ef, provider=None):
        """LeannIndex::search_with_params (leann.rs:868-896)."""
        ids, dist, cnt = self.search_batch(_f32(query).reshape(1, -1), k, ef)
        return [(int(ids[0, i]), float(dist[0, i])) for i in range(int(cnt[0]))]

    def search_batch(self, queries, k, ef=None, stats=False):
        """Batched form: queries [nq, d] -> (ids [nq,k] u64, dist [nq,k] f32, count [nq] u32[, stats])."""
        if self._h is None:
            raise IndexNotBuilt("index not built")
        q = _f32(queries)
        q = q.reshape(1, -1) if q.ndim == 1 else q
        nq, qd = q.shape
        ef = int(self.config.ef_search) if ef is None else int(ef)
        ids = np.empty((nq, k), np.uint64)
        dist = np.empty((nq, k), np.float32)
        cnt = np.empty(nq, np.uint32)
        st = np.zeros(nq, _STATS_DTYPE) if stats else None
        _check(_ffi.load().isl_index_search(self._h, _ptr(q, f32p), nq, qd, k, ef, _ptr(ids, u64p),
                                            _ptr(dist, f32p), _ptr(cnt, u32p),
                                            st.ctypes.data_as(C.POINTER(SearchStatsStruct)) if stats else None))
        return (ids, dist, cnt, SearchStats(st)) if stats else (ids, dist, cnt)

    def last_search_timing(self):
        ms = C.c_float()
        n = C.c_uint64()
        _check(_ffi.load().isl_index_last_search_timing(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # -- two-level search (docs/leann-specification.md:223-269) -----------------------------------
    def attach_pq(self, pq, codes):
        codes = np.ascontiguousarray(codes, np.uint16)
        _check(_ffi.load().isl_index_attach_pq(self._h, pq._h, _ptr(codes, u16p)))
        self._pq = pq  # keep the quantizer alive while attached

    def search_two_level_batch(self, queries, k, ef, rerank_ratio, stats=False):
        q = _f32(queries)
        q = q.reshape(1, -1) if q.ndim == 1 else q
        nq, qd = q.shape
        ids = np.empty((nq, k), np.uint64)
        dist = np.empty((nq, k), np.float32)
        cnt = np.empty(nq, np.uint32)
        st = np.zeros(nq, _STATS_DTYPE) if stats else None
        _check(_ffi.load().isl_index_search_two_level(self._h, _ptr(q, f32p), nq, qd, k, int(ef),
                                                      float(rerank_ratio), _ptr(ids, u64p), _ptr(dist, f32p),
                                                      _ptr(cnt, u32p),
                                                      st.ctypes.data_as(C.POINTER(SearchStatsStruct)) if stats else None))
        return (ids, dist, cnt, SearchStats(st)) if stats else (ids, dist, cnt)


def _adc_rerank(self, queries, k, ef, stats=False):
    """PQ ADC traversal + exact rerank of the ef survivors (isl_index_search_adc_rerank)."""
    q = _f32(queries)
    q = q.reshape(1, -1) if q.ndim == 1 else q
    nq, qd = q.shape
    ids = np.empty((nq, k), np.uint64)
    dist = np.empty((nq, k), np.float32)
    cnt = np.empty(nq, np.uint32)
    st = np.zeros(nq, _STATS_DTYPE) if stats else None
    _check(_ffi.load().isl_index_search_adc_rerank(self._h, _ptr(q, f32p), nq, qd, k, int(ef), _ptr(ids, u64p),
                                                   _ptr(dist, f32p), _ptr(cnt, u32p),
                                                   st.ctypes.data_as(C.POINTER(SearchStatsStruct)) if stats else None))
    return (ids, dist, cnt, SearchStats(st)) if stats else (ids, dist, cnt)


LeannIndex.search_adc_rerank_batch = _adc_rerank


def _set_recompute(self, encoder, token_ids, lengths):
    """EmbeddingProvider analogue (leann.rs:82-99): node i -> encoder(token_ids[i], lengths[i])."""
    if encoder is None:
        _check(_ffi.load().isl_index_set_recompute(self._h, None, None, None, 1))
        self._enc = None
        return
    t = np.ascontiguousarray(token_ids, np.int32)
    ln = np.ascontiguousarray(lengths, np.int32)
    _check(_ffi.load().isl_index_set_recompute(self._h, encoder._h, t.ctypes.data_as(_ffi.i32p),
                                               ln.ctypes.data_as(_ffi.i32p), t.shape[1]))
    self._enc = encoder  # keep the encoder alive while attached


def _drop_vectors(self):
    _check(_ffi.load().isl_index_drop_vectors(self._h))


def _adc_recompute(self, queries, k, ef, stats=False):
    """ADC traversal + bf16 recompute of the survivors + exact rerank (isl_index_search_adc_recompute)."""
    q = _f32(queries)
    q = q.reshape(1, -1) if q.ndim == 1 else q
    nq, qd = q.shape
    ids = np.empty((nq, k), np.uint64)
    dist = np.empty((nq, k), np.float32)
    cnt = np.empty(nq, np.uint32)
    st = np.zeros(nq, _STATS_DTYPE) if stats else None
    _check(_ffi.load().isl_index_search_adc_recompute(self._h, _ptr(q, f32p), nq, qd, k, int(ef), _ptr(ids, u64p),
                                                      _ptr(dist, f32p), _ptr(cnt, u32p),
                                                      st.ctypes.data_as(C.POINTER(SearchStatsStruct)) if stats else None))
    return (ids, dist, cnt, SearchStats(st)) if stats else (ids, dist, cnt)


def _recompute_exact(self, queries, k, ef, stats=False):
    """LeannIndex::search_with_params with the encoder as the provider, hop by hop (isl_index_search_recompute;
    leann.rs:899-988 with compute_embeddings_batch on every hop's frontier, :947-950)."""
    q = _f32(queries)
    q = q.reshape(1, -1) if q.ndim == 1 else q
    nq, qd = q.shape
    ids = np.empty((nq, k), np.uint64)
    dist = np.empty((nq, k), np.float32)
    cnt = np.empty(nq, np.uint32)
    st = np.zeros(nq, _STATS_DTYPE) if stats else None
    _check(_ffi.load().isl_index_search_recompute(self._h, _ptr(q, f32p), nq, qd, k, int(ef), _ptr(ids, u64p),
                                                  _ptr(dist, f32p), _ptr(cnt, u32p),
                                                  st.ctypes.data_as(C.POINTER(SearchStatsStruct)) if stats else None))
    return (ids, dist, cnt, SearchStats(st)) if stats else (ids, dist, cnt)


def _last_recompute(self):
    u = C.c_uint64()
    a, b, c = C.c_float(), C.c_float(), C.c_float()
    _check(_ffi.load().isl_index_last_recompute(self._h, C.byref(u), C.byref(a), C.byref(b), C.byref(c)))
    h, hits = C.c_uint64(), C.c_uint64()
    _check(_ffi.load().isl_index_hub_cache_info(self._h, C.byref(h), C.byref(hits)))
    return dict(unique_nodes=u.value, traverse_ms=a.value, encoder_ms=b.value, rerank_ms=c.value,
                hub_cache_nodes=h.value, hub_cache_hits=hits.value)


def _set_hub_cache(self, count):
    """`HubCache` of docs/leann-specification.md:661-690: keep the embeddings of the `count` highest
    in-degree nodes resident so that the recompute search does not run them through the encoder
    (isl_index_set_hub_cache).  0 drops the cache."""
    _check(_ffi.load().isl_index_set_hub_cache(self._h, int(count)))


def _bytes_out(fn, handle):
    n = C.c_uint64()
    _check(fn(handle, None, 0, C.byref(n)))
    buf = (C.c_uint8 * max(1, n.value))()
    _check(fn(handle, C.cast(buf, C.c_void_p), n.value, C.byref(n)))
    return bytes(bytearray(buf)[:n.value])


def _index_to_bytes(self):
    """LeannIndex::to_bytes (leann.rs:1059-1061): graph structure only, bincode layout."""
    return _bytes_out(_ffi.load().isl_index_to_bytes, self._h)


def _index_from_bytes(cls, data, vectors):
    """LeannIndex::from_bytes (leann.rs:1064-1066) + the embeddings the provider would return."""
    v = _f32(vectors)
    buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data if len(data) else b"\0")
    h = C.c_void_p()
    cfg = LeannConfig()
    _check(_ffi.load().isl_index_from_bytes(C.cast(buf, C.c_void_p), len(data), _ptr(v, f32p) if v.size else None,
                                            v.shape[1] if v.ndim == 2 else 0, C.byref(h)))
    _check(_ffi.load().isl_index_get_config(h, C.byref(cfg._s)))
    self = cls.__new__(cls)
    self.config = cfg
    self._h = h
    self._pq = None
    return self


LeannIndex.to_bytes = _index_to_bytes
LeannIndex.from_bytes = classmethod(_index_from_bytes)


def _pq_to_bytes(self):
    """ProductQuantizer::to_bytes (pq.rs:351-353)."""
    return _bytes_out(_ffi.load().isl_pq_to_bytes, self._h)


def _pq_from_bytes(cls, data):
    """ProductQuantizer::from_bytes (pq.rs:356-358)."""
    buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data if len(data) else b"\0")
    h = C.c_void_p()
    _check(_ffi.load().isl_pq_from_bytes(C.cast(buf, C.c_void_p), len(data), C.byref(h)))
    self = cls.__new__(cls)
    self._h = h
    self.config = PQConfig()
    _check(_ffi.load().isl_pq_get_config(h, C.byref(self.config._s)))
    self.dimension = int(_ffi.load().isl_pq_dimension(h))
    return self


def _hnsw_to_bytes(self):
    """HnswGraph::to_bytes (hnsw.rs:507-509)."""
    return _bytes_out(_ffi.load().isl_hnsw_to_bytes, self._h)


def _hnsw_from_bytes(cls, data):
    """HnswGraph::from_bytes (hnsw.rs:512-514)."""
    buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data if len(data) else b"\0")
    h = C.c_void_p()
    _check(_ffi.load().isl_hnsw_from_bytes(C.cast(buf, C.c_void_p), len(data), C.byref(h)))
    self = cls.__new__(cls)
    self.config = HnswConfig()
    _check(_ffi.load().isl_hnsw_get_config(h, C.byref(self.config._s)))
    self._h = h
    return self


HnswGraph.to_bytes = _hnsw_to_bytes
HnswGraph.from_bytes = classmethod(_hnsw_from_bytes)

LeannIndex.set_recompute = _set_recompute
LeannIndex.drop_vectors = _drop_vectors
LeannIndex.search_adc_recompute_batch = _adc_recompute
LeannIndex.search_recompute_batch = _recompute_exact
LeannIndex.last_recompute = _last_recompute
LeannIndex.set_hub_cache = _set_hub_cache


def _set_rerank_limit(self, limit):
    """ADC traversal + exact rerank / recompute: only the `limit` survivors with the best table distance
    get an exact distance (isl_index_set_rerank_limit); 0 = all ef survivors."""
    _check(_ffi.load().isl_index_set_rerank_limit(self._h, int(limit)))


LeannIndex.set_rerank_limit = _set_rerank_limit


def random_level(u, ml, max_layers):
    """LeannIndex::random_level (leann.rs:549-554) for an explicit uniform draw u in (0,1)."""
    lvl = math.floor(-math.log(u) * ml)
    return min(max(lvl, 0), max_layers - 1)


# ---- pq.rs --------------------------------------------------------------------------------------
class PQConfig:
    """PQConfig (pq.rs:13-65)."""

    def __init__(self, num_subquantizers=8, num_centroids=256, training_iterations=25, seed=None):
        self._s = PQConfigStruct(num_subquantizers, num_centroids, training_iterations, 0 if seed is None else int(seed),
                                 0 if seed is None else 1)

    num_subquantizers = property(lambda self: self._s.num_subquantizers)
    num_centroids = property(lambda self: self._s.num_centroids)
    training_iterations = property(lambda self: self._s.training_iterations)
    seed = property(lambda self: self._s.seed if self._s.has_seed else None)

    def validate(self, dimension):
        _check(_ffi.load().isl_pq_config_validate(C.byref(self._s), dimension))

    def bytes_per_vector(self):
        return int(_ffi.load().isl_pq_config_bytes_per_vector(C.byref(self._s)))


class PQCodebook:
    """PQCodebook (pq.rs:66-112): the centroids of one subquantizer.  `find_nearest` runs on the device through a
    one-subquantizer quantizer holding these centroids (strict `<` scan: the first of equal centroids wins)."""

    def __init__(self, subvector_dim):
        self.centroids = []
        self.subvector_dim = int(subvector_dim)

    def find_nearest(self, subvector, metric=DistanceMetric.Euclidean):
        v = _f32(subvector).reshape(-1)
        if v.size != self.subvector_dim:  # pq.rs:87-92, before any arithmetic
            raise DimensionMismatch(f"dimension mismatch: expected {self.subvector_dim}, got {v.size}")
        cb = _f32(self.centroids).reshape(1, len(self.centroids), self.subvector_dim)
        pq = ProductQuantizer(self.subvector_dim, PQConfig(1, len(self.centroids), 1, None)).with_metric(metric)
        pq.set_codebooks(cb)
        return int(pq.encode(v)[0])

    def get_centroid(self, idx):
        return _f32(self.centroids[idx]) if 0 <= idx < len(self.centroids) else None


class ProductQuantizer:
    """ProductQuantizer (pq.rs:116-359)."""

    def __init__(self, dimension, config=None):
        self.config = config or PQConfig()
        self.dimension = dimension
        h = C.c_void_p()
        _check(_ffi.load().isl_pq_new(dimension, C.byref(self.config._s), C.byref(h)))
        self._h = h

    def __del__(self):
        try:
            if self._h is not None:
                _ffi.load().isl_pq_free(self._h)
                self._h = None
        except Exception:
            pass

    def with_metric(self, metric):
        _check(_ffi.load().isl_pq_set_metric(self._h, getattr(metric, "value", metric)))
        return self

    def is_trained(self):
        return bool(_ffi.load().isl_pq_is_trained(self._h))

    def num_subquantizers(self):
        return int(_ffi.load().isl_pq_num_subquantizers(self._h))

    def compression_ratio(self):
        return float(_ffi.load().isl_pq_compression_ratio(self._h))

    def train(self, vectors):
        v = _f32(vectors)
        if v.size == 0:
            raise EmptyCollection("empty collection")  # pq.rs:176-178
        v = v.reshape(-1, v.shape[-1])
        _check(_ffi.load().isl_pq_train(self._h, _ptr(v, f32p), v.shape[0], v.shape[1]))

    def set_codebooks(self, codebooks):
        cb = _f32(codebooks)  # [m][ksub][dsub]
        _check(_ffi.load().isl_pq_set_codebooks(self._h, _ptr(cb, f32p), cb.shape[1]))

    def codebooks(self):
        k = C.c_uint64()
        _check(_ffi.load().isl_pq_get_codebooks(self._h, None, C.byref(k)))
        m = self.num_subquantizers()
        out = np.empty((m, k.value, self.dimension // m), np.float32)
        _check(_ffi.load().isl_pq_get_codebooks(self._h, _ptr(out, f32p), C.byref(k)))
        return out

    def codebook(self, subquantizer):
        """`codebooks[subquantizer]` (pq.rs:120-121) as a PQCodebook."""
        cb = PQCodebook(self.dimension // self.num_subquantizers())
        cb.centroids = list(self.codebooks()[subquantizer])
        return cb

    def encode(self, vector):
        v = _f32(vector)
        single = v.ndim == 1
        v2 = v.reshape(1, -1) if single else v
        out = np.empty((v2.shape[0], self.num_subquantizers()), np.uint16)
        _check(_ffi.load().isl_pq_encode(self._h, _ptr(v2, f32p), v2.shape[0], v2.shape[1], _ptr(out, u16p)))
        return out[0] if single else out

    def decode(self, codes):
        c = np.ascontiguousarray(codes, np.uint16)
        single = c.ndim == 1
        c2 = c.reshape(1, -1) if single else c
        out = np.empty((c2.shape[0], self.dimension), np.float32)
        _check(_ffi.load().isl_pq_decode(self._h, _ptr(c2, u16p), c2.shape[0], c2.shape[1], _ptr(out, f32p)))
        return out[0] if single else out

    def build_distance_tables(self, query):
        q = _f32(query)
        k = C.c_uint64(0)
        if self.is_trained():
            _check(_ffi.load().isl_pq_get_codebooks(self._h, None, C.byref(k)))
        out = np.empty((self.num_subquantizers(), k.value), np.float32)
        _check(_ffi.load().isl_pq_build_tables(self._h, _ptr(q, f32p), q.size, _ptr(out, f32p)))
        return out

    def table_distance(self, tables, codes):
        t = _f32(tables)
        c = np.ascontiguousarray(codes, np.uint16)
        single = c.ndim == 1
        c2 = c.reshape(1, -1) if single else c
        out = np.empty(c2.shape[0], np.float32)
        _check(_ffi.load().isl_pq_table_distance(self._h, _ptr(t, f32p), _ptr(c2, u16p), c2.shape[0], _ptr(out, f32p)))
        return float(out[0]) if single else out

    def asymmetric_distance(self, query, codes):
        q = _f32(query)
        c = np.ascontiguousarray(codes, np.uint16)
        single = c.ndim == 1
        c2 = c.reshape(1, -1) if single else c
        out = np.empty(c2.shape[0], np.float32)
        _check(_ffi.load().isl_pq_asymmetric_distance(self._h, _ptr(q, f32p), q.size, _ptr(c2, u16p), c2.shape[0],
                                                      _ptr(out, f32p)))
        return float(out[0]) if single else out


# ---- embedding/candle_provider.rs -----------------------------------------------------------------
class EncoderConfig:
    """Shape of the recompute encoder (BERT; default = BERT-base, 110M parameters)."""

    _fields = [f[0] for f in EncoderConfigStruct._fields_]

    def __init__(self, **kw):
        s = EncoderConfigStruct()
        _check(_ffi.load().isl_encoder_config_default(C.byref(s)))
        self._s = s
        for k, v in kw.items():
            if k not in EncoderConfig._fields:
                raise AttributeError(k)
            setattr(self._s, k, v)

    def __getattr__(self, name):
        if name in EncoderConfig._fields:
            return getattr(self._s, name)
        raise AttributeError(name)


class Encoder:
    """The model behind CandleEmbedder::embed_texts_raw (candle_provider.rs:353-507) for token ids:
    BERT forward on the tcgen05 tensor cores (bf16, f32 accumulate), masked mean pooling, L2 norm."""

    def __init__(self, config=None):
        self.config = config or EncoderConfig()
        h = C.c_void_p()
        _check(_ffi.load().isl_encoder_new(C.byref(self.config._s), C.byref(h)))
        self._h = h

    def free(self):
        if getattr(self, "_h", None) is not None:
            _ffi.load().isl_encoder_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def dimension(self):
        return int(_ffi.load().isl_encoder_dimension(self._h))

    def num_parameters(self):
        return int(_ffi.load().isl_encoder_num_parameters(self._h))

    def init_random(self, seed=46, stddev=0.02):
        _check(_ffi.load().isl_encoder_init_random(self._h, seed, stddev))
        return self

    def parameter_shapes(self):
        c = self.config
        H, I = c.hidden_size, c.intermediate_size
        shapes = {"embeddings.word_embeddings.weight": (c.vocab_size, H),
                  "embeddings.position_embeddings.weight": (c.max_position, H),
                  "embeddings.token_type_embeddings.weight": (c.type_vocab_size, H),
                  "embeddings.LayerNorm.weight": (H,), "embeddings.LayerNorm.bias": (H,)}
        for l in range(c.num_layers):
            p = f"encoder.layer.{l}."
            for n in ("query", "key", "value"):
                shapes[p + f"attention.self.{n}.weight"] = (H, H)
                shapes[p + f"attention.self.{n}.bias"] = (H,)
            shapes[p + "attention.output.dense.weight"] = (H, H)
            shapes[p + "attention.output.dense.bias"] = (H,)
            shapes[p + "attention.output.LayerNorm.weight"] = (H,)
            shapes[p + "attention.output.LayerNorm.bias"] = (H,)
            shapes[p + "intermediate.dense.weight"] = (I, H)
            shapes[p + "intermediate.dense.bias"] = (I,)
            shapes[p + "output.dense.weight"] = (H, I)
            shapes[p + "output.dense.bias"] = (H,)
            shapes[p + "output.LayerNorm.weight"] = (H,)
            shapes[p + "output.LayerNorm.bias"] = (H,)
        return shapes

    def get_parameter(self, name):
        shape = self.parameter_shapes()[name]
        out = np.empty(shape, np.float32)
        _check(_ffi.load().isl_encoder_get_parameter(self._h, name.encode(), _ptr(out, f32p), out.size))
        return out

    def set_parameter(self, name, value):
        v = _f32(value)
        _check(_ffi.load().isl_encoder_set_parameter(self._h, name.encode(), _ptr(v, f32p), v.size))

    def state_dict(self):
        return {n: self.get_parameter(n) for n in self.parameter_shapes()}

    def load_safetensors(self, path, prefix=""):
        """Load BERT weights from a .safetensors file (8-byte little-endian header length, JSON header
        with dtype / shape / data_offsets, raw tensor bytes) — the format CandleEmbedder reads
        (candle_provider.rs:230-300).  Tensor names are the Hugging Face ones, optionally under
        `prefix` (e.g. "bert.").  F32 / F16 / BF16 are accepted.  Returns the names that were loaded."""
        import json
        import struct

        with open(path, "rb") as f:
            raw = f.read()
        (hlen,) = struct.unpack("<Q", raw[:8])
        header = json.loads(raw[8:8 + hlen].decode("utf-8"))
        base = 8 + hlen
        shapes = self.parameter_shapes()
        loaded = []
        for name, shape in shapes.items():
            meta = header.get(prefix + name)
            if meta is None:
                continue
            lo, hi = meta["data_offsets"]
            buf = raw[base + lo:base + hi]
            dt = meta["dtype"]
            if dt == "F32":
                arr = np.frombuffer(buf, "<f4")
            elif dt == "F16":
                arr = np.frombuffer(buf, "<f2").astype(np.float32)
            elif dt == "BF16":
                arr = (np.frombuffer(buf, "<u2").astype(np.uint32) << 16).view(np.float32)
            else:
                raise SerializationError(f"unsupported safetensors dtype {dt} for {name}")
            if tuple(meta["shape"]) != tuple(shape):
                raise DimensionMismatch(f"dimension mismatch: expected {tuple(shape)}, got {tuple(meta['shape'])} for {name}")
            self.set_parameter(name, arr.reshape(shape))
            loaded.append(name)
        missing = [n for n in shapes if n not in loaded]
        if missing:
            raise SerializationError(f"safetensors file lacks {len(missing)} parameters, e.g. {missing[0]}")
        return loaded

    def embed(self, token_ids, lengths):
        """token_ids [B, S] int32 (0-padded), lengths [B] -> [B, hidden] f32."""
        t = np.ascontiguousarray(token_ids, np.int32)
        ln = np.ascontiguousarray(lengths, np.int32)
        B, S = t.shape
        out = np.empty((B, self.dimension()), np.float32)
        _check(_ffi.load().isl_encoder_embed(self._h, t.ctypes.data_as(_ffi.i32p), ln.ctypes.data_as(_ffi.i32p), B, S,
                                             _ptr(out, f32p)))
        return out

    def embed_texts_raw(self, tokenizer, texts):
        """CandleEmbedder::embed_texts_raw (candle_provider.rs:353-507): tokenise the batch
        (`tokenizer.encode_batch(texts, true)`), pad every row with zeros to the longest encoding,
        BERT forward, mean pooling weighted by the attention mask, L2 normalisation.  `tokenizer` is an
        islands_b200.tokenizer.BertWordPieceTokenizer.  Empty input -> empty output (:354-356)."""
        if len(texts) == 0:
            return np.zeros((0, self.dimension()), np.float32)
        ids, types, mask = tokenizer.encode_batch_padded(list(texts))
        lengths = mask.sum(axis=1).astype(np.int32)
        if types.any() or not np.array_equal(mask, (np.arange(mask.shape[1])[None, :] < lengths[:, None]).astype(mask.dtype)):
            raise InvalidArgument("embed_texts_raw: single sequences with right padding only")
        return self.embed(ids, lengths)

    def embed_dev(self, d_tokens_ptr, d_lengths_ptr, B, S, d_out_ptr):
        _check(_ffi.load().isl_encoder_embed_dev(self._h, C.c_void_p(d_tokens_ptr), C.c_void_p(d_lengths_ptr), B, S,
                                                 C.c_void_p(d_out_ptr)))

    def last_timing(self):
        ms = C.c_float()
        fl = C.c_double()
        _check(_ffi.load().isl_encoder_last_timing(self._h, C.byref(ms), C.byref(fl)))
        return ms.value, fl.value


def gemm_bf16_dev(d_a, d_w, m, n, k, d_bias=None, d_residual=None, gelu=False, d_out_bf16=None, d_out_f32=None):
    """isl_gemm_bf16_dev: raw device addresses; out = act(A W^T + bias) (+ residual)."""
    vp = lambda p: C.c_void_p(p) if p else None
    _check(_ffi.load().isl_gemm_bf16_dev(vp(d_a), vp(d_w), m, n, k, vp(d_bias), vp(d_residual), int(gelu), vp(d_out_bf16),
                                         vp(d_out_f32)))


# ---- search.rs ----------------------------------------------------------------------------------
def to_similarity(score):
    """SearchResult::to_similarity (search.rs:99-102)."""
    return np.float32(1.0) / (np.float32(1.0) + np.float32(score))


def merge_topk_dev(d_ids_ptr, d_dist_ptr, parts, nq, k, d_out_ids_ptr, d_out_dist_ptr, d_out_count_ptr=None):
    """Device-pointer form of merge_topk: [parts, nq, k] lists already gathered on this GPU."""
    _check(_ffi.load().isl_merge_topk_dev(C.c_void_p(d_ids_ptr), C.c_void_p(d_dist_ptr), parts, nq, k,
                                          C.c_void_p(d_out_ids_ptr), C.c_void_p(d_out_dist_ptr),
                                          C.c_void_p(d_out_count_ptr) if d_out_count_ptr else None))


def merge_topk(ids, dist, k):
    """Island / shard merge (search.rs:211-237) under the (dist,id) rule.
    ids, dist: [parts, nq, k] -> ([nq,k] ids, [nq,k] dist, [nq] count)."""
    ids = np.ascontiguousarray(ids, np.uint64)
    dist = _f32(dist)
    parts, nq, kk = ids.shape
    assert kk == k
    out_ids = np.empty((nq, k), np.uint64)
    out_dist = np.empty((nq, k), np.float32)
    cnt = np.empty(nq, np.uint32)
    _check(_ffi.load().isl_merge_topk(_ptr(ids, u64p), _ptr(dist, f32p), parts, nq, k, _ptr(out_ids, u64p),
                                      _ptr(out_dist, f32p), _ptr(cnt, u32p)))
    return out_ids, out_dist, cnt


ProductQuantizer.to_bytes = _pq_to_bytes
ProductQuantizer.from_bytes = classmethod(_pq_from_bytes)
