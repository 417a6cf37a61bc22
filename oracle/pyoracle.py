"""ctypes wrapper of the CPU oracle (oracle/libislands_oracle.so).  TEST INFRASTRUCTURE ONLY:
import this from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, never from the islands_b200 package."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libislands_oracle.so")

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u16p = C.POINTER(C.c_uint16)
f32p = C.POINTER(C.c_float)

STATS_DTYPE = np.dtype([("n_hop", "<u8"), ("n_edge", "<u8"), ("n_dist", "<u8"), ("n_adc", "<u8"), ("n_rerank", "<u8")])
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        l = C.CDLL(LIB_PATH)
        l.orc_distance.restype = C.c_float
        l.orc_distance.argtypes = [C.c_int32, f32p, f32p, C.c_uint64]
        l.orc_distance_squared.restype = C.c_float
        l.orc_distance_squared.argtypes = [C.c_int32, f32p, f32p, C.c_uint64]
        l.orc_distance_batch.restype = None
        l.orc_distance_batch.argtypes = [C.c_int32, f32p, f32p, C.c_uint64, C.c_uint32, f32p]
        l.orc_normalize.restype = None
        l.orc_normalize.argtypes = [f32p, C.c_uint64]
        l.orc_level_from_uniform.restype = C.c_uint64
        l.orc_level_from_uniform.argtypes = [C.c_double, C.c_double, C.c_uint64]
        l.orc_draw_levels.restype = None
        l.orc_draw_levels.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_uint64, u64p]
        l.orc_leann_search.restype = C.c_int32
        l.orc_leann_search.argtypes = [C.c_void_p, f32p, C.c_uint64, C.c_uint32, u64p, u64p, C.c_int64, f32p,
                                       C.c_uint64, C.c_uint32, C.c_uint32, u64p, f32p, u32p, C.c_void_p, C.c_int32]
        l.orc_leann_build.restype = C.c_int32
        l.orc_leann_build.argtypes = [C.c_void_p, f32p, C.c_uint64, C.c_uint32, u64p, u64p, u64p, u64p,
                                      C.POINTER(C.c_int64), u64p]
        l.orc_leann_build_batched.restype = C.c_int32
        l.orc_leann_build_batched.argtypes = [C.c_void_p, f32p, C.c_uint64, C.c_uint32, u64p, C.c_uint32, u64p, u64p,
                                              u64p, C.POINTER(C.c_int64), u64p, C.c_int32]
        l.orc_pq_encode.restype = None
        l.orc_pq_encode.argtypes = [C.c_int32, f32p, C.c_uint32, C.c_uint32, C.c_uint32, f32p, C.c_uint64, u16p]
        l.orc_pq_decode.restype = C.c_int32
        l.orc_pq_decode.argtypes = [f32p, C.c_uint32, C.c_uint32, C.c_uint32, u16p, C.c_uint64, f32p]
        l.orc_pq_build_tables.restype = None
        l.orc_pq_build_tables.argtypes = [f32p, C.c_uint32, C.c_uint32, C.c_uint32, f32p, f32p]
        l.orc_pq_table_distance.restype = None
        l.orc_pq_table_distance.argtypes = [f32p, C.c_uint32, C.c_uint32, u16p, C.c_uint64, f32p]
        l.orc_pq_asymmetric_distance.restype = None
        l.orc_pq_asymmetric_distance.argtypes = [f32p, C.c_uint32, C.c_uint32, C.c_uint32, f32p, u16p, C.c_uint64, f32p]
        l.orc_pq_train.restype = C.c_int32
        l.orc_pq_train.argtypes = [C.c_int32, f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                   C.c_uint64, f32p, u32p]
        l.orc_leann_search_two_level.restype = C.c_int32
        l.orc_leann_search_two_level.argtypes = [C.c_void_p, f32p, C.c_uint64, C.c_uint32, u64p, u64p, C.c_int64,
                                                 f32p, C.c_uint32, C.c_uint32, u16p, f32p, C.c_uint64, C.c_uint32,
                                                 C.c_uint32, C.c_float, u64p, f32p, u32p, C.c_void_p, C.c_int32]
        l.orc_leann_search_adc_rerank.restype = C.c_int32
        l.orc_leann_search_adc_rerank.argtypes = [C.c_void_p, f32p, C.c_uint64, C.c_uint32, u64p, u64p, C.c_int64,
                                                  f32p, C.c_uint32, C.c_uint32, u16p, f32p, C.c_uint64, C.c_uint32,
                                                  C.c_uint32, u64p, f32p, u32p, C.c_void_p, C.c_int32, C.c_uint32]
        l.orc_adc_table_round.restype = None
        l.orc_adc_table_round.argtypes = [f32p, C.c_uint64, f32p]
        l.orc_merge_topk.restype = None
        l.orc_merge_topk.argtypes = [u64p, f32p, C.c_uint32, C.c_uint64, C.c_uint32, u64p, f32p, u32p]
        l.orc_to_similarity.restype = C.c_float
        l.orc_to_similarity.argtypes = [C.c_float]
        l.orc_hnsw_new.restype = C.c_void_p
        l.orc_hnsw_new.argtypes = [C.c_void_p, C.c_uint32]
        l.orc_hnsw_free.restype = None
        l.orc_hnsw_free.argtypes = [C.c_void_p]
        l.orc_hnsw_insert.restype = C.c_int32
        l.orc_hnsw_insert.argtypes = [C.c_void_p, f32p, C.c_uint64, u64p]
        l.orc_hnsw_insert_batch.restype = C.c_int32
        l.orc_hnsw_insert_batch.argtypes = [C.c_void_p, f32p, C.c_uint64, u64p, C.c_uint32, C.c_int32]
        l.orc_hnsw_node_level.restype = C.c_int64
        l.orc_hnsw_node_level.argtypes = [C.c_void_p, C.c_uint64]
        l.orc_hnsw_len.restype = C.c_uint64
        l.orc_hnsw_len.argtypes = [C.c_void_p]
        l.orc_hnsw_entry_point.restype = C.c_int64
        l.orc_hnsw_entry_point.argtypes = [C.c_void_p]
        l.orc_hnsw_max_level.restype = C.c_uint64
        l.orc_hnsw_max_level.argtypes = [C.c_void_p]
        l.orc_hnsw_neighbors.restype = C.c_int64
        l.orc_hnsw_neighbors.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, u64p, C.c_uint64]
        l.orc_hnsw_search.restype = C.c_int32
        l.orc_hnsw_search.argtypes = [C.c_void_p, f32p, C.c_uint64, C.c_uint32, C.c_uint32, u64p, f32p, u32p, C.c_int32]
        _lib = l
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def distance(metric, a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_distance(metric, _p(a, f32p), _p(b, f32p), a.size))


def distance_squared(metric, a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_distance_squared(metric, _p(a, f32p), _p(b, f32p), a.size))


def distance_batch(metric, q, rows):
    q, rows = _f32(q), _f32(rows)
    out = np.empty(rows.shape[0], np.float32)
    lib().orc_distance_batch(metric, _p(q, f32p), _p(rows, f32p), rows.shape[0], rows.shape[1], _p(out, f32p))
    return out


def normalize(v):
    v = _f32(v).copy()
    lib().orc_normalize(_p(v, f32p), v.size)
    return v


def draw_levels(seed, n, ml, max_layers):
    out = np.empty(n, np.uint64)
    lib().orc_draw_levels(seed, n, ml, max_layers, _p(out, u64p))
    return out


def _cfgp(cfg):
    """cfg: islands_b200._ffi.LeannConfigStruct (layout shared through include/islands_b200.h)."""
    return C.cast(C.byref(cfg), C.c_void_p)


def leann_search(cfg, vectors, offsets, nbrs, entry, queries, k, ef, threads=1, stats=False):
    vectors, queries = _f32(vectors), _f32(queries)
    offsets = np.ascontiguousarray(offsets, np.uint64)
    nbrs = np.ascontiguousarray(nbrs, np.uint64)
    n, d = vectors.shape if vectors.ndim == 2 else (0, queries.shape[1])
    nq = queries.shape[0]
    ids = np.empty((nq, k), np.uint64)
    dist = np.empty((nq, k), np.float32)
    cnt = np.empty(nq, np.uint32)
    st = np.zeros(nq, STATS_DTYPE) if stats else None
    rc = lib().orc_leann_search(_cfgp(cfg), _p(vectors, f32p), n, d, _p(offsets, u64p), _p(nbrs, u64p),
                                -1 if entry is None else int(entry), _p(queries, f32p), nq, k, ef, _p(ids, u64p),
                                _p(dist, f32p), _p(cnt, u32p), st.ctypes.data if stats else None, threads)
    if rc != 0:
        raise RuntimeError(f"orc_leann_search status {rc}")
    return (ids, dist, cnt, st) if stats else (ids, dist, cnt)


def leann_build(cfg, vectors, levels, batch=1, threads=1):
    vectors = _f32(vectors)
    n, d = vectors.shape
    levels = np.ascontiguousarray(levels, np.uint64)
    offsets = np.zeros(n + 1, np.uint64)
    nbrs = np.zeros(max(1, n * int(cfg.m0)), np.uint64)
    ne = C.c_uint64()
    entry = C.c_int64()
    maxl = C.c_uint64()
    rc = lib().orc_leann_build_batched(_cfgp(cfg), _p(vectors, f32p), n, d, _p(levels, u64p), batch, _p(offsets, u64p),
                                       _p(nbrs, u64p), C.byref(ne), C.byref(entry), C.byref(maxl), threads)
    if rc != 0:
        raise RuntimeError(f"orc_leann_build status {rc}")
    return offsets, nbrs[: ne.value].copy(), (None if entry.value < 0 else entry.value), maxl.value


def pq_encode(metric, codebooks, vectors):
    cb, v = _f32(codebooks), _f32(vectors)
    m, ksub, dsub = cb.shape
    out = np.empty((v.shape[0], m), np.uint16)
    lib().orc_pq_encode(metric, _p(cb, f32p), m, ksub, dsub, _p(v, f32p), v.shape[0], _p(out, u16p))
    return out


def pq_decode(codebooks, codes):
    cb = _f32(codebooks)
    codes = np.ascontiguousarray(codes, np.uint16)
    m, ksub, dsub = cb.shape
    out = np.empty((codes.shape[0], m * dsub), np.float32)
    rc = lib().orc_pq_decode(_p(cb, f32p), m, ksub, dsub, _p(codes, u16p), codes.shape[0], _p(out, f32p))
    if rc != 0:
        raise RuntimeError(f"orc_pq_decode status {rc}")
    return out


def pq_build_tables(codebooks, query):
    cb, q = _f32(codebooks), _f32(query)
    m, ksub, dsub = cb.shape
    out = np.empty((m, ksub), np.float32)
    lib().orc_pq_build_tables(_p(cb, f32p), m, ksub, dsub, _p(q, f32p), _p(out, f32p))
    return out


def pq_table_distance(tables, codes):
    t = _f32(tables)
    codes = np.ascontiguousarray(codes, np.uint16)
    out = np.empty(codes.shape[0], np.float32)
    lib().orc_pq_table_distance(_p(t, f32p), t.shape[0], t.shape[1], _p(codes, u16p), codes.shape[0], _p(out, f32p))
    return out


def pq_asymmetric_distance(codebooks, query, codes):
    cb, q = _f32(codebooks), _f32(query)
    codes = np.ascontiguousarray(codes, np.uint16)
    m, ksub, dsub = cb.shape
    out = np.empty(codes.shape[0], np.float32)
    lib().orc_pq_asymmetric_distance(_p(cb, f32p), m, ksub, dsub, _p(q, f32p), _p(codes, u16p), codes.shape[0],
                                     _p(out, f32p))
    return out


def pq_train(metric, vectors, m, ksub, iterations, seed):
    v = _f32(vectors)
    n, d = v.shape
    k_eff = min(ksub, n)
    out = np.zeros((m, k_eff, d // m), np.float32)
    ko = C.c_uint32()
    rc = lib().orc_pq_train(metric, _p(v, f32p), n, d, m, ksub, iterations, seed, _p(out, f32p), C.byref(ko))
    if rc != 0:
        raise RuntimeError(f"orc_pq_train status {rc}")
    return out


def leann_search_two_level(cfg, vectors, offsets, nbrs, entry, codebooks, codes, queries, k, ef, rerank_ratio,
                           threads=1, stats=False):
    vectors, queries, cb = _f32(vectors), _f32(queries), _f32(codebooks)
    offsets = np.ascontiguousarray(offsets, np.uint64)
    nbrs = np.ascontiguousarray(nbrs, np.uint64)
    codes = np.ascontiguousarray(codes, np.uint16)
    n, d = vectors.shape
    nq = queries.shape[0]
    m, ksub, _ = cb.shape
    ids = np.empty((nq, k), np.uint64)
    dist = np.empty((nq, k), np.float32)
    cnt = np.empty(nq, np.uint32)
    st = np.zeros(nq, STATS_DTYPE) if stats else None
    rc = lib().orc_leann_search_two_level(_cfgp(cfg), _p(vectors, f32p), n, d, _p(offsets, u64p), _p(nbrs, u64p),
                                          -1 if entry is None else int(entry), _p(cb, f32p), m, ksub, _p(codes, u16p),
                                          _p(queries, f32p), nq, k, ef, rerank_ratio, _p(ids, u64p), _p(dist, f32p),
                                          _p(cnt, u32p), st.ctypes.data if stats else None, threads)
    if rc != 0:
        raise RuntimeError(f"orc_leann_search_two_level status {rc}")
    return (ids, dist, cnt, st) if stats else (ids, dist, cnt)


def leann_search_adc_rerank(cfg, vectors, offsets, nbrs, entry, codebooks, codes, queries, k, ef, threads=1, stats=False,
                            rerank_limit=0):
    vectors, queries, cb = _f32(vectors), _f32(queries), _f32(codebooks)
    offsets = np.ascontiguousarray(offsets, np.uint64)
    nbrs = np.ascontiguousarray(nbrs, np.uint64)
    codes = np.ascontiguousarray(codes, np.uint16)
    n, d = vectors.shape
    nq = queries.shape[0]
    m, ksub, _ = cb.shape
    ids = np.empty((nq, k), np.uint64)
    dist = np.empty((nq, k), np.float32)
    cnt = np.empty(nq, np.uint32)
    st = np.zeros(nq, STATS_DTYPE) if stats else None
    rc = lib().orc_leann_search_adc_rerank(_cfgp(cfg), _p(vectors, f32p), n, d, _p(offsets, u64p), _p(nbrs, u64p),
                                           -1 if entry is None else int(entry), _p(cb, f32p), m, ksub, _p(codes, u16p),
                                           _p(queries, f32p), nq, k, ef, _p(ids, u64p), _p(dist, f32p), _p(cnt, u32p),
                                           st.ctypes.data if stats else None, threads, int(rerank_limit))
    if rc != 0:
        raise RuntimeError(f"orc_leann_search_adc_rerank status {rc}")
    return (ids, dist, cnt, st) if stats else (ids, dist, cnt)


def adc_table_round(values):
    """bfloat16 rounding of the ADC traversal's table entries (oracle.cpp bf16_round)."""
    v = _f32(values).reshape(-1)
    out = np.empty_like(v)
    lib().orc_adc_table_round(_p(v, f32p), v.size, _p(out, f32p))
    return out


def merge_topk(ids, dist, k):
    ids = np.ascontiguousarray(ids, np.uint64)
    dist = _f32(dist)
    parts, nq, _ = ids.shape
    oi = np.empty((nq, k), np.uint64)
    od = np.empty((nq, k), np.float32)
    oc = np.empty(nq, np.uint32)
    lib().orc_merge_topk(_p(ids, u64p), _p(dist, f32p), parts, nq, k, _p(oi, u64p), _p(od, f32p), _p(oc, u32p))
    return oi, od, oc


class Hnsw:
    """orc_hnsw_* (hnsw.rs) — levels are explicit inputs."""

    def __init__(self, cfg, d):
        self._cfg = cfg
        self.d = d
        self._h = lib().orc_hnsw_new(C.cast(C.byref(cfg), C.c_void_p), d)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_hnsw_free(self._h)
            self._h = None

    def insert(self, v, level):
        v = _f32(v)
        out = C.c_uint64()
        rc = lib().orc_hnsw_insert(self._h, _p(v, f32p), level, C.byref(out))
        if rc != 0:
            raise RuntimeError(f"orc_hnsw_insert status {rc}")
        return out.value

    def insert_batch(self, vectors, levels, batch=1, threads=1):
        v = _f32(vectors)
        lv = np.ascontiguousarray(levels, np.uint64)
        rc = lib().orc_hnsw_insert_batch(self._h, _p(v, f32p), v.shape[0], _p(lv, u64p), batch, threads)
        if rc != 0:
            raise RuntimeError(f"orc_hnsw_insert_batch status {rc}")

    def node_level(self, id):
        lv = lib().orc_hnsw_node_level(self._h, id)
        return None if lv < 0 else int(lv)

    def __len__(self):
        return int(lib().orc_hnsw_len(self._h))

    def entry_point(self):
        e = lib().orc_hnsw_entry_point(self._h)
        return None if e < 0 else int(e)

    def max_level(self):
        return int(lib().orc_hnsw_max_level(self._h))

    def neighbors(self, id, layer, cap=4096):
        buf = np.empty(cap, np.uint64)
        c = lib().orc_hnsw_neighbors(self._h, id, layer, _p(buf, u64p), cap)
        return None if c < 0 else buf[:c].copy()

    def search(self, queries, k, ef, threads=1):
        q = _f32(queries)
        q = q.reshape(1, -1) if q.ndim == 1 else q
        nq = q.shape[0]
        ids = np.empty((nq, k), np.uint64)
        dist = np.empty((nq, k), np.float32)
        cnt = np.empty(nq, np.uint32)
        rc = lib().orc_hnsw_search(self._h, _p(q, f32p), nq, k, ef, _p(ids, u64p), _p(dist, f32p), _p(cnt, u32p), threads)
        if rc != 0:
            raise RuntimeError(f"orc_hnsw_search status {rc}")
        return ids, dist, cnt
