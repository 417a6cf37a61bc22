"""CPU oracle of the reference's to_bytes layout — TEST INFRASTRUCTURE ONLY.

`bincode::serialize` (bincode 1.x API; leann.rs:1059-1061, pq.rs:351-353, hnsw.rs:507-509) of the
serde-derived structs, restated independently of the product's C++ writer: little-endian fixed-width
integers, usize as u64, f32 / f64 raw, bool one byte, Vec / HashMap = u64 length + items, Option = one
tag byte + value, unit enum variant = u32 index, struct fields in declaration order.
PARITY UNPINNED: the Rust toolchain is absent, so no bytes produced by the reference exist to pin
this against (SURVEY F11 also notes that Cargo.toml pins bincode 3.0.0 while the code uses the 1.x API).
"""
import struct

import numpy as np


def _u64(v): return struct.pack("<Q", int(v))
def _u32(v): return struct.pack("<I", int(v))
def _f32(v): return struct.pack("<f", float(v))
def _f64(v): return struct.pack("<d", float(v))
def _bool(v): return b"\x01" if v else b"\x00"
def _opt_u64(v): return b"\x00" if v is None else b"\x01" + _u64(v)
def _vec_u64(a): return _u64(len(a)) + np.asarray(a, "<u8").tobytes()
def _vec_f32(a): return _u64(len(a)) + np.asarray(a, "<f4").tobytes()


def leann_config(c):
    """LeannConfig (leann.rs:322-371)."""
    return (_u64(c.m) + _u64(c.m0) + _u64(c.ef_construction) + _f64(c.ml) + _u64(c.max_layers) + _u32(c.metric)
            + _u64(c.ef_search) + _u64(c.beam_width) + _f32(c.prune_ratio) + _u32(c.pruning_strategy)
            + _bool(c.high_degree_pruning) + _f32(c.hub_percentile) + _bool(c.is_compact) + _bool(c.is_recompute))


def leann_index(cfg, node_offsets, neighbors, levels, entry_point, max_level, dimension):
    """LeannIndex { config, graph: CsrGraph, dimension } (leann.rs:493-500, :193-208)."""
    n = len(node_offsets) - 1
    deg = np.diff(np.asarray(node_offsets, np.uint64).astype(np.int64))
    return (leann_config(cfg) + _vec_u64(node_offsets) + _vec_u64(neighbors) + _vec_u64(levels) + _opt_u64(entry_point)
            + _u64(max_level) + _u64(n) + _vec_u64(deg) + _opt_u64(dimension))


def product_quantizer(num_subquantizers, num_centroids, training_iterations, seed, codebooks, dimension, metric, trained):
    """ProductQuantizer (pq.rs:116-129); codebooks [m][ksub][dsub] or None when untrained."""
    out = _u64(num_subquantizers) + _u64(num_centroids) + _u64(training_iterations) + _opt_u64(seed)
    dsub = dimension // num_subquantizers
    if trained:
        out += _u64(len(codebooks))
        for cb in codebooks:
            out += _u64(len(cb))
            for cen in cb:
                out += _vec_f32(cen)
            out += _u64(dsub)
    else:
        out += _u64(0)
    return out + _u64(dimension) + _u64(dsub) + _u32(metric) + _bool(trained)


def hnsw_graph(cfg, vectors, levels, neighbors_of, entry_point, max_level):
    """HnswGraph (hnsw.rs:151-164) with the node map in ascending id order; neighbors_of(id, layer) -> ids."""
    n = len(vectors)
    out = _u64(cfg.m) + _u64(cfg.m0) + _u64(cfg.ef_construction) + _f64(cfg.ml) + _u32(cfg.metric) + _u64(cfg.max_layers)
    out += _u64(n)
    for i in range(n):
        out += _u64(i) + _u64(i) + _vec_f32(vectors[i]) + _u64(int(levels[i]) + 1)
        for layer in range(int(levels[i]) + 1):
            out += _vec_u64(neighbors_of(i, layer))
        out += _u64(levels[i])
    dim = None if n == 0 else len(vectors[0])
    return out + _opt_u64(entry_point) + _u64(max_level) + _opt_u64(dim) + _u64(n)
