"""The flat index files of docs/leann-specification.md:967-1027 — graph (`.hnsw`), PQ codebook (`.pq`) and PQ codes
(`.codes`) — the mmap-friendly companions of the bincode images (`to_bytes`, islands_b200.core) and of the `META`
container (islands_b200.storage).  Host-side format code only: every array section is a little-endian u32 / u8 / f32
run at a 4-byte-aligned file offset, so a reader maps the file (`numpy.memmap`) and hands the views to
`LeannIndex.from_csr` / `ProductQuantizer.set_codebooks` / `LeannIndex.attach_pq` without parsing.

The reference has no code for these files (the spec is a design document), so this module IS the definition where the
spec leaves a choice; the choices, all stated once here:

* integers little-endian, structures packed in the order the spec lists the fields;
* the `.hnsw` header is 64 bytes as the spec says — its field list adds up to 26 bytes + `_reserved`, so `_reserved`
  is the 38 bytes that make 64 (the spec's "40" does not add up); `entry_point` = 0xFFFFFFFF for an empty graph;
  `metric` 0 = L2, 1 = Cosine, 2 = InnerProduct as in the spec, 3 = Manhattan (the fourth `DistanceMetric` of
  distance.rs:9-19, which the spec does not number);
* a layer header is 16 bytes: `layer_id u8, 3 zero bytes, num_nodes u32, edges_start u64` (the natural alignment of the
  three fields, which keeps `row_ptr` 4-byte aligned); `edges_start` is the absolute file offset of the layer's
  `edges`; every layer's `row_ptr` runs over ALL node ids (`num_nodes + 1` entries; a node that does not reach the
  layer has an empty row), so a row is addressed by the node id on every layer;
* `SourceRef` (the spec never spells it out) is the 12-byte record `{source u32, chunk_start u32, chunk_end u32}` of
  `NodeMetadata` (spec :43-54): `source` is the row of the node in the token table the recompute search embeds
  (`isl_index_set_recompute`), the chunk bounds are the caller's;
* `.pq`: the optional `centroid_norms` section (Σc² per centroid, f32, the left-to-right fold of distance.rs:79) is
  present exactly when the file is long enough to hold it;
* `.codes` holds one byte per code (`num_centroids ≤ 256`, pq.rs:58-64); wider codes have no place in the spec's
  format and are refused.
"""
import os
import struct
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from .core import SerializationError
from .storage import DeserializationError

GRAPH_MAGIC = b"HNSW"
CODEBOOK_MAGIC = b"PQCB"
CODES_MAGIC = b"PQCD"
FORMAT_VERSION = 1
NO_ENTRY_POINT = 0xFFFFFFFF

_GRAPH_HEADER = struct.Struct("<4sIIBIBIHH38x")   # 64 bytes
_LAYER_HEADER = struct.Struct("<B3xIQ")           # 16 bytes
_CODEBOOK_HEADER = struct.Struct("<4sIHHH18x")    # 32 bytes
_CODES_HEADER = struct.Struct("<4sIIB3x")         # 16 bytes
assert _GRAPH_HEADER.size == 64 and _LAYER_HEADER.size == 16
assert _CODEBOOK_HEADER.size == 32 and _CODES_HEADER.size == 16

SOURCE_REF_DTYPE = np.dtype([("source", "<u4"), ("chunk_start", "<u4"), ("chunk_end", "<u4")])

# spec numbering (0 = L2, 1 = Cosine, 2 = InnerProduct) <-> DistanceMetric values of islands_b200.core
_METRIC_TO_FILE = {0: 1, 1: 0, 2: 2, 3: 3}
_METRIC_FROM_FILE = {v: k for k, v in _METRIC_TO_FILE.items()}


def _u32_array(a, what):
    a = np.asarray(a)
    if a.size and (a.min() < 0 or a.max() > 0xFFFFFFFF):
        raise SerializationError(f"{what} does not fit 32 bits (the file format stores u32)")
    return np.ascontiguousarray(a, dtype="<u4")


def _open_for_write(path):
    parent = os.path.dirname(os.fspath(path))
    if parent:
        os.makedirs(parent, exist_ok=True)
    return open(path, "wb")


def _view(path, size, dtype, offset, count, mmap):
    """`count` items of `dtype` at `offset`; a read-only memory map, or a copy when `mmap` is false."""
    dtype = np.dtype(dtype)
    if offset + count * dtype.itemsize > size:
        raise DeserializationError("failed to fill whole buffer")  # truncated file
    if count == 0:
        return np.zeros(0, dtype)
    if mmap:
        return np.memmap(path, dtype=dtype, mode="r", offset=offset, shape=(count,))
    with open(path, "rb") as f:
        f.seek(offset)
        return np.fromfile(f, dtype=dtype, count=count)


# ---- .hnsw ------------------------------------------------------------------------------------------------------------

@dataclass
class GraphFile:
    """One `.hnsw` file: header fields, per-layer CSR (`row_ptr` over all node ids, `edges`), hub ids, source refs."""
    num_nodes: int
    entry_point: Optional[int]
    metric: int                      # DistanceMetric value of islands_b200.core
    dimension: int
    m: int
    ef_construction: int
    layers: List[Tuple[np.ndarray, np.ndarray]] = field(default_factory=list)
    hub_ids: np.ndarray = field(default_factory=lambda: np.zeros(0, "<u4"))
    source_refs: Optional[np.ndarray] = None   # SOURCE_REF_DTYPE[num_nodes]

    @property
    def num_layers(self) -> int:
        return len(self.layers)

    def to_csr(self, layer: int = 0):
        """(`node_offsets`, `neighbors`) of a layer as the u64 arrays of `CsrGraph` (leann.rs:193-208) — what
        `LeannIndex.from_csr` takes (ids are u64 at the ABI, u32 in the file and on the device)."""
        row_ptr, edges = self.layers[layer]
        return np.asarray(row_ptr, np.uint64), np.asarray(edges, np.uint64)

    def levels(self) -> np.ndarray:
        """Top layer of every node (`CsrGraph::levels`): the highest layer on which its row is not empty; a node with
        no edges anywhere is on layer 0."""
        lv = np.zeros(self.num_nodes, np.uint64)
        for li, (row_ptr, _) in enumerate(self.layers):
            if li:
                lv[np.diff(np.asarray(row_ptr, np.int64)) > 0] = li
        return lv


def write_graph_file(path, graph: GraphFile) -> int:
    """Write `graph` as a `.hnsw` file; returns the number of bytes written."""
    n = int(graph.num_nodes)
    if not 0 <= n < NO_ENTRY_POINT:
        raise SerializationError("num_nodes does not fit the u32 header field")
    if not 0 < len(graph.layers) <= 255:
        raise SerializationError("a graph file holds 1..255 layers")
    if not (0 <= graph.m <= 0xFFFF and 0 <= graph.ef_construction <= 0xFFFF):
        raise SerializationError("m / ef_construction do not fit the u16 header fields")
    if graph.metric not in _METRIC_TO_FILE:
        raise SerializationError(f"unknown metric {graph.metric}")
    ep = NO_ENTRY_POINT if graph.entry_point is None else int(graph.entry_point)
    if graph.entry_point is not None and not 0 <= ep < n:
        raise SerializationError("entry_point is not a node of the graph")
    layers = []
    for li, (row_ptr, edges) in enumerate(graph.layers):
        rp = _u32_array(row_ptr, f"layer {li}: row_ptr (more than 2^32 - 1 edges in one layer)")
        ed = _u32_array(edges, f"layer {li}: edge ids")
        if rp.size != n + 1 or (n + 1 and rp[0] != 0) or int(rp[-1]) != ed.size or np.any(np.diff(rp.astype(np.int64)) < 0):
            raise SerializationError(f"layer {li}: row_ptr must hold num_nodes + 1 ascending offsets ending at len(edges)")
        if ed.size and int(ed.max()) >= n:
            raise SerializationError(f"layer {li}: edge id out of range")
        layers.append((rp, ed))
    hubs = _u32_array(graph.hub_ids, "hub ids")
    if hubs.size and int(hubs.max()) >= n:
        raise SerializationError("hub id out of range")
    refs = graph.source_refs
    if refs is None:
        refs = np.zeros(n, SOURCE_REF_DTYPE)
        refs["source"] = np.arange(n, dtype=np.uint32)  # node i recomputes from token row i
    refs = np.ascontiguousarray(refs, SOURCE_REF_DTYPE)
    if refs.shape != (n,):
        raise SerializationError("source_refs must hold one record per node")

    with _open_for_write(path) as f:
        f.write(_GRAPH_HEADER.pack(GRAPH_MAGIC, FORMAT_VERSION, n, len(layers), ep, _METRIC_TO_FILE[graph.metric],
                                   int(graph.dimension), int(graph.m), int(graph.ef_construction)))
        pos = _GRAPH_HEADER.size
        for li, (rp, ed) in enumerate(layers):
            edges_start = pos + _LAYER_HEADER.size + rp.nbytes
            f.write(_LAYER_HEADER.pack(li, n, edges_start))
            rp.tofile(f)
            ed.tofile(f)
            pos = edges_start + ed.nbytes
        f.write(struct.pack("<I", hubs.size))
        hubs.tofile(f)
        refs.tofile(f)
        return pos + 4 + hubs.nbytes + refs.nbytes


def read_graph_file(path, mmap: bool = True) -> GraphFile:
    """Open a `.hnsw` file.  With `mmap` (default) every array of the result is a read-only view of the mapped file."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        head = f.read(_GRAPH_HEADER.size)
    if len(head) != _GRAPH_HEADER.size:
        raise DeserializationError("failed to fill whole buffer")
    magic, version, n, num_layers, ep, metric, dim, m, efc = _GRAPH_HEADER.unpack(head)
    if magic != GRAPH_MAGIC:
        raise DeserializationError("not a graph file (magic is not HNSW)")
    if version != FORMAT_VERSION:
        raise DeserializationError(f"unsupported graph file version {version}")
    if metric not in _METRIC_FROM_FILE:
        raise DeserializationError(f"unknown metric code {metric}")
    if num_layers == 0:
        raise DeserializationError("a graph file holds at least one layer")
    if ep != NO_ENTRY_POINT and ep >= n:
        raise DeserializationError("entry_point is not a node of the graph")
    pos = _GRAPH_HEADER.size
    layers = []
    for li in range(num_layers):
        lh = _view(path, size, np.uint8, pos, _LAYER_HEADER.size, False).tobytes()
        layer_id, ln, edges_start = _LAYER_HEADER.unpack(lh)
        if layer_id != li or ln != n or edges_start != pos + _LAYER_HEADER.size + 4 * (n + 1):
            raise DeserializationError(f"layer {li}: inconsistent layer header")
        row_ptr = _view(path, size, "<u4", pos + _LAYER_HEADER.size, n + 1, mmap)
        if int(row_ptr[0]) != 0:
            raise DeserializationError(f"layer {li}: row_ptr does not start at 0")
        edges = _view(path, size, "<u4", edges_start, int(row_ptr[-1]), mmap)
        layers.append((row_ptr, edges))
        pos = edges_start + 4 * int(row_ptr[-1])
    (num_hubs,) = struct.unpack("<I", _view(path, size, np.uint8, pos, 4, False).tobytes())
    hubs = _view(path, size, "<u4", pos + 4, num_hubs, mmap)
    refs = _view(path, size, SOURCE_REF_DTYPE, pos + 4 + 4 * num_hubs, n, mmap)
    return GraphFile(n, None if ep == NO_ENTRY_POINT else ep, _METRIC_FROM_FILE[metric], dim, m, efc, layers, hubs, refs)


def graph_file_from_csr(csr, config, dimension, hub_ids=(), source_refs=None) -> GraphFile:
    """The single-layer LEANN graph (`CsrGraph`, leann.rs:193-208) as a `GraphFile`: layer 0 carries the CSR rows;
    `levels` are not edges and are not stored (the LEANN search never reads them, leann.rs:868-988)."""
    metric = getattr(config.metric, "value", config.metric)
    return GraphFile(int(csr.num_nodes), None if csr.entry_point is None else int(csr.entry_point), int(metric),
                     int(dimension), int(config.m), int(config.ef_construction),
                     [(np.asarray(csr.node_offsets), np.asarray(csr.neighbors))], np.asarray(hub_ids), source_refs)


def layer_from_export(degrees, neighbors) -> Tuple[np.ndarray, np.ndarray]:
    """One layer as `HnswGraph.export_layer` hands it out — `degrees [n]` (-1 where the node does not reach the layer)
    and `neighbors [n][M]` padded rows — as the file's (`row_ptr`, `edges`): an absent node gets an empty row."""
    deg = np.maximum(np.asarray(degrees, np.int64), 0)
    nb = np.asarray(neighbors)
    row_ptr = np.zeros(deg.size + 1, np.int64)
    np.cumsum(deg, out=row_ptr[1:])
    keep = np.arange(nb.shape[1] if nb.ndim == 2 else 0)[None, :] < deg[:, None]
    return row_ptr, (nb[keep] if nb.size else np.zeros(0, np.uint64))


def graph_file_from_hnsw(graph, hub_ids=(), source_refs=None) -> GraphFile:
    """Every layer of an `HnswGraph` handle (hnsw.rs:151-164: per-node connection lists per layer) as a `GraphFile`."""
    cfg = graph.config
    layers = [layer_from_export(*graph.export_layer(layer)) for layer in range(int(graph.max_level) + 1)]
    return GraphFile(len(graph), graph.entry_point, int(getattr(cfg.metric, "value", cfg.metric)), int(graph.dimension() or 0),
                     int(cfg.m), int(cfg.ef_construction), layers, np.asarray(hub_ids), source_refs)


def hubs_by_in_degree(row_ptr, edges, fraction: float) -> np.ndarray:
    """The `ceil(fraction · n)` nodes of highest in-degree, ties to the lower id, ascending by id — the set the
    hub-embedding cache keeps resident (docs/leann-specification.md:661-690; `isl_index_set_hub_cache`)."""
    n = len(row_ptr) - 1
    count = min(n, int(np.ceil(fraction * n))) if n and fraction > 0 else 0
    if count == 0:
        return np.zeros(0, "<u4")
    indeg = np.bincount(np.asarray(edges, np.int64), minlength=n)
    order = np.lexsort((np.arange(n), -indeg))  # in-degree descending, id ascending
    return np.sort(order[:count]).astype("<u4")


# ---- .pq --------------------------------------------------------------------------------------------------------------

def write_codebook_file(path, codebooks, with_norms: bool = True) -> int:
    """`codebooks`: f32 `[num_subspaces][num_centroids][subspace_dim]` (`ProductQuantizer.codebooks()`, pq.rs:116-126)."""
    cb = np.ascontiguousarray(codebooks, "<f4")
    if cb.ndim != 3 or min(cb.shape) < 1 or max(cb.shape) > 0xFFFF:
        raise SerializationError("codebooks must be [num_subspaces][num_centroids][subspace_dim], each 1..65535")
    with _open_for_write(path) as f:
        f.write(_CODEBOOK_HEADER.pack(CODEBOOK_MAGIC, FORMAT_VERSION, cb.shape[0], cb.shape[1], cb.shape[2]))
        cb.tofile(f)
        written = _CODEBOOK_HEADER.size + cb.nbytes
        if with_norms:
            norms = np.zeros(cb.shape[:2], np.float32)
            for j in range(cb.shape[2]):  # left-to-right f32 fold, one rounding per multiply and per add
                norms = (norms + cb[:, :, j] * cb[:, :, j]).astype(np.float32)
            norms.astype("<f4").tofile(f)
            written += norms.nbytes
        return written


def read_codebook_file(path, mmap: bool = True):
    """→ (`codebooks [m][ksub][dsub]`, `centroid_norms [m][ksub]` or None)."""
    size = os.path.getsize(path)
    head = _view(path, size, np.uint8, 0, _CODEBOOK_HEADER.size, False).tobytes()
    magic, version, m, ksub, dsub = _CODEBOOK_HEADER.unpack(head)
    if magic != CODEBOOK_MAGIC:
        raise DeserializationError("not a codebook file (magic is not PQCB)")
    if version != FORMAT_VERSION:
        raise DeserializationError(f"unsupported codebook file version {version}")
    if min(m, ksub, dsub) < 1:
        raise DeserializationError("empty codebook")
    cb = _view(path, size, "<f4", _CODEBOOK_HEADER.size, m * ksub * dsub, mmap).reshape(m, ksub, dsub)
    norms_at = _CODEBOOK_HEADER.size + 4 * m * ksub * dsub
    rest = size - norms_at
    if rest == 0:
        return cb, None
    if rest != 4 * m * ksub:
        raise DeserializationError("trailing bytes are not a centroid_norms section")
    return cb, _view(path, size, "<f4", norms_at, m * ksub, mmap).reshape(m, ksub)


# ---- .codes -----------------------------------------------------------------------------------------------------------

def write_codes_file(path, codes) -> int:
    """`codes`: `[num_vectors][num_subspaces]`, every value < 256 (`ProductQuantizer.encode`, pq.rs:221-244)."""
    c = np.asarray(codes)
    if c.ndim != 2 or not 1 <= c.shape[1] <= 255 or c.shape[0] > 0xFFFFFFFF:
        raise SerializationError("codes must be [num_vectors][num_subspaces] with 1..255 subspaces")
    if c.size and (c.min() < 0 or c.max() > 255):
        raise SerializationError("the codes file stores one byte per code (num_centroids <= 256)")
    c8 = np.ascontiguousarray(c, np.uint8)
    with _open_for_write(path) as f:
        f.write(_CODES_HEADER.pack(CODES_MAGIC, FORMAT_VERSION, c8.shape[0], c8.shape[1]))
        c8.tofile(f)
    return _CODES_HEADER.size + c8.nbytes


def read_codes_file(path, mmap: bool = True) -> np.ndarray:
    """→ u8 `[num_vectors][num_subspaces]` (widen with `.astype(np.uint16)` for `LeannIndex.attach_pq`)."""
    size = os.path.getsize(path)
    head = _view(path, size, np.uint8, 0, _CODES_HEADER.size, False).tobytes()
    magic, version, n, m = _CODES_HEADER.unpack(head)
    if magic != CODES_MAGIC:
        raise DeserializationError("not a codes file (magic is not PQCD)")
    if version != FORMAT_VERSION:
        raise DeserializationError(f"unsupported codes file version {version}")
    if m == 0:
        raise DeserializationError("num_subspaces is zero")
    if size != _CODES_HEADER.size + n * m:
        raise DeserializationError("file length does not match num_vectors * num_subspaces")
    return _view(path, size, np.uint8, _CODES_HEADER.size, n * m, mmap).reshape(n, m)
