"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/islands_b200.h declares, and fails loudly (no CPU fallback) without a GPU."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "islands_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"#ifdef ISL_TEST_HOOKS.*?#endif", "", text, flags=re.S)  # test hooks are not part of the product ABI
    return sorted(set(re.findall(r"\b(isl_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from islands_b200 import _ffi

    lib = _ffi.load()
    declared = _declared_symbols()
    assert len(declared) >= 50
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in islands_b200.h but not exported"
        assert name in _ffi.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_ffi.SIGNATURES) == set(declared)
    assert lib.isl_abi_version() == 2


def test_library_is_built_for_sm_100a_only():
    import subprocess

    from islands_b200 import _ffi

    out = subprocess.run(["cuobjdump", "--list-elf", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_config_defaults_and_validation():
    """leann.rs:1091-1146, hnsw.rs:533-569, pq.rs:479-520 (host logic, no GPU needed)."""
    from islands_b200 import HnswConfig, InvalidConfig, LeannConfig, PQConfig

    c = LeannConfig()
    assert (c.m, c.m0, c.ef_construction, c.ef_search, c.beam_width, c.max_layers) == (30, 60, 128, 64, 1, 16)
    assert c.metric == 0 and c.prune_ratio == 0.0 and c.pruning_strategy == 0
    assert c.high_degree_pruning == 1 and abs(c.hub_percentile - 0.02) < 1e-3
    assert c.is_compact == 1 and c.is_recompute == 1
    assert abs(c.ml - 1.0 / np.log(30.0)) < 1e-15
    c.validate()
    f = LeannConfig.fast()
    assert (f.m, f.m0, f.ef_construction, f.ef_search) == (16, 32, 100, 32) and f.prune_ratio > 0
    a = LeannConfig.accurate()
    assert (a.m, a.m0, a.ef_construction, a.ef_search) == (48, 96, 400, 128) and a.prune_ratio == 0
    for bad in (dict(m=0), dict(m=30, m0=10), dict(ef_construction=5), dict(prune_ratio=1.5),
                dict(prune_ratio=-0.1), dict(beam_width=0), dict(hub_percentile=1.5)):
        with pytest.raises(InvalidConfig):
            LeannConfig(**bad).validate()
    h = HnswConfig()
    assert (h.m, h.m0, h.ef_construction, h.max_layers) == (16, 32, 200, 16)
    h.validate()
    with pytest.raises(InvalidConfig):
        HnswConfig(m=0).validate()
    with pytest.raises(InvalidConfig):
        HnswConfig(m=16, m0=8).validate()
    p = PQConfig()
    assert (p.num_subquantizers, p.num_centroids, p.training_iterations, p.seed) == (8, 256, 25, None)
    p.validate(128)
    with pytest.raises(InvalidConfig):
        p.validate(100)  # not divisible (pq.rs:493-494)
    with pytest.raises(InvalidConfig):
        PQConfig(num_subquantizers=0).validate(128)
    with pytest.raises(InvalidConfig):
        PQConfig(num_centroids=0).validate(128)
    with pytest.raises(InvalidConfig):
        PQConfig(num_centroids=70000).validate(128)
    assert PQConfig(8, 256).bytes_per_vector() == 8      # pq.rs:505-520
    assert PQConfig(8, 512).bytes_per_vector() == 16


def test_host_side_argument_errors_need_no_gpu():
    from islands_b200 import DimensionMismatch, DistanceMetric, to_similarity

    with pytest.raises(DimensionMismatch):  # distance.rs:206-212 — checked before any device work
        DistanceMetric(DistanceMetric.Cosine).calculate([1.0, 2.0], [1.0, 2.0, 3.0])
    with pytest.raises(DimensionMismatch):
        DistanceMetric(DistanceMetric.Euclidean).calculate_squared([1.0, 2.0], [1.0, 2.0, 3.0])
    assert abs(to_similarity(0.0) - 1.0) < 1e-6 and abs(to_similarity(1.0) - 0.5) < 1e-6  # search.rs:311-324


def test_csr_graph_mirror():
    """CsrGraph accessors (leann.rs:1172-1220)."""
    from islands_b200 import CsrGraph

    g = CsrGraph()
    assert g.num_nodes == 0 and g.entry_point is None
    assert g.add_node([], 0) == 0 and g.entry_point == 0
    assert g.add_node([0], 1) == 1 and g.entry_point == 1
    g2 = CsrGraph()
    g2.add_node([], 0)
    g2.add_node([0], 0)
    g2.add_node([0, 1], 0)
    assert list(g2.get_neighbors(0)) == [] and list(g2.get_neighbors(1)) == [0]
    assert list(g2.get_neighbors(2)) == [0, 1] and g2.get_neighbors(999) is None
    assert g2.storage_bytes() == 8 * (4 + 3 + 3 + 3)


def test_no_cpu_fallback_without_gpu():
    from islands_b200 import CudaError, DistanceMetric, _ffi

    if _ffi.load().isl_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(CudaError):
        DistanceMetric(0).calculate([1.0, 0.0], [0.0, 1.0])


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under islands_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "islands_b200")):
        if "lib" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("pyoracle", "libislands_oracle", "oracle.h", "import oracle", "from oracle"):
                    assert needle not in text, (f, needle)
                if f == "api_shard.cu":
                    # the one run-time library load of the product: NCCL (no link-time dependency on it)
                    assert re.findall(r"dlopen\((\w+)", text) == ["nme"]
                    assert re.search(r'names\[\] = \{"libnccl\.so\.2", "libnccl\.so"\}', text)
                else:
                    assert "dlopen(" not in text, (f, "dlopen")


def test_rust_sys_declarations_are_current_and_complete():
    """bindings/rust/islands-b200-sys/src/lib.rs (the `extern "C"` module of INTEGRATION.md §1) is generated from the
    header: the committed file must be what scripts/gen_rust_sys.py produces now, and declare every symbol of
    the ctypes table (which test_header_library_and_table_agree ties to the header and the library)."""
    import importlib.util
    import re

    from islands_b200 import _ffi

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_rust_sys", os.path.join(root, "scripts", "gen_rust_sys.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text, names = gen.generate()
    with open(gen.OUT) as f:
        assert f.read() == text, "run python scripts/gen_rust_sys.py"
    declared = set(re.findall(r"pub fn (isl_\w+)\(", text))
    assert declared == set(_ffi.SIGNATURES)
    assert "pub struct IslLeannConfig" in text and "pub hub_percentile: f32" in text


def test_rust_wrapper_uses_only_declared_symbols_with_the_right_arity():
    """bindings/rust/islands-b200/src/lib.rs (the safe wrapper with the reference's signatures) cannot be compiled
    here (no rustc): at least every `sys::isl_*` call it makes must name a function of the generated -sys crate
    and pass as many arguments as that declaration takes, and every `sys::ISL_*` constant must exist."""
    sys_text = open(os.path.join(ROOT, "bindings", "rust", "islands-b200-sys", "src", "lib.rs")).read()
    decl = {m.group(1): (0 if not m.group(2).strip() else m.group(2).count(":"))
            for m in re.finditer(r"pub fn (isl_\w+)\(([^)]*)\)", sys_text)}
    consts = set(re.findall(r"pub const (ISL_\w+):", sys_text))
    structs = set(re.findall(r"pub struct (Isl\w+)", sys_text))
    text = open(os.path.join(ROOT, "bindings", "rust", "islands-b200", "src", "lib.rs")).read()
    text = re.sub(r"//[^\n]*", "", text)
    calls = 0
    for m in re.finditer(r"sys::(isl_\w+)\s*\(", text):
        name, i, depth, args, cur = m.group(1), m.end(), 1, 0, ""
        assert name in decl, name
        while depth:
            ch = text[i]
            depth += ch in "([{"
            depth -= ch in ")]}"
            if ch == "," and depth == 1:
                args += bool(cur.strip())
                cur = ""
            elif depth:
                cur += ch
            i += 1
        args += bool(cur.strip())
        assert args == decl[name], (name, args, decl[name])
        calls += 1
    assert calls >= 55
    for c in re.findall(r"sys::(ISL_\w+)", text):
        assert c in consts, c
    for s_ in re.findall(r"sys::(Isl\w+)", text):
        assert s_ in structs, s_
    # the reference surface the wrapper promises (src/core/mod.rs:60-99)
    for needle in ("pub fn build<P: EmbeddingProvider>(&mut self, provider: &P, num_vectors: usize) -> CoreResult<()>",
                   "pub fn search<P: EmbeddingProvider>(&self, query: &[f32], k: usize, provider: &P) -> CoreResult<Vec<(u64, f32)>>",
                   "pub fn search_with_params<P: EmbeddingProvider>(&self, query: &[f32], k: usize, ef: usize, _provider: &P) -> CoreResult<Vec<(u64, f32)>>",
                   "pub fn encode(&self, vector: &[f32]) -> CoreResult<Vec<u16>>", "pub fn decode(&self, codes: &[u16]) -> CoreResult<Vec<f32>>",
                   "pub fn set_neighbors(&mut self, node_id: u64, new_neighbors: Vec<u64>)", "DimensionMismatch { expected: usize, actual: usize }",
                   # hnsw.rs:166-514 and search.rs:106-248 — what IndexerService calls (service.rs:622, 655-657, 781-785)
                   "pub fn insert(&mut self, vector: Vec<f32>) -> CoreResult<u64>",
                   "pub fn search(&self, query: &[f32], k: usize, ef: usize) -> CoreResult<Vec<(u64, f32)>>",
                   "pub fn get_node(&self, id: u64) -> Option<HnswNode>", "pub fn from_bytes(bytes: &[u8]) -> CoreResult<Self>",
                   "pub fn search(&self, query: &[f32]) -> CoreResult<Vec<SearchResult>>",
                   "pub fn search_batch(&self, queries: &[Vec<f32>]) -> CoreResult<Vec<Vec<SearchResult>>>",
                   "pub fn search(&self, query: &[f32]) -> CoreResult<Vec<(String, SearchResult)>>",
                   "pub fn add_index(&mut self, name: impl Into<String>, graph: HnswGraph)", "pub mod prelude"):
        assert needle in text, needle


def test_error_payloads_travel_beside_the_message():
    """CoreError::DimensionMismatch{expected, actual} (error.rs:12-18): the numbers come back through
    isl_last_error_detail, so a binding can rebuild the full variant (no device needed for this check)."""
    import ctypes as C

    from islands_b200 import DimensionMismatch, DistanceMetric, _ffi

    with pytest.raises(DimensionMismatch):
        DistanceMetric(1).calculate([1.0, 2.0, 3.0], [1.0, 2.0])  # distance.rs:39-44, checked before any device work
    a, b = C.c_uint64(), C.c_uint64()
    _ffi.load().isl_last_error_detail(C.byref(a), C.byref(b))
    assert (a.value, b.value) == (3, 2)


def test_csr_graph_set_neighbors_host_mirror():
    """CsrGraph::set_neighbors (leann.rs:256-293): in place for equal length, rebuild otherwise, unknown node ignored."""
    from islands_b200 import CsrGraph

    g = CsrGraph()
    g.add_node([], 0)
    g.add_node([0], 0)
    g.add_node([0, 1], 1)
    g.set_neighbors(2, [1, 0])
    assert list(g.get_neighbors(2)) == [1, 0] and list(g.node_offsets) == [0, 0, 1, 3]
    g.set_neighbors(0, [1, 2])
    assert list(g.get_neighbors(0)) == [1, 2] and list(g.get_neighbors(1)) == [0] and list(g.get_neighbors(2)) == [1, 0]
    assert list(g.node_offsets) == [0, 2, 3, 5] and list(g.degree_counts) == [2, 1, 2]
    g.set_neighbors(1, [])
    assert list(g.node_offsets) == [0, 2, 2, 4] and list(g.get_neighbors(2)) == [1, 0]
    g.set_neighbors(99, [0])
    assert g.num_nodes == 3 and g.neighbors.size == 4


def test_nothing_unwinds_through_the_abi():
    """Errors never abort: every `isl_status` entry point is a function-try-block closed by ISL_ABI_GUARD,
    so a C++ exception raised below it (std::bad_alloc / std::length_error from a host container, anything else) comes
    back as a status with a message.  Shown with the test hook that throws inside an entry point, and kept true for
    every entry point by a scan of the sources."""
    from islands_b200 import _ffi

    lib = _ffi.load()
    assert lib.isl_test_raise(3) == 0
    for kind, text in ((0, b"std::bad_alloc"), (1, b"host exception: "), (2, b"unknown type")):
        assert lib.isl_test_raise(kind) == 9  # ISL_INVALID_ARGUMENT
        assert text in lib.isl_last_error()
    guarded = 0
    src = os.path.join(ROOT, "islands_b200", "csrc")
    for name in sorted(os.listdir(src)):
        if not name.endswith(".cu"):
            continue
        lines = open(os.path.join(src, name)).read().split("\n")
        for i, line in enumerate(lines):
            if re.match(r"^isl_status isl_\w+\(", line):
                j = i
                while not lines[j].rstrip().endswith("{"):
                    j += 1
                assert lines[j].rstrip().endswith(") try {"), (name, i + 1)
                k = j + 1
                while not lines[k].startswith("}"):
                    k += 1
                assert lines[k].startswith("} ISL_ABI_GUARD"), (name, k + 1)
                guarded += 1
    assert guarded >= 87


def test_last_error_is_per_thread():
    """`isl_last_error` / `isl_last_error_detail` are thread-local (header): concurrent failing calls on different
    threads each read back their own message and payload (no device needed: the length check comes first)."""
    import ctypes as C
    import threading

    from islands_b200 import _ffi

    lib = _ffi.load()
    errors = []

    def worker(t):
        a = (C.c_float * (t + 2))()
        b = (C.c_float * 1)()
        out = C.c_float()
        for _ in range(300):
            st = lib.isl_distance_calculate(1, a, t + 2, b, 1, C.byref(out))
            x, y = C.c_uint64(), C.c_uint64()
            lib.isl_last_error_detail(C.byref(x), C.byref(y))
            msg = lib.isl_last_error().decode()
            if st != 1 or (x.value, y.value) != (t + 2, 1) or f"expected {t + 2}, got 1" not in msg:
                errors.append((t, st, x.value, y.value, msg))
                return

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors[:3]
