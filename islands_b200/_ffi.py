"""ctypes binding of the C ABI in include/islands_b200.h (libislands_b200.so, built in-tree).

There is no fallback: if the shared library is missing this module raises ImportError with the
build command, and if no CUDA device is usable every compute entry point returns
ISL_CUDA_ERROR, which is raised as CudaError.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ISL_DEV_LIB_PATH") or os.path.join(_HERE, "lib", "libislands_b200.so")

ISL_OK = 0
ISL_NO_ENTRY = -1
ISL_INVALID_ID = 0xFFFFFFFFFFFFFFFF

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u16p = C.POINTER(C.c_uint16)
f32p = C.POINTER(C.c_float)


class LeannConfigStruct(C.Structure):
    """isl_leann_config == LeannConfig (src/core/leann.rs:322-371), same field order."""

    _fields_ = [
        ("m", C.c_uint64),
        ("m0", C.c_uint64),
        ("ef_construction", C.c_uint64),
        ("ml", C.c_double),
        ("max_layers", C.c_uint64),
        ("metric", C.c_int32),
        ("ef_search", C.c_uint64),
        ("beam_width", C.c_uint64),
        ("prune_ratio", C.c_float),
        ("pruning_strategy", C.c_int32),
        ("high_degree_pruning", C.c_int32),
        ("hub_percentile", C.c_float),
        ("is_compact", C.c_int32),
        ("is_recompute", C.c_int32),
        ("prune_seed", C.c_uint64),
    ]


class HnswConfigStruct(C.Structure):
    """isl_hnsw_config == HnswConfig (src/core/hnsw.rs:15-28)."""

    _fields_ = [
        ("m", C.c_uint64),
        ("m0", C.c_uint64),
        ("ef_construction", C.c_uint64),
        ("ml", C.c_double),
        ("metric", C.c_int32),
        ("max_layers", C.c_uint64),
    ]


class PQConfigStruct(C.Structure):
    """isl_pq_config == PQConfig (src/core/pq.rs:13-22); seed: Option<u64> as (seed, has_seed)."""

    _fields_ = [
        ("num_subquantizers", C.c_uint64),
        ("num_centroids", C.c_uint64),
        ("training_iterations", C.c_uint64),
        ("seed", C.c_uint64),
        ("has_seed", C.c_int32),
    ]


class EncoderConfigStruct(C.Structure):
    """isl_encoder_config: BERT shape of the recompute encoder."""

    _fields_ = [
        ("vocab_size", C.c_uint32),
        ("hidden_size", C.c_uint32),
        ("num_layers", C.c_uint32),
        ("num_heads", C.c_uint32),
        ("intermediate_size", C.c_uint32),
        ("max_position", C.c_uint32),
        ("type_vocab_size", C.c_uint32),
        ("layer_norm_eps", C.c_float),
        ("normalize", C.c_int32),
        ("precision", C.c_int32),
    ]


class BuildStatsStruct(C.Structure):
    """isl_build_stats."""

    _fields_ = [
        ("n_hop", C.c_uint64),
        ("n_edge", C.c_uint64),
        ("n_dist", C.c_uint64),
        ("edges", C.c_uint64),
        ("rounds", C.c_uint64),
        ("search_ms", C.c_float),
        ("rounds_ms", C.c_float),
    ]


class SearchStatsStruct(C.Structure):
    _fields_ = [
        ("n_hop", C.c_uint64),
        ("n_edge", C.c_uint64),
        ("n_dist", C.c_uint64),
        ("n_adc", C.c_uint64),
        ("n_rerank", C.c_uint64),
    ]


_LCP = C.POINTER(LeannConfigStruct)
_HCP = C.POINTER(HnswConfigStruct)
_PCP = C.POINTER(PQConfigStruct)
_SSP = C.POINTER(SearchStatsStruct)
_ECP = C.POINTER(EncoderConfigStruct)
i32p = C.POINTER(C.c_int32)
_VP = C.c_void_p
_VPP = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol declared in include/islands_b200.h is listed here and
# tests/test_abi.py checks the two against each other.
SIGNATURES = {
    "isl_abi_version": (C.c_int, []),
    "isl_last_error": (C.c_char_p, []),
    "isl_last_error_detail": (None, [u64p, u64p]),
    "isl_device_count": (C.c_int, []),
    "isl_kernel_launch_count": (C.c_uint64, []),
    "isl_kernel_launch_count_reset": (None, []),
    "isl_leann_config_default": (C.c_int, [_LCP]),
    "isl_leann_config_fast": (C.c_int, [_LCP]),
    "isl_leann_config_accurate": (C.c_int, [_LCP]),
    "isl_leann_config_validate": (C.c_int, [_LCP]),
    "isl_hnsw_config_default": (C.c_int, [_HCP]),
    "isl_hnsw_config_validate": (C.c_int, [_HCP]),
    "isl_pq_config_default": (C.c_int, [_PCP]),
    "isl_pq_config_validate": (C.c_int, [_PCP, C.c_uint64]),
    "isl_pq_config_bytes_per_vector": (C.c_uint64, [_PCP]),
    "isl_distance_calculate": (C.c_int, [C.c_int32, f32p, C.c_uint64, f32p, C.c_uint64, f32p]),
    "isl_distance_calculate_squared": (C.c_int, [C.c_int32, f32p, C.c_uint64, f32p, C.c_uint64, f32p]),
    "isl_distance_batch": (C.c_int, [C.c_int32, f32p, f32p, C.c_uint64, C.c_uint32, f32p]),
    "isl_distance_batch_dev": (C.c_int, [C.c_int32, _VP, _VP, C.c_uint64, C.c_uint32, _VP]),
    "isl_normalize_rows": (C.c_int, [f32p, C.c_uint64, C.c_uint32]),
    "isl_index_from_csr": (C.c_int, [_LCP, C.c_uint32, C.c_uint64, u64p, u64p, u64p, C.c_int64, f32p, _VPP]),
    "isl_index_build": (C.c_int, [_LCP, C.c_uint32, C.c_uint64, f32p, u64p, C.c_uint64, C.c_uint32, _VPP]),
    "isl_index_build_dev": (C.c_int, [_LCP, C.c_uint32, C.c_uint64, _VP, u64p, C.c_uint64, C.c_uint32, _VPP]),
    "isl_index_free": (None, [_VP]),
    "isl_index_len": (C.c_uint64, [_VP]),
    "isl_index_dimension": (C.c_uint32, [_VP]),
    "isl_index_num_edges": (C.c_uint64, [_VP]),
    "isl_index_entry_point": (C.c_int64, [_VP]),
    "isl_index_max_level": (C.c_uint64, [_VP]),
    "isl_index_storage_bytes": (C.c_uint64, [_VP]),
    "isl_index_export_csr": (C.c_int, [_VP, u64p, u64p, u64p, u64p]),
    "isl_index_get_neighbors": (C.c_int, [_VP, C.c_uint64, u64p, C.c_uint64, u64p]),
    "isl_index_search": (C.c_int, [_VP, f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u64p, f32p, u32p, _SSP]),
    "isl_index_search_dev": (C.c_int, [_VP, _VP, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _VP, _VP, _VP, _VP]),
    "isl_index_search_default": (C.c_int, [_VP, f32p, C.c_uint64, C.c_uint32, C.c_uint32, u64p, f32p, u32p]),
    "isl_index_last_search_timing": (C.c_int, [_VP, f32p, u64p]),
    "isl_pq_new": (C.c_int, [C.c_uint32, _PCP, _VPP]),
    "isl_pq_free": (None, [_VP]),
    "isl_pq_set_metric": (C.c_int, [_VP, C.c_int32]),
    "isl_pq_is_trained": (C.c_int32, [_VP]),
    "isl_pq_num_subquantizers": (C.c_uint64, [_VP]),
    "isl_pq_compression_ratio": (C.c_float, [_VP]),
    "isl_pq_train": (C.c_int, [_VP, f32p, C.c_uint64, C.c_uint32]),
    "isl_pq_set_codebooks": (C.c_int, [_VP, f32p, C.c_uint64]),
    "isl_pq_get_codebooks": (C.c_int, [_VP, f32p, u64p]),
    "isl_pq_encode": (C.c_int, [_VP, f32p, C.c_uint64, C.c_uint32, u16p]),
    "isl_pq_decode": (C.c_int, [_VP, u16p, C.c_uint64, C.c_uint64, f32p]),
    "isl_pq_build_tables": (C.c_int, [_VP, f32p, C.c_uint32, f32p]),
    "isl_pq_table_distance": (C.c_int, [_VP, f32p, u16p, C.c_uint64, f32p]),
    "isl_pq_asymmetric_distance": (C.c_int, [_VP, f32p, C.c_uint32, u16p, C.c_uint64, f32p]),
    "isl_index_attach_pq": (C.c_int, [_VP, _VP, u16p]),
    "isl_index_search_two_level": (C.c_int, [_VP, f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, u64p, f32p, u32p, _SSP]),
    "isl_index_search_adc_rerank": (C.c_int, [_VP, f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u64p, f32p, u32p, _SSP]),
    "isl_hnsw_new": (C.c_int, [_HCP, _VPP]),
    "isl_hnsw_free": (None, [_VP]),
    "isl_hnsw_len": (C.c_uint64, [_VP]),
    "isl_hnsw_dimension": (C.c_uint32, [_VP]),
    "isl_hnsw_entry_point": (C.c_int64, [_VP]),
    "isl_hnsw_max_level": (C.c_uint64, [_VP]),
    "isl_hnsw_insert_batch": (C.c_int, [_VP, f32p, C.c_uint64, C.c_uint32, u64p, C.c_uint64, C.c_uint32, u64p]),
    "isl_hnsw_insert_batch_dev": (C.c_int, [_VP, _VP, C.c_uint64, C.c_uint32, u64p, C.c_uint64, C.c_uint32, u64p]),
    "isl_hnsw_node_level": (C.c_int, [_VP, C.c_uint64, u64p]),
    "isl_hnsw_get_neighbors": (C.c_int, [_VP, C.c_uint64, C.c_uint64, u64p, C.c_uint64, u64p]),
    "isl_hnsw_export_layer": (C.c_int, [_VP, C.c_uint64, C.POINTER(C.c_int64), u64p]),
    "isl_hnsw_search": (C.c_int, [_VP, f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u64p, f32p, u32p]),
    "isl_hnsw_search_dev": (C.c_int, [_VP, _VP, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _VP, _VP, _VP]),
    "isl_hnsw_last_search_timing": (C.c_int, [_VP, f32p]),
    "isl_encoder_config_default": (C.c_int, [_ECP]),
    "isl_encoder_new": (C.c_int, [_ECP, _VPP]),
    "isl_encoder_free": (None, [_VP]),
    "isl_encoder_dimension": (C.c_uint32, [_VP]),
    "isl_encoder_num_parameters": (C.c_uint64, [_VP]),
    "isl_encoder_init_random": (C.c_int, [_VP, C.c_uint64, C.c_float]),
    "isl_encoder_set_parameter": (C.c_int, [_VP, C.c_char_p, f32p, C.c_uint64]),
    "isl_encoder_get_parameter": (C.c_int, [_VP, C.c_char_p, f32p, C.c_uint64]),
    "isl_encoder_embed": (C.c_int, [_VP, i32p, i32p, C.c_uint64, C.c_uint32, f32p]),
    "isl_encoder_embed_dev": (C.c_int, [_VP, _VP, _VP, C.c_uint64, C.c_uint32, _VP]),
    "isl_encoder_last_timing": (C.c_int, [_VP, f32p, C.POINTER(C.c_double)]),
    "isl_gemm_bf16_dev": (C.c_int, [_VP, _VP, C.c_uint32, C.c_uint32, C.c_uint32, _VP, _VP, C.c_int32, _VP, _VP]),
    "isl_index_set_recompute": (C.c_int, [_VP, _VP, i32p, i32p, C.c_uint32]),
    "isl_index_drop_vectors": (C.c_int, [_VP]),
    "isl_index_search_adc_recompute": (C.c_int, [_VP, f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u64p, f32p, u32p, _SSP]),
    "isl_index_search_recompute": (C.c_int, [_VP, f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u64p, f32p, u32p, _SSP]),
    "isl_index_last_recompute": (C.c_int, [_VP, u64p, f32p, f32p, f32p]),
    "isl_index_set_hub_cache": (C.c_int, [_VP, C.c_uint64]),
    "isl_index_set_rerank_limit": (C.c_int, [_VP, C.c_uint32]),
    "isl_index_hub_cache_info": (C.c_int, [_VP, u64p, u64p]),
    "isl_index_to_bytes": (C.c_int, [_VP, _VP, C.c_uint64, u64p]),
    "isl_index_from_bytes": (C.c_int, [_VP, C.c_uint64, f32p, C.c_uint32, _VPP]),
    "isl_pq_to_bytes": (C.c_int, [_VP, _VP, C.c_uint64, u64p]),
    "isl_pq_from_bytes": (C.c_int, [_VP, C.c_uint64, _VPP]),
    "isl_hnsw_to_bytes": (C.c_int, [_VP, _VP, C.c_uint64, u64p]),
    "isl_hnsw_from_bytes": (C.c_int, [_VP, C.c_uint64, _VPP]),
    "isl_index_get_config": (C.c_int, [_VP, _LCP]),
    "isl_pq_get_config": (C.c_int, [_VP, _PCP]),
    "isl_pq_dimension": (C.c_uint32, [_VP]),
    "isl_hnsw_get_config": (C.c_int, [_VP, _HCP]),
    "isl_hnsw_get_vector": (C.c_int, [_VP, C.c_uint64, f32p]),
    "isl_index_get_vector": (C.c_int, [_VP, C.c_uint64, f32p]),
    "isl_merge_topk": (C.c_int, [u64p, f32p, C.c_uint32, C.c_uint64, C.c_uint32, u64p, f32p, u32p]),
    "isl_merge_topk_dev": (C.c_int, [_VP, _VP, C.c_uint32, C.c_uint64, C.c_uint32, _VP, _VP, _VP]),
    "isl_set_caller_stream": (C.c_int, [_VP]),
    "isl_index_last_build_stats": (C.c_int, [_VP, C.POINTER(BuildStatsStruct)]),
    "isl_index_set_neighbors": (C.c_int, [_VP, C.c_uint64, u64p, C.c_uint64]),
    "isl_shard_unique_id": (C.c_int, [_VP, C.c_uint64]),
    "isl_shard_init": (C.c_int, [C.c_int, C.c_int, _VP, _VPP]),
    "isl_shard_free": (None, [_VP]),
    "isl_shard_rank": (C.c_int, [_VP]),
    "isl_shard_world": (C.c_int, [_VP]),
    "isl_shard_enable_peer_exchange": (C.c_int, [_VP, C.c_uint64]),
    "isl_index_search_sharded": (C.c_int, [_VP, _VP, C.c_uint64, f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u64p, f32p, u32p]),
    "isl_index_search_sharded_dev": (C.c_int, [_VP, _VP, C.c_uint64, _VP, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _VP, _VP, _VP]),
    "isl_index_search_sharded_adc": (C.c_int, [_VP, _VP, C.c_uint64, C.c_int32, f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u64p, f32p, u32p]),
    "isl_shard_last_timing": (C.c_int, [_VP, f32p, f32p, f32p]),
    "isl_index_search_packed_dev": (C.c_int, [_VP, C.c_uint64, _VP, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _VP]),
    "isl_merge_packed_dev": (C.c_int, [_VP, C.c_uint32, C.c_uint64, C.c_uint32, _VP, _VP, _VP]),
}

# Test hooks: exported by the library, declared in the header only under ISL_TEST_HOOKS, not part of the product ABI.
TEST_HOOKS = {
    "isl_std_rng_draw": (C.c_int, [C.c_uint64, C.POINTER(C.c_uint8), C.c_uint64, C.c_uint64, u64p]),
    "isl_adc_table_round": (C.c_int, [f32p, C.c_uint64, f32p]),
    "isl_test_raise": (C.c_int, [C.c_int32]),
}

_lib = None


def load(strict=True):
    """Load libislands_b200.so and attach signatures.  strict=False tolerates symbols that are
    declared here but missing from the library (only the ABI test uses that, to report them)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C islands_b200/csrc`. islands_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    missing = []
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    for name, (res, args) in TEST_HOOKS.items():
        fn = getattr(lib, name, None)
        if fn is not None:
            fn.restype = res
            fn.argtypes = args
    if missing and strict and not os.environ.get("ISL_DEV_ALLOW_MISSING"):
        raise ImportError(f"libislands_b200.so lacks symbols declared in islands_b200.h: {missing}")
    _lib = lib
    return lib


def last_error():
    msg = load().isl_last_error()
    return msg.decode("utf-8", "replace") if msg else ""
