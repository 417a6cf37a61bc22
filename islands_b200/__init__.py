"""islands_b200 — B200-native (sm_100a) LEANN / HNSW search hot path behind the reference's
`src/core` API.  Compute lives in lib/libislands_b200.so (C ABI: include/islands_b200.h);
this package is the host-side mirror of the reference interface.  No CPU fallback."""
from .core import (Encoder, EncoderConfig, gemm_bf16_dev, CoreError, CsrGraph, CudaError, DimensionMismatch, DistanceMetric, EmptyCollection,
                   HnswConfig, HnswGraph, HnswNode, IndexNotBuilt, InMemoryEmbeddingProvider, InvalidArgument, InvalidConfig,
                   LeannConfig, LeannIndex, NodeNotFound, PQCodebook, PQConfig, PQError, ProductQuantizer,
                   PruningStrategy, SerializationError, merge_topk, merge_topk_dev, normalize_vector, normalized,
                   random_level, to_similarity)

from .registry import IslandRegistry, StoredIndex
from .search import MultiIndexSearcher, SearchConfig, Searcher, SearchResult
from .storage import DeserializationError, FileSystemStorage, IndexMetadata, IndexReader, IndexWriter
from .files import (GraphFile, graph_file_from_csr, graph_file_from_hnsw, hubs_by_in_degree, read_codebook_file, read_codes_file,
                    read_graph_file, write_codebook_file, write_codes_file, write_graph_file)

__all__ = [n for n in dir() if not n.startswith("_")]
