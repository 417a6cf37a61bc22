//! Safe wrapper over `libislands_b200.so` with the signatures of panbanda/islands' `src/core`
//! (`src/core/mod.rs:60-99`): `LeannIndex::{new, build, search, search_with_params}`, `CsrGraph`,
//! `LeannConfig`, `EmbeddingProvider`, `HnswGraph` / `HnswConfig` / `HnswNode`, `Searcher` / `MultiIndexSearcher` /
//! `SearchConfig` / `SearchResult`, `ProductQuantizer`, `DistanceMetric` / `Distance`, `CoreError`, the `prelude`.
//! A maintainer of the reference swaps `use crate::core::leann::LeannIndex` for
//! `use islands_b200::LeannIndex` and keeps the call sites.
//!
//! There is no Rust toolchain in the image this repository is developed in: this file is source for the
//! reference's build and is not compiled by this repository's tests.  Everything below is argument marshalling;
//! the arithmetic lives in the CUDA library, and a missing GPU surfaces as `CoreError::SearchError`, never as a
//! CPU fallback.
#![allow(clippy::missing_safety_doc)]

use islands_b200_sys as sys;
use std::ffi::CStr;
use std::ptr;

// ---------------------------------------------------------------------------------------------
// error.rs:9-62
// ---------------------------------------------------------------------------------------------
#[derive(thiserror::Error, Debug)]
pub enum CoreError {
    #[error("Vector dimension mismatch: expected {expected}, got {actual}")]
    DimensionMismatch { expected: usize, actual: usize },
    #[error("Empty vector collection")]
    EmptyCollection,
    #[error("Invalid configuration: {0}")]
    InvalidConfig(String),
    #[error("Index not built")]
    IndexNotBuilt,
    #[error("Node not found: {0}")]
    NodeNotFound(u64),
    #[error("Serialization error: {0}")]
    Serialization(String),
    #[error("Deserialization error: {0}")]
    Deserialization(String),
    #[error("IO error: {0}")]
    Io(#[from] std::io::Error),
    #[error("HNSW graph error: {0}")]
    HnswError(String),
    #[error("Product quantization error: {0}")]
    PQError(String),
    /// Device failures (`ISL_CUDA_ERROR`) and ABI argument errors have no variant of their own in the reference;
    /// they travel as `SearchError` with the library's message.
    #[error("Search error: {0}")]
    SearchError(String),
    #[error("Embedding error: {0}")]
    EmbeddingError(String),
}

pub type CoreResult<T> = Result<T, CoreError>;

fn last_message() -> String {
    unsafe {
        let p = sys::isl_last_error();
        if p.is_null() {
            String::new()
        } else {
            CStr::from_ptr(p).to_string_lossy().into_owned()
        }
    }
}

/// `isl_status` -> the `CoreError` variant it mirrors, payloads included (`isl_last_error_detail`).
fn check(status: i32) -> CoreResult<()> {
    if status == sys::ISL_OK {
        return Ok(());
    }
    let msg = last_message();
    let (mut a, mut b) = (0u64, 0u64);
    unsafe { sys::isl_last_error_detail(&mut a, &mut b) };
    Err(match status {
        sys::ISL_DIM_MISMATCH => CoreError::DimensionMismatch { expected: a as usize, actual: b as usize },
        sys::ISL_EMPTY_COLLECTION => CoreError::EmptyCollection,
        sys::ISL_INVALID_CONFIG => CoreError::InvalidConfig(msg),
        sys::ISL_INDEX_NOT_BUILT => CoreError::IndexNotBuilt,
        sys::ISL_NODE_NOT_FOUND => CoreError::NodeNotFound(a),
        sys::ISL_PQ_ERROR => CoreError::PQError(msg),
        sys::ISL_SERIALIZATION => CoreError::Deserialization(msg),
        _ => CoreError::SearchError(msg),
    })
}

// ---------------------------------------------------------------------------------------------
// distance.rs:7-139
// ---------------------------------------------------------------------------------------------
#[derive(Debug, Clone, Copy, PartialEq, Eq, Default)]
pub enum DistanceMetric {
    #[default]
    Cosine,
    Euclidean,
    DotProduct,
    Manhattan,
}

impl DistanceMetric {
    fn code(self) -> i32 {
        match self {
            DistanceMetric::Cosine => sys::ISL_METRIC_COSINE,
            DistanceMetric::Euclidean => sys::ISL_METRIC_EUCLIDEAN,
            DistanceMetric::DotProduct => sys::ISL_METRIC_DOT,
            DistanceMetric::Manhattan => sys::ISL_METRIC_MANHATTAN,
        }
    }
    fn from_code(c: i32) -> Self {
        match c {
            sys::ISL_METRIC_EUCLIDEAN => DistanceMetric::Euclidean,
            sys::ISL_METRIC_DOT => DistanceMetric::DotProduct,
            sys::ISL_METRIC_MANHATTAN => DistanceMetric::Manhattan,
            _ => DistanceMetric::Cosine,
        }
    }
}

/// distance.rs:22-35
pub trait Distance: Send + Sync {
    fn calculate(&self, a: &[f32], b: &[f32]) -> CoreResult<f32>;
    fn calculate_squared(&self, a: &[f32], b: &[f32]) -> CoreResult<f32>;
    fn batch_calculate(&self, query: &[f32], vectors: &[&[f32]]) -> CoreResult<Vec<f32>>;
}

impl Distance for DistanceMetric {
    fn calculate(&self, a: &[f32], b: &[f32]) -> CoreResult<f32> {
        let mut out = 0f32;
        check(unsafe { sys::isl_distance_calculate(self.code(), a.as_ptr(), a.len() as u64, b.as_ptr(), b.len() as u64, &mut out) })?;
        Ok(out)
    }
    fn calculate_squared(&self, a: &[f32], b: &[f32]) -> CoreResult<f32> {
        let mut out = 0f32;
        check(unsafe {
            sys::isl_distance_calculate_squared(self.code(), a.as_ptr(), a.len() as u64, b.as_ptr(), b.len() as u64, &mut out)
        })?;
        Ok(out)
    }
    fn batch_calculate(&self, query: &[f32], vectors: &[&[f32]]) -> CoreResult<Vec<f32>> {
        let d = query.len();
        let mut rows = Vec::with_capacity(vectors.len() * d);
        for v in vectors {
            if v.len() != d {
                return Err(CoreError::DimensionMismatch { expected: d, actual: v.len() });
            }
            rows.extend_from_slice(v);
        }
        let mut out = vec![0f32; vectors.len()];
        check(unsafe {
            sys::isl_distance_batch(self.code(), query.as_ptr(), rows.as_ptr(), vectors.len() as u64, d as u32, out.as_mut_ptr())
        })?;
        Ok(out)
    }
}

/// distance.rs:125-132
pub fn normalize_vector(v: &mut [f32]) {
    let _ = unsafe { sys::isl_normalize_rows(v.as_mut_ptr(), 1, v.len() as u32) };
}

/// distance.rs:135-139
pub fn normalized(v: &[f32]) -> Vec<f32> {
    let mut out = v.to_vec();
    normalize_vector(&mut out);
    out
}

// ---------------------------------------------------------------------------------------------
// leann.rs:82-159 — the recompute seam
// ---------------------------------------------------------------------------------------------
pub trait EmbeddingProvider: Send + Sync {
    fn compute_embedding(&self, id: u64) -> CoreResult<Vec<f32>>;
    fn compute_embeddings_batch(&self, ids: &[u64]) -> CoreResult<Vec<Vec<f32>>> {
        ids.iter().map(|&id| self.compute_embedding(id)).collect()
    }
    fn dimension(&self) -> usize;
}

pub struct InMemoryEmbeddingProvider {
    embeddings: Vec<Vec<f32>>,
    dimension: usize,
}

impl InMemoryEmbeddingProvider {
    pub fn new(embeddings: Vec<Vec<f32>>) -> CoreResult<Self> {
        if embeddings.is_empty() {
            return Err(CoreError::EmptyCollection);
        }
        let dimension = embeddings[0].len();
        Ok(Self { embeddings, dimension })
    }
    pub fn with_dimension(dimension: usize) -> Self {
        Self { embeddings: Vec::new(), dimension }
    }
    pub fn add(&mut self, embedding: Vec<f32>) -> CoreResult<u64> {
        if embedding.len() != self.dimension {
            return Err(CoreError::DimensionMismatch { expected: self.dimension, actual: embedding.len() });
        }
        self.embeddings.push(embedding);
        Ok((self.embeddings.len() - 1) as u64)
    }
}

impl EmbeddingProvider for InMemoryEmbeddingProvider {
    fn compute_embedding(&self, id: u64) -> CoreResult<Vec<f32>> {
        self.embeddings.get(id as usize).cloned().ok_or(CoreError::NodeNotFound(id))
    }
    fn dimension(&self) -> usize {
        self.dimension
    }
}

// ---------------------------------------------------------------------------------------------
// leann.rs:168-461 — PruningStrategy, CsrGraph, LeannConfig
// ---------------------------------------------------------------------------------------------
#[derive(Debug, Clone, Copy, PartialEq, Eq, Default)]
pub enum PruningStrategy {
    #[default]
    Global,
    Local,
    /// Draws from `thread_rng` in the reference (leann.rs:1043); the GPU library draws from a counter stream seeded by
    /// `LeannConfig::prune_seed` (include/islands_b200.h).
    Proportional,
}

#[derive(Debug, Clone, Default)]
pub struct CsrGraph {
    pub node_offsets: Vec<usize>,
    pub neighbors: Vec<u64>,
    pub levels: Vec<usize>,
    pub entry_point: Option<u64>,
    pub max_level: usize,
    pub num_nodes: usize,
    pub degree_counts: Vec<usize>,
}

impl CsrGraph {
    pub fn new() -> Self {
        Self { node_offsets: vec![0], ..Default::default() }
    }
    pub fn get_neighbors(&self, node_id: u64) -> Option<&[u64]> {
        let id = node_id as usize;
        if id >= self.num_nodes {
            return None;
        }
        Some(&self.neighbors[self.node_offsets[id]..self.node_offsets[id + 1]])
    }
    pub fn add_node(&mut self, neighbors: Vec<u64>, level: usize) -> u64 {
        let id = self.num_nodes as u64;
        self.num_nodes += 1;
        self.levels.push(level);
        self.degree_counts.push(neighbors.len());
        self.neighbors.extend(neighbors);
        self.node_offsets.push(self.neighbors.len());
        if self.entry_point.is_none() || level > self.max_level {
            self.entry_point = Some(id);
            self.max_level = level;
        }
        id
    }
    pub fn set_neighbors(&mut self, node_id: u64, new_neighbors: Vec<u64>) {
        let id = node_id as usize;
        if id >= self.num_nodes {
            return;
        }
        let (s, e) = (self.node_offsets[id], self.node_offsets[id + 1]);
        if new_neighbors.len() == e - s {
            self.neighbors[s..e].copy_from_slice(&new_neighbors);
        } else {
            let delta = new_neighbors.len() as isize - (e - s) as isize;
            self.neighbors.splice(s..e, new_neighbors.iter().copied());
            for o in self.node_offsets[id + 1..].iter_mut() {
                *o = (*o as isize + delta) as usize;
            }
        }
        self.degree_counts[id] = new_neighbors.len();
    }
    pub fn storage_bytes(&self) -> usize {
        8 * (self.node_offsets.len() + self.neighbors.len() + self.levels.len() + self.degree_counts.len())
    }
}

#[derive(Debug, Clone)]
pub struct LeannConfig {
    pub m: usize,
    pub m0: usize,
    pub ef_construction: usize,
    pub ml: f64,
    pub max_layers: usize,
    pub metric: DistanceMetric,
    pub ef_search: usize,
    pub beam_width: usize,
    pub prune_ratio: f32,
    pub pruning_strategy: PruningStrategy,
    pub high_degree_pruning: bool,
    pub hub_percentile: f32,
    pub is_compact: bool,
    pub is_recompute: bool,
    /// Not a reference field: seed of the `PruningStrategy::Proportional` draw stream.
    pub prune_seed: u64,
}

impl Default for LeannConfig {
    fn default() -> Self {
        Self::paper_default()
    }
}

impl LeannConfig {
    fn from_raw(c: &sys::IslLeannConfig) -> Self {
        Self {
            m: c.m as usize,
            m0: c.m0 as usize,
            ef_construction: c.ef_construction as usize,
            ml: c.ml,
            max_layers: c.max_layers as usize,
            metric: DistanceMetric::from_code(c.metric),
            ef_search: c.ef_search as usize,
            beam_width: c.beam_width as usize,
            prune_ratio: c.prune_ratio,
            pruning_strategy: match c.pruning_strategy {
                sys::ISL_PRUNE_LOCAL => PruningStrategy::Local,
                sys::ISL_PRUNE_PROPORTIONAL => PruningStrategy::Proportional,
                _ => PruningStrategy::Global,
            },
            high_degree_pruning: c.high_degree_pruning != 0,
            hub_percentile: c.hub_percentile,
            is_compact: c.is_compact != 0,
            is_recompute: c.is_recompute != 0,
            prune_seed: c.prune_seed,
        }
    }
    fn raw(&self) -> sys::IslLeannConfig {
        sys::IslLeannConfig {
            m: self.m as u64,
            m0: self.m0 as u64,
            ef_construction: self.ef_construction as u64,
            ml: self.ml,
            max_layers: self.max_layers as u64,
            metric: self.metric.code(),
            ef_search: self.ef_search as u64,
            beam_width: self.beam_width as u64,
            prune_ratio: self.prune_ratio,
            pruning_strategy: match self.pruning_strategy {
                PruningStrategy::Global => sys::ISL_PRUNE_GLOBAL,
                PruningStrategy::Local => sys::ISL_PRUNE_LOCAL,
                PruningStrategy::Proportional => sys::ISL_PRUNE_PROPORTIONAL,
            },
            high_degree_pruning: self.high_degree_pruning as i32,
            hub_percentile: self.hub_percentile,
            is_compact: self.is_compact as i32,
            is_recompute: self.is_recompute as i32,
            prune_seed: self.prune_seed,
        }
    }
    fn preset(f: unsafe extern "C" fn(*mut sys::IslLeannConfig) -> i32) -> Self {
        let mut raw = std::mem::MaybeUninit::<sys::IslLeannConfig>::zeroed();
        unsafe {
            f(raw.as_mut_ptr());
            Self::from_raw(&raw.assume_init())
        }
    }
    /// leann.rs:386-403
    pub fn paper_default() -> Self {
        Self::preset(sys::isl_leann_config_default)
    }
    /// leann.rs:406-416
    pub fn fast() -> Self {
        Self::preset(sys::isl_leann_config_fast)
    }
    /// leann.rs:419-429
    pub fn accurate() -> Self {
        Self::preset(sys::isl_leann_config_accurate)
    }
    /// leann.rs:432-460
    pub fn validate(&self) -> CoreResult<()> {
        check(unsafe { sys::isl_leann_config_validate(&self.raw()) })
    }
}

// ---------------------------------------------------------------------------------------------
// leann.rs:493-1067 — LeannIndex
// ---------------------------------------------------------------------------------------------
pub struct LeannIndex {
    config: LeannConfig,
    handle: *mut sys::IslIndex,
    // The library keeps raw pointers to an attached quantizer / encoder: the index shares their ownership so that
    // neither can be freed while it may still be read (fields drop after `Drop::drop` has freed the handle).
    pq: Option<std::sync::Arc<ProductQuantizer>>,
    encoder: Option<std::sync::Arc<Encoder>>,
}

// The handle is immutable after build; searches lease per-call scratch inside the library (csrc/api.cu).
unsafe impl Send for LeannIndex {}
unsafe impl Sync for LeannIndex {}

impl Drop for LeannIndex {
    fn drop(&mut self) {
        unsafe { sys::isl_index_free(self.handle) };
    }
}

/// Levels are drawn from `thread_rng` in the reference (leann.rs:549-554); here from a seed (same formula).
pub struct BuildOptions {
    pub levels: Option<Vec<u64>>,
    pub seed: u64,
    /// 1 = the reference's sequential insertion order; > 1 = rounds of that many inserts (GPU-parallel).
    pub batch: u32,
}

impl Default for BuildOptions {
    fn default() -> Self {
        Self { levels: None, seed: 0, batch: 4096 }
    }
}

impl LeannIndex {
    pub fn new(config: LeannConfig) -> CoreResult<Self> {
        config.validate()?;
        Ok(Self { config, handle: ptr::null_mut(), pq: None, encoder: None })
    }
    pub fn with_defaults() -> CoreResult<Self> {
        Self::new(LeannConfig::default())
    }
    pub fn len(&self) -> usize {
        unsafe { sys::isl_index_len(self.handle) as usize }
    }
    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }
    pub fn dimension(&self) -> Option<usize> {
        match unsafe { sys::isl_index_dimension(self.handle) } {
            0 => None,
            d => Some(d as usize),
        }
    }
    pub fn storage_bytes(&self) -> usize {
        if self.handle.is_null() {
            8
        } else {
            unsafe { sys::isl_index_storage_bytes(self.handle) as usize }
        }
    }
    pub fn is_recompute(&self) -> bool {
        self.config.is_recompute
    }
    pub fn is_compact(&self) -> bool {
        self.config.is_compact
    }
    pub fn config(&self) -> &LeannConfig {
        &self.config
    }

    /// leann.rs:560-631.  The provider is drained once (`compute_embeddings_batch(0..n)`) and the embeddings become
    /// resident in HBM; construction then runs on the GPU.
    pub fn build<P: EmbeddingProvider>(&mut self, provider: &P, num_vectors: usize) -> CoreResult<()> {
        self.build_with(provider, num_vectors, &BuildOptions::default())
    }

    pub fn build_with<P: EmbeddingProvider>(&mut self, provider: &P, num_vectors: usize, opt: &BuildOptions) -> CoreResult<()> {
        if num_vectors == 0 {
            return Ok(()); // leann.rs:565-567
        }
        let ids: Vec<u64> = (0..num_vectors as u64).collect();
        let rows = provider.compute_embeddings_batch(&ids)?;
        let dim = provider.dimension();
        let mut flat = Vec::with_capacity(num_vectors * dim);
        for r in &rows {
            if r.len() != dim {
                return Err(CoreError::DimensionMismatch { expected: dim, actual: r.len() });
            }
            flat.extend_from_slice(r);
        }
        if let Some(l) = &opt.levels {
            if l.len() != num_vectors {
                return Err(CoreError::InvalidConfig("levels must have one entry per vector".into()));
            }
        }
        let mut h: *mut sys::IslIndex = ptr::null_mut();
        check(unsafe {
            sys::isl_index_build(
                &self.config.raw(),
                dim as u32,
                num_vectors as u64,
                flat.as_ptr(),
                opt.levels.as_ref().map_or(ptr::null(), |l| l.as_ptr()),
                opt.seed,
                opt.batch,
                &mut h,
            )
        })?;
        unsafe { sys::isl_index_free(self.handle) };
        self.handle = h;
        self.pq = None; // attachments belonged to the handle that was just freed
        self.encoder = None;
        Ok(())
    }

    /// Adopt an existing graph ("identical graphs" entry point of the parity tests).
    pub fn from_csr<P: EmbeddingProvider>(config: LeannConfig, graph: &CsrGraph, provider: &P) -> CoreResult<Self> {
        config.validate()?;
        let n = graph.num_nodes;
        let ids: Vec<u64> = (0..n as u64).collect();
        let rows = provider.compute_embeddings_batch(&ids)?;
        let dim = provider.dimension();
        let flat: Vec<f32> = rows.into_iter().flatten().collect();
        let off: Vec<u64> = graph.node_offsets.iter().map(|&x| x as u64).collect();
        let lv: Vec<u64> = graph.levels.iter().map(|&x| x as u64).collect();
        let mut h: *mut sys::IslIndex = ptr::null_mut();
        check(unsafe {
            sys::isl_index_from_csr(
                &config.raw(),
                dim as u32,
                n as u64,
                off.as_ptr(),
                graph.neighbors.as_ptr(),
                lv.as_ptr(),
                graph.entry_point.map_or(-1, |e| e as i64),
                flat.as_ptr(),
                &mut h,
            )
        })?;
        Ok(Self { config, handle: h, pq: None, encoder: None })
    }

    /// graph: the CSR arrays in the reference's layout.
    pub fn graph(&self) -> CoreResult<CsrGraph> {
        let n = self.len();
        let e = unsafe { sys::isl_index_num_edges(self.handle) } as usize;
        let (mut off, mut nb, mut lv, mut deg) = (vec![0u64; n + 1], vec![0u64; e], vec![0u64; n], vec![0u64; n]);
        check(unsafe { sys::isl_index_export_csr(self.handle, off.as_mut_ptr(), nb.as_mut_ptr(), lv.as_mut_ptr(), deg.as_mut_ptr()) })?;
        let ep = unsafe { sys::isl_index_entry_point(self.handle) };
        Ok(CsrGraph {
            node_offsets: off.into_iter().map(|x| x as usize).collect(),
            neighbors: nb,
            levels: lv.into_iter().map(|x| x as usize).collect(),
            entry_point: if ep < 0 { None } else { Some(ep as u64) },
            max_level: unsafe { sys::isl_index_max_level(self.handle) } as usize,
            num_nodes: n,
            degree_counts: deg.into_iter().map(|x| x as usize).collect(),
        })
    }

    /// leann.rs:858-865.  The provider argument is kept for signature compatibility: the embeddings it returned at
    /// build time are resident on the device (for true on-demand recompute attach an encoder: `isl_index_set_recompute`).
    pub fn search<P: EmbeddingProvider>(&self, query: &[f32], k: usize, provider: &P) -> CoreResult<Vec<(u64, f32)>> {
        self.search_with_params(query, k, self.config.ef_search, provider)
    }

    /// leann.rs:868-896
    pub fn search_with_params<P: EmbeddingProvider>(&self, query: &[f32], k: usize, ef: usize, _provider: &P) -> CoreResult<Vec<(u64, f32)>> {
        if self.is_empty() {
            return Ok(vec![]);
        }
        let (ids, dist, count) = self.search_batch(query, 1, k, ef)?;
        Ok((0..count[0] as usize).map(|i| (ids[i], dist[i])).collect())
    }

    /// Batched form (what a GPU wants): `queries` is `[nq][dim]` row-major; returns ids / distances `[nq][k]`
    /// (padded with `u64::MAX` / `+inf`) and the result count per query.
    pub fn search_batch(&self, queries: &[f32], nq: usize, k: usize, ef: usize) -> CoreResult<(Vec<u64>, Vec<f32>, Vec<u32>)> {
        let dim = if nq == 0 { 0 } else { queries.len() / nq };
        let (mut ids, mut dist, mut count) = (vec![u64::MAX; nq * k], vec![f32::INFINITY; nq * k], vec![0u32; nq]);
        check(unsafe {
            sys::isl_index_search(
                self.handle,
                queries.as_ptr(),
                nq as u64,
                dim as u32,
                k as u32,
                ef as u32,
                ids.as_mut_ptr(),
                dist.as_mut_ptr(),
                count.as_mut_ptr(),
                ptr::null_mut(),
            )
        })?;
        Ok((ids, dist, count))
    }

    /// Sharded search (service.rs:777-801): collective over the ranks of `shard`; `id_base` = first global id of
    /// this rank's shard.  Every rank passes the same queries and receives the same merged top-k.
    pub fn search_sharded(&self, shard: &ShardComm, id_base: u64, queries: &[f32], nq: usize, k: usize, ef: usize)
                          -> CoreResult<(Vec<u64>, Vec<f32>, Vec<u32>)> {
        let dim = if nq == 0 { 0 } else { queries.len() / nq };
        let (mut ids, mut dist, mut count) = (vec![u64::MAX; nq * k], vec![f32::INFINITY; nq * k], vec![0u32; nq]);
        check(unsafe {
            sys::isl_index_search_sharded(self.handle, shard.handle, id_base, queries.as_ptr(), nq as u64, dim as u32, k as u32,
                                          ef as u32, ids.as_mut_ptr(), dist.as_mut_ptr(), count.as_mut_ptr())
        })?;
        Ok((ids, dist, count))
    }

    /// leann.rs:1059-1061 (graph only, bincode layout)
    pub fn to_bytes(&self) -> CoreResult<Vec<u8>> {
        let mut len = 0u64;
        check(unsafe { sys::isl_index_to_bytes(self.handle, ptr::null_mut(), 0, &mut len) })
            .map_err(|e| CoreError::Serialization(e.to_string()))?;
        let mut out = vec![0u8; len as usize];
        check(unsafe { sys::isl_index_to_bytes(self.handle, out.as_mut_ptr(), len, &mut len) })
            .map_err(|e| CoreError::Serialization(e.to_string()))?;
        Ok(out)
    }

    /// leann.rs:1064-1066 + the embeddings the provider returns (the bytes hold the graph only).
    pub fn from_bytes<P: EmbeddingProvider>(bytes: &[u8], provider: &P, num_vectors: usize) -> CoreResult<Self> {
        let ids: Vec<u64> = (0..num_vectors as u64).collect();
        let flat: Vec<f32> = provider.compute_embeddings_batch(&ids)?.into_iter().flatten().collect();
        let mut h: *mut sys::IslIndex = ptr::null_mut();
        check(unsafe { sys::isl_index_from_bytes(bytes.as_ptr(), bytes.len() as u64, flat.as_ptr(), provider.dimension() as u32, &mut h) })?;
        let mut raw = std::mem::MaybeUninit::<sys::IslLeannConfig>::zeroed();
        check(unsafe { sys::isl_index_get_config(h, raw.as_mut_ptr()) })?;
        Ok(Self { config: LeannConfig::from_raw(unsafe { &raw.assume_init() }), handle: h, pq: None, encoder: None })
    }
}

// ---------------------------------------------------------------------------------------------
// embedding/candle_provider.rs:353-507 — the recompute encoder, and the searches that use it as the
// EmbeddingProvider (leann.rs:82-99, 947-950)
// ---------------------------------------------------------------------------------------------
/// BERT shape of the recompute encoder (defaults: BERT-base, 110M parameters) and its arithmetic mode.
#[derive(Debug, Clone, Copy, PartialEq, Eq, Default)]
pub enum EncoderPrecision {
    /// bf16 operands, f32 accumulation.
    #[default]
    Bf16,
    /// Split precision (three bf16 products per f32 product): embeddings agree with an f32 forward to ~2e-5.
    Bf16x3,
}

pub struct Encoder {
    handle: *mut sys::IslEncoder,
}

// The library serialises forwards on one encoder handle (a mutex inside it); weights are set through `&mut self`.
unsafe impl Send for Encoder {}
unsafe impl Sync for Encoder {}

impl Drop for Encoder {
    fn drop(&mut self) {
        unsafe { sys::isl_encoder_free(self.handle) };
    }
}

impl Encoder {
    /// BERT-base shape; weights are set by name (`set_parameter`, Hugging Face names) or drawn (`init_random`).
    pub fn new(precision: EncoderPrecision) -> CoreResult<Self> {
        let mut raw = std::mem::MaybeUninit::<sys::IslEncoderConfig>::zeroed();
        check(unsafe { sys::isl_encoder_config_default(raw.as_mut_ptr()) })?;
        let mut cfg = unsafe { raw.assume_init() };
        cfg.precision = match precision {
            EncoderPrecision::Bf16 => 0,
            EncoderPrecision::Bf16x3 => 1,
        };
        let mut h: *mut sys::IslEncoder = ptr::null_mut();
        check(unsafe { sys::isl_encoder_new(&cfg, &mut h) })?;
        Ok(Self { handle: h })
    }
    /// `EmbeddingProvider::dimension`
    pub fn dimension(&self) -> usize {
        unsafe { sys::isl_encoder_dimension(self.handle) as usize }
    }
    pub fn num_parameters(&self) -> u64 {
        unsafe { sys::isl_encoder_num_parameters(self.handle) }
    }
    pub fn init_random(&mut self, seed: u64, stddev: f32) -> CoreResult<()> {
        check(unsafe { sys::isl_encoder_init_random(self.handle, seed, stddev) })
    }
    pub fn set_parameter(&mut self, name: &str, data: &[f32]) -> CoreResult<()> {
        let name = std::ffi::CString::new(name).map_err(|e| CoreError::EmbeddingError(e.to_string()))?;
        check(unsafe { sys::isl_encoder_set_parameter(self.handle, name.as_ptr(), data.as_ptr(), data.len() as u64) })
    }
    /// The model half of `embed_texts_raw` (candle_provider.rs:404-507): `token_ids` is `[batch][seq_len]` zero padded,
    /// `lengths[b]` the number of real tokens; returns `[batch][dimension]` pooled, L2-normalised embeddings.
    pub fn embed(&self, token_ids: &[i32], lengths: &[i32], seq_len: usize) -> CoreResult<Vec<f32>> {
        if token_ids.len() != lengths.len() * seq_len {
            return Err(CoreError::EmbeddingError("token_ids must hold lengths.len() rows of seq_len ids".into()));
        }
        let mut out = vec![0f32; lengths.len() * self.dimension()];
        check(unsafe {
            sys::isl_encoder_embed(self.handle, token_ids.as_ptr(), lengths.as_ptr(), lengths.len() as u64, seq_len as u32, out.as_mut_ptr())
        })
        .map_err(|e| CoreError::EmbeddingError(e.to_string()))?;
        Ok(out)
    }
}

impl LeannIndex {
    /// Attach a trained quantizer and the codes `[len()][num_subquantizers]` of the indexed vectors: enables the
    /// two-level search and "PQ ADC traversal + exact rerank".  The index shares ownership of the quantizer.
    pub fn attach_pq(&mut self, pq: std::sync::Arc<ProductQuantizer>, codes: &[u16]) -> CoreResult<()> {
        if codes.len() != self.len() * pq.num_subquantizers() {
            return Err(CoreError::PQError("one code row per indexed vector".into()));
        }
        check(unsafe { sys::isl_index_attach_pq(self.handle, pq.handle, codes.as_ptr()) })?;
        self.pq = Some(pq);
        Ok(())
    }
    /// Attach the encoder as this index's `EmbeddingProvider`: node i is embedded from row i of `token_ids`
    /// (`[len()][seq_len]`) whenever a recompute search needs it.  The index shares ownership of the encoder.
    pub fn set_recompute(&mut self, encoder: std::sync::Arc<Encoder>, token_ids: &[i32], lengths: &[i32], seq_len: usize) -> CoreResult<()> {
        if lengths.len() != self.len() || token_ids.len() != lengths.len() * seq_len {
            return Err(CoreError::EmbeddingError("one token row of seq_len ids and one length per node".into()));
        }
        check(unsafe { sys::isl_index_set_recompute(self.handle, encoder.handle, token_ids.as_ptr(), lengths.as_ptr(), seq_len as u32) })?;
        self.encoder = Some(encoder);
        Ok(())
    }
    /// LEANN's storage saving: free the resident embeddings; graph, codes and token rows remain and only the
    /// recompute searches keep working.
    pub fn drop_vectors(&mut self) -> CoreResult<()> {
        check(unsafe { sys::isl_index_drop_vectors(self.handle) })
    }
    /// Keep the embeddings of the `count` highest in-degree nodes resident (docs/leann-specification.md:661-690).
    pub fn set_hub_cache(&mut self, count: u64) -> CoreResult<()> {
        check(unsafe { sys::isl_index_set_hub_cache(self.handle, count) })
    }
    /// Recompute / rerank only the best `limit` survivors of the ADC traversal (0 = all of them).
    pub fn set_rerank_limit(&mut self, limit: u32) -> CoreResult<()> {
        check(unsafe { sys::isl_index_set_rerank_limit(self.handle, limit) })
    }

    fn batched(
        &self,
        queries: &[f32],
        nq: usize,
        k: usize,
        ef: usize,
        call: unsafe extern "C" fn(*const sys::IslIndex, *const f32, u64, u32, u32, u32, *mut u64, *mut f32, *mut u32, *mut sys::IslSearchStats) -> i32,
    ) -> CoreResult<(Vec<u64>, Vec<f32>, Vec<u32>)> {
        let dim = if nq == 0 { 0 } else { queries.len() / nq };
        let (mut ids, mut dist, mut count) = (vec![u64::MAX; nq * k], vec![f32::INFINITY; nq * k], vec![0u32; nq]);
        check(unsafe {
            call(self.handle, queries.as_ptr(), nq as u64, dim as u32, k as u32, ef as u32, ids.as_mut_ptr(), dist.as_mut_ptr(),
                 count.as_mut_ptr(), ptr::null_mut())
        })?;
        Ok((ids, dist, count))
    }
    /// `search_with_params` exactly as the reference runs it with a model behind the provider (leann.rs:899-988): every
    /// hop's unvisited neighbours are embedded by the encoder (one pass per frontier of the whole batch).
    pub fn search_recompute_batch(&self, queries: &[f32], nq: usize, k: usize, ef: usize) -> CoreResult<(Vec<u64>, Vec<f32>, Vec<u32>)> {
        self.batched(queries, nq, k, ef, sys::isl_index_search_recompute)
    }
    /// The cheap variant: ADC traversal, encoder over the batch's distinct survivors, exact rerank.
    pub fn search_adc_recompute_batch(&self, queries: &[f32], nq: usize, k: usize, ef: usize) -> CoreResult<(Vec<u64>, Vec<f32>, Vec<u32>)> {
        self.batched(queries, nq, k, ef, sys::isl_index_search_adc_recompute)
    }
    /// "PQ ADC traversal + exact rerank" over the resident embeddings (include/islands_b200.h).
    pub fn search_adc_rerank_batch(&self, queries: &[f32], nq: usize, k: usize, ef: usize) -> CoreResult<(Vec<u64>, Vec<f32>, Vec<u32>)> {
        self.batched(queries, nq, k, ef, sys::isl_index_search_adc_rerank)
    }
}

/// One rank's membership in a sharded index (NCCL communicator inside the library).
pub struct ShardComm {
    handle: *mut sys::IslShard,
}

unsafe impl Send for ShardComm {}

impl ShardComm {
    /// `ncclGetUniqueId`: call on one rank, hand the 128 bytes to the others by any host channel.
    pub fn unique_id() -> CoreResult<[u8; 128]> {
        let mut id = [0u8; 128];
        check(unsafe { sys::isl_shard_unique_id(id.as_mut_ptr() as *mut _, 128) })?;
        Ok(id)
    }
    pub fn new(rank: i32, world: i32, unique_id: &[u8; 128]) -> CoreResult<Self> {
        let mut h: *mut sys::IslShard = ptr::null_mut();
        check(unsafe { sys::isl_shard_init(rank, world, unique_id.as_ptr() as *const _, &mut h) })?;
        Ok(Self { handle: h })
    }
    /// Exchange by peer stores over NVLink instead of `ncclAllGather` (collective; ranks of one node).
    pub fn enable_peer_exchange(&mut self, max_records: u64) -> CoreResult<()> {
        check(unsafe { sys::isl_shard_enable_peer_exchange(self.handle, max_records) })
    }
}

impl Drop for ShardComm {
    fn drop(&mut self) {
        unsafe { sys::isl_shard_free(self.handle) };
    }
}

// ---------------------------------------------------------------------------------------------
// hnsw.rs:15-125 — HnswConfig, HnswNode
// ---------------------------------------------------------------------------------------------
#[derive(Debug, Clone)]
pub struct HnswConfig {
    pub m: usize,
    pub m0: usize,
    pub ef_construction: usize,
    pub ml: f64,
    pub metric: DistanceMetric,
    pub max_layers: usize,
}

impl Default for HnswConfig {
    /// hnsw.rs:37-48 (the values come from the library: `isl_hnsw_config_default`).
    fn default() -> Self {
        let mut raw = std::mem::MaybeUninit::<sys::IslHnswConfig>::zeroed();
        unsafe {
            sys::isl_hnsw_config_default(raw.as_mut_ptr());
            Self::from_raw(&raw.assume_init())
        }
    }
}

impl HnswConfig {
    fn from_raw(c: &sys::IslHnswConfig) -> Self {
        Self {
            m: c.m as usize,
            m0: c.m0 as usize,
            ef_construction: c.ef_construction as usize,
            ml: c.ml,
            metric: DistanceMetric::from_code(c.metric),
            max_layers: c.max_layers as usize,
        }
    }
    fn raw(&self) -> sys::IslHnswConfig {
        sys::IslHnswConfig {
            m: self.m as u64,
            m0: self.m0 as u64,
            ef_construction: self.ef_construction as u64,
            ml: self.ml,
            metric: self.metric.code(),
            max_layers: self.max_layers as u64,
        }
    }
    /// hnsw.rs:30-35: the reference's constructor drops its argument and returns the defaults; kept.
    pub fn new(_config: Self) -> Self {
        Self::default()
    }
    /// hnsw.rs:52-59
    pub fn fast() -> Self {
        Self { m: 12, m0: 24, ef_construction: 100, ..Self::default() }
    }
    /// hnsw.rs:62-69
    pub fn accurate() -> Self {
        Self { m: 32, m0: 64, ef_construction: 400, ..Self::default() }
    }
    /// hnsw.rs:72-85
    pub fn validate(&self) -> CoreResult<()> {
        check(unsafe { sys::isl_hnsw_config_validate(&self.raw()) })
    }
}

/// hnsw.rs:88-125.  The graph lives in HBM: a node is materialised on request (`HnswGraph::get_node` returns it by
/// value, where the reference hands out a borrow of its `HashMap` entry).
#[derive(Debug, Clone)]
pub struct HnswNode {
    pub id: u64,
    pub vector: Vec<f32>,
    pub connections: Vec<Vec<u64>>,
    pub level: usize,
}

impl HnswNode {
    pub fn neighbors_at(&self, layer: usize) -> Option<&[u64]> {
        self.connections.get(layer).map(Vec::as_slice)
    }
}

// ---------------------------------------------------------------------------------------------
// hnsw.rs:151-515 — HnswGraph (what `IndexerService` stores per island: service.rs:622, 655-657, 781-785)
// ---------------------------------------------------------------------------------------------
pub struct HnswGraph {
    config: HnswConfig,
    handle: *mut sys::IslHnsw,
    /// The reference draws node levels from `thread_rng` (hnsw.rs:206-211); the library applies the same formula to
    /// draw `node id` of a counter stream with this seed.
    level_seed: u64,
}

// Searches lease per-call scratch inside the library; inserts take `&mut self`.
unsafe impl Send for HnswGraph {}
unsafe impl Sync for HnswGraph {}

impl Drop for HnswGraph {
    fn drop(&mut self) {
        unsafe { sys::isl_hnsw_free(self.handle) };
    }
}

impl HnswGraph {
    /// hnsw.rs:167-178
    pub fn new(config: HnswConfig) -> CoreResult<Self> {
        config.validate()?;
        let mut h: *mut sys::IslHnsw = ptr::null_mut();
        check(unsafe { sys::isl_hnsw_new(&config.raw(), &mut h) })?;
        Ok(Self { config, handle: h, level_seed: 0 })
    }
    pub fn with_defaults() -> CoreResult<Self> {
        Self::new(HnswConfig::default())
    }
    pub fn with_level_seed(mut self, seed: u64) -> Self {
        self.level_seed = seed;
        self
    }
    pub fn config(&self) -> &HnswConfig {
        &self.config
    }
    pub fn len(&self) -> usize {
        unsafe { sys::isl_hnsw_len(self.handle) as usize }
    }
    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }
    pub fn dimension(&self) -> Option<usize> {
        match unsafe { sys::isl_hnsw_dimension(self.handle) } {
            0 => None,
            d => Some(d as usize),
        }
    }
    pub fn entry_point(&self) -> Option<u64> {
        let ep = unsafe { sys::isl_hnsw_entry_point(self.handle) };
        if ep < 0 {
            None
        } else {
            Some(ep as u64)
        }
    }
    pub fn max_level(&self) -> usize {
        unsafe { sys::isl_hnsw_max_level(self.handle) as usize }
    }

    /// hnsw.rs:201-203: vector, level and the neighbour list of every layer the node reaches, copied from the device.
    pub fn get_node(&self, id: u64) -> Option<HnswNode> {
        let dim = self.dimension()?;
        let mut level = 0u64;
        check(unsafe { sys::isl_hnsw_node_level(self.handle, id, &mut level) }).ok()?;
        let mut vector = vec![0f32; dim];
        check(unsafe { sys::isl_hnsw_get_vector(self.handle, id, vector.as_mut_ptr()) }).ok()?;
        let cap = self.config.m0.max(self.config.m);
        let mut connections = Vec::with_capacity(level as usize + 1);
        for layer in 0..=level {
            let mut list = vec![0u64; cap];
            let mut count = 0u64;
            check(unsafe { sys::isl_hnsw_get_neighbors(self.handle, id, layer, list.as_mut_ptr(), cap as u64, &mut count) }).ok()?;
            list.truncate(count as usize);
            connections.push(list);
        }
        Some(HnswNode { id, vector, connections, level: level as usize })
    }

    /// hnsw.rs:214-250: one sequential insert (`batch = 1` is the reference's loop).
    pub fn insert(&mut self, vector: Vec<f32>) -> CoreResult<u64> {
        self.insert_batch(&vector, 1, 1)
    }

    /// `count` inserts of `[count][dim]` row-major vectors in one call; `batch > 1` inserts rounds of that many
    /// nodes against one graph snapshot (GPU-parallel construction).  Returns the id of the first new node.
    pub fn insert_batch(&mut self, vectors: &[f32], count: usize, batch: u32) -> CoreResult<u64> {
        if count == 0 {
            return Ok(self.len() as u64);
        }
        let dim = vectors.len() / count;
        let mut first = 0u64;
        check(unsafe {
            sys::isl_hnsw_insert_batch(self.handle, vectors.as_ptr(), count as u64, dim as u32, ptr::null(), self.level_seed, batch, &mut first)
        })?;
        Ok(first)
    }

    /// hnsw.rs:458-504
    pub fn search(&self, query: &[f32], k: usize, ef: usize) -> CoreResult<Vec<(u64, f32)>> {
        if self.is_empty() {
            return Ok(vec![]); // hnsw.rs:459-461
        }
        let (ids, dist, count) = self.search_batch(query, 1, k, ef)?;
        Ok((0..count[0] as usize).map(|i| (ids[i], dist[i])).collect())
    }

    /// Batched form: `queries` is `[nq][dim]` row-major; ids / distances `[nq][k]` (padded with `u64::MAX` / `+inf`)
    /// and the result count per query.
    pub fn search_batch(&self, queries: &[f32], nq: usize, k: usize, ef: usize) -> CoreResult<(Vec<u64>, Vec<f32>, Vec<u32>)> {
        let dim = if nq == 0 { 0 } else { queries.len() / nq };
        let (mut ids, mut dist, mut count) = (vec![u64::MAX; nq * k], vec![f32::INFINITY; nq * k], vec![0u32; nq]);
        if nq == 0 || self.is_empty() {
            return Ok((ids, dist, count));
        }
        check(unsafe {
            sys::isl_hnsw_search(self.handle, queries.as_ptr(), nq as u64, dim as u32, k as u32, ef as u32, ids.as_mut_ptr(),
                                 dist.as_mut_ptr(), count.as_mut_ptr())
        })?;
        Ok((ids, dist, count))
    }

    /// hnsw.rs:507-509 (bincode layout)
    pub fn to_bytes(&self) -> CoreResult<Vec<u8>> {
        let mut len = 0u64;
        check(unsafe { sys::isl_hnsw_to_bytes(self.handle, ptr::null_mut(), 0, &mut len) })
            .map_err(|e| CoreError::Serialization(e.to_string()))?;
        let mut out = vec![0u8; len as usize];
        check(unsafe { sys::isl_hnsw_to_bytes(self.handle, out.as_mut_ptr(), len, &mut len) })
            .map_err(|e| CoreError::Serialization(e.to_string()))?;
        Ok(out)
    }

    /// hnsw.rs:512-514
    pub fn from_bytes(bytes: &[u8]) -> CoreResult<Self> {
        let mut h: *mut sys::IslHnsw = ptr::null_mut();
        check(unsafe { sys::isl_hnsw_from_bytes(bytes.as_ptr(), bytes.len() as u64, &mut h) })?;
        let mut raw = std::mem::MaybeUninit::<sys::IslHnswConfig>::zeroed();
        let got = check(unsafe { sys::isl_hnsw_get_config(h, raw.as_mut_ptr()) });
        if let Err(e) = got {
            unsafe { sys::isl_hnsw_free(h) };
            return Err(e);
        }
        Ok(Self { config: HnswConfig::from_raw(unsafe { &raw.assume_init() }), handle: h, level_seed: 0 })
    }
}

/// mod.rs:79-85 (compatibility aliases)
pub type Index = HnswGraph;
pub type IndexConfig = HnswConfig;
pub type IndexBuilder = HnswConfig;
pub type Error = CoreError;

// ---------------------------------------------------------------------------------------------
// search.rs:9-249 — SearchConfig, SearchResult, Searcher, MultiIndexSearcher
// ---------------------------------------------------------------------------------------------
#[derive(Debug, Clone)]
pub struct SearchConfig {
    pub top_k: usize,
    pub ef: usize,
    pub include_vectors: bool,
    pub include_metadata: bool,
    pub min_similarity: Option<f32>,
}

impl Default for SearchConfig {
    fn default() -> Self {
        Self { top_k: 10, ef: 100, include_vectors: false, include_metadata: true, min_similarity: None }
    }
}

impl SearchConfig {
    /// search.rs:36-42
    pub fn fast(k: usize) -> Self {
        Self { top_k: k, ef: k * 2, ..Self::default() }
    }
    /// search.rs:45-51
    pub fn accurate(k: usize) -> Self {
        Self { top_k: k, ef: k * 10, ..Self::default() }
    }
}

/// search.rs:54-103
#[derive(Debug, Clone, serde::Serialize, serde::Deserialize)]
pub struct SearchResult {
    pub id: u64,
    pub score: f32,
    pub vector: Option<Vec<f32>>,
    pub metadata: Option<serde_json::Value>,
    pub text: Option<String>,
}

impl SearchResult {
    pub fn new(id: u64, score: f32) -> Self {
        Self { id, score, vector: None, metadata: None, text: None }
    }
    pub fn with_vector(mut self, vector: Vec<f32>) -> Self {
        self.vector = Some(vector);
        self
    }
    pub fn with_metadata(mut self, metadata: serde_json::Value) -> Self {
        self.metadata = Some(metadata);
        self
    }
    pub fn with_text(mut self, text: impl Into<String>) -> Self {
        self.text = Some(text.into());
        self
    }
    /// 1 / (1 + distance)
    pub fn to_similarity(&self) -> f32 {
        1.0 / (1.0 + self.score)
    }
}

/// Raw `(id, distance)` rows of one graph -> `SearchResult`s under `config` (vectors attached on request, then the
/// `min_similarity` filter: search.rs:155-176).
fn decorate(graph: &HnswGraph, config: &SearchConfig, raw: impl Iterator<Item = (u64, f32)>) -> Vec<SearchResult> {
    raw.map(|(id, distance)| {
        let hit = SearchResult::new(id, distance);
        // the node is only fetched from the device when its vector was asked for
        match config.include_vectors.then(|| graph.get_node(id)).flatten() {
            Some(node) => hit.with_vector(node.vector),
            None => hit,
        }
    })
    .filter(|hit| config.min_similarity.map_or(true, |floor| hit.to_similarity() >= floor))
    .collect()
}

/// search.rs:106-182
pub struct Searcher<'a> {
    graph: &'a HnswGraph,
    config: SearchConfig,
}

impl<'a> Searcher<'a> {
    pub fn new(graph: &'a HnswGraph) -> Self {
        Self { graph, config: SearchConfig::default() }
    }
    pub fn with_config(graph: &'a HnswGraph, config: SearchConfig) -> Self {
        Self { graph, config }
    }
    pub fn top_k(mut self, k: usize) -> Self {
        self.config.top_k = k;
        self
    }
    pub fn ef(mut self, ef: usize) -> Self {
        self.config.ef = ef;
        self
    }
    pub fn include_vectors(mut self) -> Self {
        self.config.include_vectors = true;
        self
    }
    pub fn min_similarity(mut self, threshold: f32) -> Self {
        self.config.min_similarity = Some(threshold);
        self
    }

    /// search.rs:150-176
    pub fn search(&self, query: &[f32]) -> CoreResult<Vec<SearchResult>> {
        let raw = self.graph.search(query, self.config.top_k, self.config.ef)?;
        Ok(decorate(self.graph, &self.config, raw.into_iter()))
    }

    /// search.rs:179-181 maps `search` over the queries one by one; here the whole batch is ONE library call
    /// (one kernel launch over all queries), with the same per-query results.
    pub fn search_batch(&self, queries: &[Vec<f32>]) -> CoreResult<Vec<Vec<SearchResult>>> {
        let nq = queries.len();
        if nq == 0 || self.graph.is_empty() {
            return Ok(vec![Vec::new(); nq]);
        }
        let dim = self.graph.dimension().unwrap_or(queries[0].len());
        let mut flat = Vec::with_capacity(nq * dim);
        for q in queries {
            if q.len() != dim {
                return Err(CoreError::DimensionMismatch { expected: dim, actual: q.len() });
            }
            flat.extend_from_slice(q);
        }
        let k = self.config.top_k;
        let (ids, dist, count) = self.graph.search_batch(&flat, nq, k, self.config.ef)?;
        Ok((0..nq)
            .map(|q| {
                let rows = (0..count[q] as usize).map(|i| (ids[q * k + i], dist[q * k + i]));
                decorate(self.graph, &self.config, rows)
            })
            .collect())
    }
}

/// search.rs:185-254: every island is searched and the lists are merged by score, island order breaking ties (the
/// reference's stable sort).  `min_similarity` is not applied here, as in the reference.
pub struct MultiIndexSearcher {
    graphs: Vec<(String, HnswGraph)>,
    config: SearchConfig,
}

impl Default for MultiIndexSearcher {
    fn default() -> Self {
        Self::new()
    }
}

impl MultiIndexSearcher {
    pub fn new() -> Self {
        Self { graphs: Vec::new(), config: SearchConfig::default() }
    }
    pub fn add_index(&mut self, name: impl Into<String>, graph: HnswGraph) {
        self.graphs.push((name.into(), graph));
    }
    pub fn with_config(mut self, config: SearchConfig) -> Self {
        self.config = config;
        self
    }
    /// search.rs:211-237
    pub fn search(&self, query: &[f32]) -> CoreResult<Vec<(String, SearchResult)>> {
        let unfiltered = SearchConfig { min_similarity: None, ..self.config.clone() };
        let mut merged: Vec<(String, SearchResult)> = Vec::new();
        for (name, graph) in &self.graphs {
            let raw = graph.search(query, self.config.top_k, self.config.ef)?;
            merged.extend(decorate(graph, &unfiltered, raw.into_iter()).into_iter().map(|hit| (name.clone(), hit)));
        }
        merged.sort_by(|a, b| a.1.score.partial_cmp(&b.1.score).unwrap_or(std::cmp::Ordering::Equal));
        merged.truncate(self.config.top_k);
        Ok(merged)
    }
    pub fn num_indexes(&self) -> usize {
        self.graphs.len()
    }
    pub fn total_vectors(&self) -> usize {
        self.graphs.iter().map(|(_, g)| g.len()).sum()
    }
}

/// mod.rs:88-99
pub mod prelude {
    pub use super::{
        CoreError, CoreResult, CsrGraph, Distance, DistanceMetric, EmbeddingProvider, HnswConfig, HnswGraph, InMemoryEmbeddingProvider,
        LeannConfig, LeannIndex, ProductQuantizer, PruningStrategy, SearchConfig, SearchResult, Searcher,
    };
}

// ---------------------------------------------------------------------------------------------
// pq.rs:13-359 — PQConfig, ProductQuantizer
// ---------------------------------------------------------------------------------------------
#[derive(Debug, Clone)]
pub struct PQConfig {
    pub num_subquantizers: usize,
    pub num_centroids: usize,
    pub training_iterations: usize,
    pub seed: Option<u64>,
}

impl Default for PQConfig {
    fn default() -> Self {
        Self { num_subquantizers: 8, num_centroids: 256, training_iterations: 25, seed: None }
    }
}

impl PQConfig {
    fn raw(&self) -> sys::IslPqConfig {
        sys::IslPqConfig {
            num_subquantizers: self.num_subquantizers as u64,
            num_centroids: self.num_centroids as u64,
            training_iterations: self.training_iterations as u64,
            seed: self.seed.unwrap_or(0),
            has_seed: self.seed.is_some() as i32,
        }
    }
    /// pq.rs:37-55
    pub fn validate(&self, dimension: usize) -> CoreResult<()> {
        check(unsafe { sys::isl_pq_config_validate(&self.raw(), dimension as u64) })
    }
    /// pq.rs:58-64
    pub fn bytes_per_vector(&self) -> usize {
        unsafe { sys::isl_pq_config_bytes_per_vector(&self.raw()) as usize }
    }
}

/// pq.rs:66-112: the centroids of one subquantizer.
#[derive(Debug, Clone)]
pub struct PQCodebook {
    pub centroids: Vec<Vec<f32>>,
    pub subvector_dim: usize,
}

impl PQCodebook {
    pub fn new(subvector_dim: usize) -> Self {
        Self { centroids: Vec::new(), subvector_dim }
    }
    /// pq.rs:86-106, on the device: a one-subquantizer quantizer holding these centroids encodes the subvector (strict
    /// `<` scan from `f32::MAX`: the first of equal centroids wins).
    pub fn find_nearest(&self, subvector: &[f32], metric: &DistanceMetric) -> CoreResult<usize> {
        if subvector.len() != self.subvector_dim {
            return Err(CoreError::DimensionMismatch { expected: self.subvector_dim, actual: subvector.len() });
        }
        let config = PQConfig { num_subquantizers: 1, num_centroids: self.centroids.len(), training_iterations: 1, seed: None };
        let mut pq = ProductQuantizer::new(self.subvector_dim, config)?.with_metric(*metric);
        pq.set_codebooks(std::slice::from_ref(self))?;
        Ok(pq.encode(subvector)?[0] as usize)
    }
    pub fn get_centroid(&self, idx: usize) -> Option<&[f32]> {
        self.centroids.get(idx).map(Vec::as_slice)
    }
}

pub struct ProductQuantizer {
    handle: *mut sys::IslPq,
    dimension: usize,
}

unsafe impl Send for ProductQuantizer {}
unsafe impl Sync for ProductQuantizer {}

impl Drop for ProductQuantizer {
    fn drop(&mut self) {
        unsafe { sys::isl_pq_free(self.handle) };
    }
}

impl ProductQuantizer {
    /// pq.rs:133-149
    pub fn new(dimension: usize, config: PQConfig) -> CoreResult<Self> {
        let mut h: *mut sys::IslPq = ptr::null_mut();
        check(unsafe { sys::isl_pq_new(dimension as u32, &config.raw(), &mut h) })?;
        Ok(Self { handle: h, dimension })
    }
    /// pq.rs:152-155
    pub fn with_metric(self, metric: DistanceMetric) -> Self {
        let _ = unsafe { sys::isl_pq_set_metric(self.handle, metric.code()) };
        self
    }
    pub fn is_trained(&self) -> bool {
        unsafe { sys::isl_pq_is_trained(self.handle) != 0 }
    }
    pub fn num_subquantizers(&self) -> usize {
        unsafe { sys::isl_pq_num_subquantizers(self.handle) as usize }
    }
    pub fn compression_ratio(&self) -> f32 {
        unsafe { sys::isl_pq_compression_ratio(self.handle) }
    }
    /// pq.rs:175-218
    pub fn train(&mut self, vectors: &[Vec<f32>]) -> CoreResult<()> {
        if vectors.is_empty() {
            return Err(CoreError::EmptyCollection);
        }
        let mut flat = Vec::with_capacity(vectors.len() * self.dimension);
        for v in vectors {
            if v.len() != self.dimension {
                return Err(CoreError::DimensionMismatch { expected: self.dimension, actual: v.len() });
            }
            flat.extend_from_slice(v);
        }
        check(unsafe { sys::isl_pq_train(self.handle, flat.as_ptr(), vectors.len() as u64, self.dimension as u32) })
    }
    /// Install codebooks trained elsewhere (one per subquantizer, equal sizes); marks the quantizer trained.
    pub fn set_codebooks(&mut self, codebooks: &[PQCodebook]) -> CoreResult<()> {
        let ksub = codebooks.first().map_or(0, |c| c.centroids.len());
        let mut flat = Vec::new();
        for cb in codebooks {
            if cb.centroids.len() != ksub {
                return Err(CoreError::PQError("codebooks must hold the same number of centroids".into()));
            }
            for centroid in &cb.centroids {
                if centroid.len() != cb.subvector_dim {
                    return Err(CoreError::DimensionMismatch { expected: cb.subvector_dim, actual: centroid.len() });
                }
                flat.extend_from_slice(centroid);
            }
        }
        if codebooks.len() != self.num_subquantizers() || flat.len() != ksub * self.dimension {
            return Err(CoreError::PQError("one codebook per subquantizer, subvector_dim = dimension / num_subquantizers".into()));
        }
        check(unsafe { sys::isl_pq_set_codebooks(self.handle, flat.as_ptr(), ksub as u64) })
    }
    /// `codebooks` (pq.rs:120-121), copied out of the library.
    pub fn codebooks(&self) -> CoreResult<Vec<PQCodebook>> {
        let m = self.num_subquantizers();
        let mut ksub = 0u64;
        check(unsafe { sys::isl_pq_get_codebooks(self.handle, ptr::null_mut(), &mut ksub) })?;
        let dsub = if m == 0 { 0 } else { self.dimension / m };
        let mut flat = vec![0f32; m * ksub as usize * dsub];
        check(unsafe { sys::isl_pq_get_codebooks(self.handle, flat.as_mut_ptr(), &mut ksub) })?;
        Ok(flat
            .chunks((ksub as usize * dsub).max(1))
            .take(m)
            .map(|sub| PQCodebook { centroids: sub.chunks(dsub.max(1)).map(<[f32]>::to_vec).collect(), subvector_dim: dsub })
            .collect())
    }
    /// pq.rs:221-244
    pub fn encode(&self, vector: &[f32]) -> CoreResult<Vec<u16>> {
        let mut codes = vec![0u16; self.num_subquantizers()];
        check(unsafe { sys::isl_pq_encode(self.handle, vector.as_ptr(), 1, vector.len() as u32, codes.as_mut_ptr()) })?;
        Ok(codes)
    }
    /// Batched encode: `[n][dimension]` row-major -> `[n][m]` codes.
    pub fn encode_batch(&self, vectors: &[f32], n: usize) -> CoreResult<Vec<u16>> {
        let mut codes = vec![0u16; n * self.num_subquantizers()];
        check(unsafe { sys::isl_pq_encode(self.handle, vectors.as_ptr(), n as u64, self.dimension as u32, codes.as_mut_ptr()) })?;
        Ok(codes)
    }
    /// pq.rs:247-271
    pub fn decode(&self, codes: &[u16]) -> CoreResult<Vec<f32>> {
        let mut out = vec![0f32; self.dimension];
        check(unsafe { sys::isl_pq_decode(self.handle, codes.as_ptr(), 1, codes.len() as u64, out.as_mut_ptr()) })?;
        Ok(out)
    }
    /// pq.rs:275-304
    pub fn asymmetric_distance(&self, query: &[f32], codes: &[u16]) -> CoreResult<f32> {
        if codes.len() != self.num_subquantizers() {
            return Err(CoreError::PQError("Codes length mismatch".into()));
        }
        let mut out = 0f32;
        check(unsafe { sys::isl_pq_asymmetric_distance(self.handle, query.as_ptr(), query.len() as u32, codes.as_ptr(), 1, &mut out) })?;
        Ok(out)
    }
    /// pq.rs:307-338: tables[m][ksub]
    pub fn build_distance_tables(&self, query: &[f32]) -> CoreResult<Vec<Vec<f32>>> {
        let m = self.num_subquantizers();
        let mut ksub = 0u64;
        check(unsafe { sys::isl_pq_get_codebooks(self.handle, ptr::null_mut(), &mut ksub) })?;
        let mut flat = vec![0f32; m * ksub as usize];
        check(unsafe { sys::isl_pq_build_tables(self.handle, query.as_ptr(), query.len() as u32, flat.as_mut_ptr()) })?;
        Ok(flat.chunks(ksub as usize).map(|c| c.to_vec()).collect())
    }
    /// pq.rs:341-348 (host arithmetic: m table lookups, left fold, sqrt — too small to ship to the device)
    pub fn table_distance(&self, tables: &[Vec<f32>], codes: &[u16]) -> f32 {
        let mut s = 0f32;
        for (t, &c) in tables.iter().zip(codes) {
            s += t[c as usize];
        }
        s.sqrt()
    }
    pub fn to_bytes(&self) -> CoreResult<Vec<u8>> {
        let mut len = 0u64;
        check(unsafe { sys::isl_pq_to_bytes(self.handle, ptr::null_mut(), 0, &mut len) })?;
        let mut out = vec![0u8; len as usize];
        check(unsafe { sys::isl_pq_to_bytes(self.handle, out.as_mut_ptr(), len, &mut len) })?;
        Ok(out)
    }
    pub fn from_bytes(bytes: &[u8]) -> CoreResult<Self> {
        let mut h: *mut sys::IslPq = ptr::null_mut();
        check(unsafe { sys::isl_pq_from_bytes(bytes.as_ptr(), bytes.len() as u64, &mut h) })?;
        let dimension = unsafe { sys::isl_pq_dimension(h) } as usize;
        Ok(Self { handle: h, dimension })
    }
}
