"""Island registry: the search half of the reference's IndexerService (src/indexer/service.rs) —
`graphs: HashMap<String, StoredIndex{graph: HnswGraph, files}>`, `build_index_with_embeddings`
(:608-674: one HnswGraph per repository, embeddings inserted in order) and `search_with_embeddings`
(:737-818: optional `index_names` filter, ef = max(top_k, 100), score = 1 - distance, results of all
islands sorted by score descending and truncated).  Git, storage, watching and the MCP / CLI layers
are out of scope.  Graph construction and search run on the GPU (isl_hnsw_*).
"""
import numpy as np

from .core import HnswConfig, HnswGraph, _f32


class StoredIndex:
    """StoredIndex (service.rs): the island's graph and its (path, content) files, id = position."""

    def __init__(self, graph, files):
        self.graph = graph
        self.files = list(files)


class IslandRegistry:
    def __init__(self, embed=None, insert_batch=1024):
        """embed: callable texts -> [n, d] f32 (the EmbeddingProvider: e.g. tokenizer + Encoder.embed)."""
        self.graphs = {}  # insertion-ordered, like the iteration the tests rely on
        self.embed = embed
        self.insert_batch = insert_batch

    # ---- build_index_with_embeddings (service.rs:608-674) ------------------------------------------
    def add_island(self, name, files, embeddings=None, config=None, levels=None, seed=0):
        files = list(files)
        if embeddings is None:
            if self.embed is None:
                raise RuntimeError("Embedder not initialized. Call init_embedder() first.")  # service.rs:617-621
            embeddings = self.embed([content for _, content in files])
        emb = _f32(embeddings)
        graph = HnswGraph(config or HnswConfig())  # HnswConfig::default() (service.rs:624)
        if emb.shape[0]:
            graph.insert_batch(emb, levels=levels, seed=seed, batch=self.insert_batch)
        self.graphs[str(name)] = StoredIndex(graph, files)
        return graph

    def remove_island(self, name):
        self.graphs.pop(str(name), None)

    def list_indexes(self):
        return list(self.graphs.keys())

    # ---- search_with_embeddings (service.rs:737-818) -----------------------------------------------
    def search(self, query_embedding, index_names=None, top_k=10):
        """-> list of {"score", "index", "path", "snippet"} dicts, best first."""
        return self.search_batch(_f32(query_embedding).reshape(1, -1), index_names, top_k)[0]

    def search_batch(self, query_embeddings, index_names=None, top_k=10):
        q = _f32(query_embeddings)
        names = list(index_names) if index_names is not None else list(self.graphs.keys())  # service.rs:768-771
        ef = max(top_k, 100)  # service.rs:780
        per_query = [[] for _ in range(q.shape[0])]
        for name in names:
            stored = self.graphs.get(name)
            if stored is None or len(stored.graph) == 0:
                continue
            ids, dist, cnt = stored.graph.search_batch(q, top_k, ef)  # one GPU call per island for the batch
            for i in range(q.shape[0]):
                for j in range(int(cnt[i])):
                    nid = int(ids[i, j])
                    if nid < len(stored.files):  # service.rs:788
                        path, content = stored.files[nid]
                        score = np.float32(1.0) - dist[i, j]  # service.rs:791
                        per_query[i].append((score, name, path, content[:200]))
        out = []
        for rows in per_query:
            rows.sort(key=lambda r: -r[0])  # stable, descending by score (service.rs:800)
            out.append([{"score": float(s), "index": nm, "path": p, "snippet": sn} for s, nm, p, sn in rows[:top_k]])
        return out
