"""IslandRegistry (search half of IndexerService, service.rs:608-674, 737-818) and the safetensors
loader of the recompute encoder."""
import json
import struct

import numpy as np
import pytest

from conftest import uniform

pytestmark = pytest.mark.gpu


def test_registry_search_matches_reference_recipe(gpu_lib, orc):
    from islands_b200 import HnswConfig, IslandRegistry

    cfg = HnswConfig(m=8, m0=16, ef_construction=40, ml=0.9)
    d = 16
    rng = np.random.RandomState(3)
    reg = IslandRegistry(insert_batch=8)
    oracles = {}
    for name, n in (("repo/a", 120), ("repo/b", 80)):
        v = uniform(rng, n, d)
        lv = orc.draw_levels(len(name), n, cfg.ml, cfg.max_layers)
        files = [(f"{name}/file{i}.rs", f"fn f{i}() {{}} " * 30) for i in range(n)]
        reg.add_island(name, files, embeddings=v, config=cfg, levels=lv)
        og = orc.Hnsw(cfg._s, d)
        og.insert_batch(v, lv, batch=8, threads=8)
        oracles[name] = (og, files)
    assert reg.list_indexes() == ["repo/a", "repo/b"]
    q = uniform(rng, 6, d)

    def reference(qi, names, top_k):  # service.rs:775-801 on the oracle graphs
        rows = []
        for nm in names:
            og, files = oracles[nm]
            ids, dist, cnt = og.search(q[qi:qi + 1], top_k, max(top_k, 100))
            for j in range(cnt[0]):
                path, content = files[int(ids[0, j])]
                rows.append((np.float32(1.0) - dist[0, j], nm, path, content[:200]))
        rows.sort(key=lambda r: -r[0])
        return [{"score": float(s), "index": nm, "path": p, "snippet": sn} for s, nm, p, sn in rows[:top_k]]

    for qi in range(6):
        assert reg.search(q[qi], top_k=5) == reference(qi, ["repo/a", "repo/b"], 5)
        assert reg.search(q[qi], index_names=["repo/b"], top_k=3) == reference(qi, ["repo/b"], 3)
        assert reg.search(q[qi], index_names=["nope"], top_k=3) == []
    assert len(reg.search(q[0], top_k=5)[0]["snippet"]) == 200
    with pytest.raises(RuntimeError):
        IslandRegistry().add_island("x", [("p", "c")])  # no embedder (service.rs:617-621)


def test_encoder_loads_safetensors(gpu_lib, tmp_path):
    from islands_b200 import Encoder, EncoderConfig, SerializationError

    cfg = EncoderConfig(vocab_size=300, hidden_size=64, num_layers=1, num_heads=1, intermediate_size=128, max_position=16)
    src = Encoder(cfg).init_random(seed=9, stddev=0.05)
    sd = src.state_dict()
    header, blobs, off = {}, [], 0
    for i, (name, arr) in enumerate(sd.items()):
        if i % 2:  # alternate F32 and BF16 tensors
            b = (arr.astype(np.float32).view(np.uint32) >> 16).astype("<u2").tobytes()
            dt = "BF16"
        else:
            b = arr.astype("<f4").tobytes()
            dt = "F32"
        header["bert." + name] = {"dtype": dt, "shape": list(arr.shape), "data_offsets": [off, off + len(b)]}
        blobs.append(b)
        off += len(b)
    hj = json.dumps(header).encode()
    path = tmp_path / "model.safetensors"
    path.write_bytes(struct.pack("<Q", len(hj)) + hj + b"".join(blobs))
    dst = Encoder(cfg)
    loaded = dst.load_safetensors(str(path), prefix="bert.")
    assert len(loaded) == len(sd)
    for i, (name, arr) in enumerate(sd.items()):
        want = (arr.view(np.uint32) >> 16 << 16).view(np.float32) if i % 2 else arr
        assert np.array_equal(dst.get_parameter(name), want), name
    tok = np.random.RandomState(0).randint(1, 300, size=(5, 12)).astype(np.int32)
    ln = np.array([12, 9, 5, 3, 1], np.int32)
    out = dst.embed(tok, ln)
    assert out.shape == (5, 64) and np.isfinite(out).all()
    with pytest.raises(SerializationError):
        Encoder(cfg).load_safetensors(str(path), prefix="missing.")
