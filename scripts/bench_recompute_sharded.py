"""BASELINE configs[4] at scale: on-demand recompute over a node-range-sharded, GRAPH-ONLY index (torchrun, one rank
per GPU).  Every rank owns `--shard-nodes` nodes: a token row per node, the graph built over the encoder's outputs,
PQ codes — and then DROPS the f32 embeddings.  A query batch runs isl_index_search_sharded_adc(recompute): ADC
traversal on every shard, the shard's distinct survivors through the random-init BERT-base encoder (110M parameters,
bf16 tcgen05), exact rerank against the recomputed rows, ONE exchange of the per-shard records, merge.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \
      scripts/bench_recompute_sharded.py --shard-nodes 12500000 [--out FILE]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shard-nodes", type=int, default=200_000)
    ap.add_argument("--nq", type=int, default=256)
    ap.add_argument("--seq", type=int, default=64)
    ap.add_argument("--efs", default="128,256")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    import bench
    from islands_b200 import Encoder, EncoderConfig, LeannConfig, LeannIndex, PQConfig, ProductQuantizer
    from islands_b200.shard import make_shard_comm

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    n, nq, S, k = a.shard_nodes, a.nq, a.seq, 10
    t_all = time.perf_counter()

    # clustered token table (near-duplicate chunks: what gives a code corpus its neighbourhoods); queries shared by all ranks
    def draw(gen, base, count):
        t = base[torch.randint(0, base.shape[0], (count,), generator=gen, device=dev)].clone()
        flip = torch.rand((count, S), generator=gen, device=dev) < 0.15
        t[flip] = torch.randint(1, 30000, (int(flip.sum()),), generator=gen, device=dev, dtype=torch.int32)
        return t.contiguous(), torch.full((count,), S, device=dev, dtype=torch.int32)

    g0 = torch.Generator(device=dev)
    g0.manual_seed(45)
    clusters = max(64, n * world // 50)
    base = torch.randint(1, 30000, (clusters, S), generator=g0, device=dev, dtype=torch.int32)  # the same prototypes everywhere
    qtok, qln = draw(g0, base, nq)
    g1 = torch.Generator(device=dev)
    g1.manual_seed(1000 + rank)
    tok, ln = draw(g1, base, n)
    del base

    enc = Encoder(EncoderConfig()).init_random(seed=46, stddev=0.02)  # same weights on every rank
    emb = torch.empty((n, 768), device=dev)
    qemb = torch.empty((nq, 768), device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    enc.embed_dev(tok.data_ptr(), ln.data_ptr(), n, S, emb.data_ptr())
    t_embed = time.perf_counter() - t0
    enc.embed_dev(qtok.data_ptr(), qln.data_ptr(), nq, S, qemb.data_ptr())

    t0 = time.perf_counter()
    index = LeannIndex(LeannConfig())
    index.build_dev(emb.data_ptr(), n, 768, seed=7, batch=4096)
    t_build = time.perf_counter() - t0

    # global ground truth over all shards (brute force on the embeddings, before they are dropped)
    sc, li = bench.ground_truth_scores(torch, emb, qemb, k)
    g_sc = torch.empty((world * nq, k), dtype=sc.dtype, device=dev)
    g_id = torch.empty((world * nq, k), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(g_sc, sc.contiguous())
    dist.all_gather_into_tensor(g_id, (li + rank * n).contiguous())
    g_sc = g_sc.view(world, nq, k).permute(1, 0, 2).reshape(nq, -1)
    g_id = g_id.view(world, nq, k).permute(1, 0, 2).reshape(nq, -1)
    gt = g_id.gather(1, g_sc.topk(k, dim=1).indices).cpu().numpy()

    t0 = time.perf_counter()
    pq = ProductQuantizer(768, PQConfig(32, 256, 8, 1))
    pq.train(emb[:20000].cpu().numpy())
    codes = np.empty((n, 32), np.uint16)
    step = 1 << 20
    for s in range(0, n, step):  # the codes are made from device rows a block at a time: no host copy of the shard
        codes[s:s + step] = pq.encode(emb[s:s + step].cpu().numpy())
    index.attach_pq(pq, codes)
    t_pq = time.perf_counter() - t0
    del codes
    qh = qemb.cpu().numpy()
    comm = make_shard_comm(device=dev)

    def recall(ids):
        return float(np.mean([len(set(ids[i].tolist()) & set(gt[i].tolist())) / k for i in range(nq)]))

    efs = [int(v) for v in a.efs.split(",")]
    stored = {}
    for ef in efs:  # with the stored vectors still resident: the sharded "PQ ADC traversal + exact rerank"
        index.search_sharded_adc(comm, rank * n, qh, k, ef)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        ids, dst, cnt = index.search_sharded_adc(comm, rank * n, qh, k, ef)
        dt = time.perf_counter() - t0
        stored[str(ef)] = dict(batch_ms=round(dt * 1e3, 2), qps=round(nq / dt, 1), recall_at_10=recall(ids), ids=ids, dst=dst)

    index.set_recompute(enc, tok.cpu().numpy(), ln.cpu().numpy())
    del emb
    index.drop_vectors()  # graph + codes + token rows are all that is left of the shard
    torch.cuda.empty_cache()
    free_b, total_b = torch.cuda.mem_get_info()
    rec = {}
    for ef in efs:
        index.search_sharded_adc(comm, rank * n, qh, k, ef, recompute=True)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        ids, dst, cnt = index.search_sharded_adc(comm, rank * n, qh, k, ef, recompute=True)
        dt = time.perf_counter() - t0
        info = index.last_recompute()
        tt = torch.tensor([dt, float(info["unique_nodes"]), info["encoder_ms"]], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        same = bool(np.array_equal(ids, stored[str(ef)]["ids"]) and np.array_equal(dst.view(np.uint32), stored[str(ef)]["dst"].view(np.uint32)))
        rec[str(ef)] = dict(batch_ms=round(float(tt[0]) * 1e3, 1), qps=round(nq / float(tt[0]), 1), recall_at_10=recall(ids),
                            sequences_encoded_per_rank_max=int(tt[1]), encoder_ms_max=round(float(tt[2]), 1),
                            stage_ms_rank0=[round(v, 3) for v in comm.last_timing()],
                            identical_to_stored_vector_search=same)
    for v in stored.values():
        v.pop("ids"), v.pop("dst")
    line = dict(what="BASELINE configs[4]: on-demand recompute, random-init 110M bf16 encoder over the ADC survivors, node-range-sharded graph-only index",
                world=world, shard_nodes=n, total_nodes=world * n, queries=nq, seq_len=S, pq="m=32 ksub=256",
                index_embed_s=round(t_embed, 1), index_embed_seq_per_s=round(n / t_embed), build_s=round(t_build, 1), pq_s=round(t_pq, 1),
                device_memory_used_after_drop_gb=round((total_b - free_b) / 1e9, 1),
                sharded_adc_rerank_stored_vectors=stored, sharded_adc_recompute=rec, total_s=round(time.perf_counter() - t_all, 1))
    if rank == 0:
        text = json.dumps(line)
        print(text)
        if a.out:
            with open(a.out, "w") as f:
                f.write(text + "\n")
    comm.free()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
