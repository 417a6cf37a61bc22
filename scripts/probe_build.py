"""Measurement scaffold: GPU build (isl_index_build_dev) at scale + recall-vs-ef sweep."""
import argparse, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from islands_b200 import LeannConfig, LeannIndex  # noqa: E402


def make_data(n, nq, d, clusters, dev, sigma=0.15, latent=0):
    g = torch.Generator(device=dev); g.manual_seed(42)
    if latent:
        A = torch.randn((latent, d), generator=g, device=dev) / latent ** 0.5
        if clusters:
            c = torch.randn((clusters, latent), generator=g, device=dev)
            zx = c[torch.randint(0, clusters, (n,), generator=g, device=dev)] + 0.5 * torch.randn((n, latent), generator=g, device=dev)
            zq = c[torch.randint(0, clusters, (nq,), generator=g, device=dev)] + 0.5 * torch.randn((nq, latent), generator=g, device=dev)
        else:
            zx = torch.randn((n, latent), generator=g, device=dev); zq = torch.randn((nq, latent), generator=g, device=dev)
        x = zx @ A + sigma * torch.randn((n, d), generator=g, device=dev)
        q = zq @ A + sigma * torch.randn((nq, d), generator=g, device=dev)
    elif clusters:
        centers = torch.rand((clusters, d), generator=g, device=dev) * 2 - 1
        x = centers[torch.randint(0, clusters, (n,), generator=g, device=dev)] + sigma * torch.randn((n, d), generator=g, device=dev)
        q = centers[torch.randint(0, clusters, (nq,), generator=g, device=dev)] + sigma * torch.randn((nq, d), generator=g, device=dev)
    else:
        x = torch.rand((n, d), generator=g, device=dev) * 2 - 1
        q = torch.rand((nq, d), generator=g, device=dev) * 2 - 1
    return x.contiguous(), q.contiguous()


def ground_truth(x, q, k, chunk=1024):
    xn = torch.nn.functional.normalize(x, dim=1); qn = torch.nn.functional.normalize(q, dim=1)
    return torch.cat([(qn[s:s + chunk] @ xn.T).topk(k, dim=1).indices for s in range(0, q.shape[0], chunk)]).cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000000); ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--nq", type=int, default=2000); ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--efs", type=str, default="64,256,1024"); ap.add_argument("--clusters", type=int, default=0)
    ap.add_argument("--sigma", type=float, default=0.15); ap.add_argument("--latent", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    x, q = make_data(a.n, a.nq, a.d, a.clusters, dev, a.sigma, a.latent)
    gt = ground_truth(x, q, 10)
    torch.cuda.synchronize()
    cfg = LeannConfig()
    idx = LeannIndex(cfg)
    t0 = time.time()
    idx.build_dev(x.data_ptr(), a.n, a.d, seed=7, batch=a.batch)
    bt = time.time() - t0
    g = idx.graph
    deg = g.degree_counts.astype(np.int64)
    print(json.dumps(dict(event="build", n=a.n, batch=a.batch, clusters=a.clusters, seconds=round(bt, 2), edges=int(g.neighbors.size),
                          mean_deg=float(deg.mean()), min_deg=int(deg.min()), entry=g.entry_point, max_level=g.max_level)), flush=True)
    ids = torch.empty((a.nq, 10), dtype=torch.int64, device=dev); dist = torch.empty((a.nq, 10), dtype=torch.float32, device=dev)
    cnt = torch.empty((a.nq,), dtype=torch.int32, device=dev); stats = torch.zeros((a.nq, 5), dtype=torch.int64, device=dev)
    for ef in [int(e) for e in a.efs.split(",")]:
        for _ in range(2):
            t0 = time.time()
            idx.search_batch_dev(q.data_ptr(), a.nq, a.d, 10, ef, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), stats.data_ptr())
            wall = time.time() - t0
        ms, _ = idx.last_search_timing()
        got = ids.cpu().numpy(); st = stats.cpu().numpy()
        recall = np.mean([len(set(got[i].tolist()) & set(gt[i].tolist())) / 10 for i in range(a.nq)])
        nh, ne, nd = st[:, 0].sum(), st[:, 1].sum(), st[:, 2].sum()
        b = nd * 4 * a.d + ne * 4 + nh * 16 + a.nq * (4 * a.d + 120)
        print(json.dumps(dict(event="search", ef=ef, kernel_ms=round(ms, 2), wall_ms=round(wall * 1e3, 2), qps=round(a.nq / ms * 1e3, 1), recall10=round(float(recall), 4),
                              n_dist=float(nd) / a.nq, n_hop=float(nh) / a.nq, gbps=round(b / ms / 1e6, 1), frac=round(b / ms / 1e6 / 6543.1, 4))), flush=True)


if __name__ == "__main__":
    main()
