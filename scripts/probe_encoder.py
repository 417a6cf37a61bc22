"""Dev probe: encoder throughput (BERT-base shape, random init) and the GEMMs on their own."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from islands_b200 import Encoder, EncoderConfig, gemm_bf16_dev
dev = torch.device("cuda:0")
peak = 1404.3
for (m, n, k) in [(131072, 2304, 768), (131072, 768, 768), (131072, 3072, 768), (131072, 768, 3072), (8192, 3072, 768)]:
    a = torch.randn((m, k), device=dev).to(torch.bfloat16); w = torch.randn((n, k), device=dev).to(torch.bfloat16)
    o = torch.empty((m, n), device=dev, dtype=torch.bfloat16)
    torch.cuda.synchronize()
    for _ in range(3): gemm_bf16_dev(a.data_ptr(), w.data_ptr(), m, n, k, d_out_bf16=o.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): gemm_bf16_dev(a.data_ptr(), w.data_ptr(), m, n, k, d_out_bf16=o.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    for _ in range(3): torch.matmul(a, w.T)
    t0.record()
    for _ in range(10): torch.matmul(a, w.T)
    t1.record(); torch.cuda.synchronize()
    cms = t0.elapsed_time(t1) / 10
    print(json.dumps(dict(gemm=[m, n, k], ms=round(ms, 3), tflops=round(2 * m * n * k / ms / 1e9, 1), cublas_ms=round(cms, 3), cublas_tflops=round(2 * m * n * k / cms / 1e9, 1))), flush=True)
enc = Encoder(EncoderConfig()).init_random()
for B, S in [(2048, 64), (1024, 128), (256, 64)]:
    tok = torch.randint(1, 30000, (B, S), device=dev, dtype=torch.int32); ln = torch.full((B,), S, device=dev, dtype=torch.int32)
    out = torch.empty((B, 768), device=dev)
    torch.cuda.synchronize()
    for _ in range(3):
        enc.embed_dev(tok.data_ptr(), ln.data_ptr(), B, S, out.data_ptr())
    ms, fl = enc.last_timing()
    print(json.dumps(dict(B=B, S=S, ms=round(ms, 2), tflops=round(fl / ms / 1e9, 1), frac_sustained=round(fl / ms / 1e9 / peak, 3), seq_per_s=round(B / ms * 1e3))), flush=True)
