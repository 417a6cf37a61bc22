import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from islands_b200 import Encoder, EncoderConfig
dev = torch.device("cuda:0")
enc = Encoder(EncoderConfig()).init_random()
B, S = int(os.environ.get("B", 2048)), int(os.environ.get("S", 64))
tok = torch.randint(1, 30000, (B, S), device=dev, dtype=torch.int32); ln = torch.full((B,), S, device=dev, dtype=torch.int32)
out = torch.empty((B, 768), device=dev)
torch.cuda.synchronize()
for _ in range(int(os.environ.get("REP", 3))):
    enc.embed_dev(tok.data_ptr(), ln.data_ptr(), B, S, out.data_ptr())
ms, fl = enc.last_timing()
print(json.dumps(dict(B=B, S=S, ms=round(ms, 2), tflops=round(fl / ms / 1e9, 1), frac_sustained=round(fl / ms / 1e9 / 1404.3, 3))))
