import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from islands_b200 import gemm_bf16_dev
dev = torch.device("cuda:0")
m, n, k = 131072, 2304, 768
a = torch.randn((m, k), device=dev).to(torch.bfloat16); w = torch.randn((n, k), device=dev).to(torch.bfloat16)
o = torch.empty((m, n), device=dev, dtype=torch.bfloat16)
torch.cuda.synchronize()
for _ in range(3):
    gemm_bf16_dev(a.data_ptr(), w.data_ptr(), m, n, k, d_out_bf16=o.data_ptr())
for _ in range(3):
    o2 = torch.matmul(a, w.T)
torch.cuda.synchronize()
print("ok")
