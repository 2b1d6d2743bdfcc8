"""Run one GEMM shape a few times (for ncu captures).  usage: prof_gemm.py TILE_N EPI [M N K]"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chinesener_b200 import ops

tile, epi = int(sys.argv[1]), int(sys.argv[2])
M, N, K = (int(x) for x in sys.argv[3:6]) if len(sys.argv) >= 6 else (8192, 2304, 768)
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
res = torch.randn(M, N, device="cuda") if epi == ops.EPI_RES_F32 else None
o = torch.empty(M, N, device="cuda", dtype=torch.float32 if epi in (ops.EPI_F32, ops.EPI_RES_F32) else torch.bfloat16)
for _ in range(6):
    ops.gemm_bf16(a, w, bias, residual=res, epilogue=epi, tile_n=tile, out=o)
torch.cuda.synchronize()
