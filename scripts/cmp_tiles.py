"""Bitwise comparison of the same GEMM under different tile shapes (tuning aid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chinesener_b200 import ops

g = torch.Generator().manual_seed(0)
for (M, N, K) in [(400, 768, 768), (400, 2304, 768), (400, 3072, 768), (400, 768, 3072)]:
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).cuda()
    ref = a.double() @ w.double().t()
    outs = {t: ops.gemm_bf16(a, w, None, epilogue=ops.EPI_F32, tile_n=t) for t in (128, 192, 256)}
    for t, o in outs.items():
        d = (o.double() - ref).abs().max().item()
        same = {u: int((o != outs[u]).sum()) for u in outs if u != t}
        print(M, N, K, "tile", t, "max|err| vs fp64", f"{d:.3e}", "elements differing from other tiles", same)
