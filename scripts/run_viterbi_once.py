"""Three calls of the large-batch Viterbi kernel at the roofline shape (for an ncu capture: -k regex:crf_viterbi -s 1 -c 1)."""
import sys

import torch

sys.path.insert(0, ".")
from chinesener_b200 import ops  # noqa: E402

B, L, K = 262144, 128, 10
g = torch.Generator(device="cuda").manual_seed(1234)
x = torch.randn(B, L, K, device="cuda", generator=g)
tr = torch.randn(K, K, device="cuda", generator=g) * 0.5
lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
for _ in range(3):
    ops.crf_viterbi(x, lens, tr)
torch.cuda.synchronize()
