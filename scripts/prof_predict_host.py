"""Host-side profile of one PREDICT step of bert_bilstm_crf (GPU parked behind a spin kernel)."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chinesener_b200 import engine, synthetic  # noqa: E402

params = dict(synthetic.data_params(128, 10), pretrain_dir="")
est = engine.Estimator("bert_bilstm_crf", params)
batches = [{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in synthetic.msra_batch(64, 128, seed=5 + i).items()}
           for i in range(4)]
for b in batches:
    est.predict(b)
torch.cuda.synchronize()
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
torch.cuda._sleep(120_000_000)
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
n = 0
for out in est.predict_iter((batches[i % 4] for i in range(8)), streams=4):
    n += 1
pr.disable()
print("host ms per step (incl. profiler overhead):", (time.perf_counter() - t0) * 1e3 / 8)
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
