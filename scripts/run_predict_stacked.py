"""A few stacked PREDICT calls of bert_bilstm_crf (G batches of 64 sentences per call) for `ncu` launch lists:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python scripts/run_predict_stacked.py 4 3"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chinesener_b200 import engine, ops, synthetic  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 4
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
est = engine.Estimator("bert_bilstm_crf", dict(synthetic.data_params(128, 10), pretrain_dir=""))
batches = [{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in synthetic.msra_batch(64, 128, seed=1234 + i).items()} for i in range(G)]
est.predict(batches[0])                       # creates variables / weight packs
dev = est.stack_to_device(batches)
ops.DEFAULT_TILE = ops.TILE_AUTO_THROUGHPUT
for _ in range(calls):
    est.predict_device(dev)
torch.cuda.synchronize()
print("tokens per call", int(dev['mask'].total_tokens))
