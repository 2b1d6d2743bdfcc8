"""TRAIN-step profile helper: host enqueue time vs device time of `bert_bilstm_crf` train steps.
Run plainly for the timing JSON, or under `ncu --metrics gpu__time_duration.sum` for the launch list."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chinesener_b200 import _lib, engine, synthetic  # noqa: E402


def main(steps=5, B=64, L=128):
    params = dict(synthetic.data_params(L, 10), pretrain_dir="", num_train_steps=10000, warmup_ratio=0.1)
    est = engine.Estimator("bert_bilstm_crf", params)
    batches = [est.to_device(synthetic.msra_batch(B, L, seed=77 + i)) for i in range(2)]
    for i in range(3):
        est.train_step(batches[i % 2])
    torch.cuda.synchronize()
    # device time, GPU not starved: park it behind a spin kernel while the host enqueues
    torch.cuda._sleep(60_000_000)
    l0 = _lib.LAUNCHES
    h0 = time.perf_counter()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        est.train_step(batches[i % 2])
    e.record()
    host_ms = (time.perf_counter() - h0) * 1e3 / steps
    torch.cuda.synchronize()
    # plain back-to-back steps (what bench.py's `train` object measures)
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record()
    for i in range(steps):
        est.train_step(batches[i % 2])
    e2.record()
    torch.cuda.synchronize()
    if os.environ.get("NER_PROF_HOST"):
        import cProfile
        import pstats
        torch.cuda._sleep(60_000_000)
        pr = cProfile.Profile()
        pr.enable()
        for i in range(3):
            est.train_step(batches[i % 2])
        pr.disable()
        torch.cuda.synchronize()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
    print(json.dumps({"host_enqueue_ms_per_step": host_ms, "library_launches_per_step": (_lib.LAUNCHES - l0) / (2 * steps),
                      "ms_per_step_back_to_back": s2.elapsed_time(e2) / steps,
                      "note": "first figure = host time to enqueue one step while the GPU is parked"}))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 5)
