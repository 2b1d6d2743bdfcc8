"""CPU, world_size 2 over gloo: the data-parallel host logic — one all-reduce over the flat gradient
buffer, rank-sharded synthetic batches, max-over-ranks timing reduction (bench.py's N>1 plumbing)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from chinesener_b200 import synthetic, variables
        from chinesener_b200.tools import train_utils
        store = variables.VariableStore("cpu", seed=7)
        store.get_variable("logits/kernel", (4, 3), variables.glorot_uniform)
        store.get_variable("crf_layer/transitions", (3, 3), variables.xavier)
        store.get_variable("bert/x/LayerNorm/gamma", (5,), variables.ones)
        fs = train_utils.FlatState(store, lambda n: 0 if "crf" in n else 1)
        # identical initial weights on every rank (same seed), gradients differ per rank
        for n in fs.names:
            store.grads[n].fill_(float(rank + 1))
        w = train_utils.allreduce_gradients(fs.grads)
        ok = (w == world) and float(fs.grads.max()) == 3.0       # (alignment padding between variables stays 0)
        # views stayed views: the per-variable gradient tensors see the reduced values
        ok = ok and all(bool(torch.all(store.grads[n] == 3.0)) for n in fs.names)
        ok = ok and fs.names[0] == "crf_layer/transitions"        # group 0 sorted first
        # rank-sharded batches are different, timing reduction is a MAX
        a = synthetic.msra_batch(4, 16, seed=1234 + 100 * rank)["token_ids"]
        gathered = [torch.zeros_like(a) for _ in range(world)]
        dist.all_gather(gathered, a)
        ok = ok and not torch.equal(gathered[0], gathered[1])
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and float(t) == float(world)
        q.put((rank, ok, float(store.vars["logits/kernel"].sum())))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == res[1][2]          # identical initial weights across ranks
