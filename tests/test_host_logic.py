"""CPU: host-side logic of the train ops and the lazy loss (no kernels involved)."""
import numpy as np
import torch

from chinesener_b200 import variables
from chinesener_b200.tools import train_utils


def test_noam_scheme_matches_reference_formula():
    # tools/transformer/modules.py:209-217: init_lr * w^0.5 * min(step * w^-1.5, step^-0.5), step = global_step + 1
    w, lr0 = 400, 1e-3
    for gs in (0, 1, 10, 399, 400, 401, 5000):
        step = gs + 1.0
        ref = lr0 * w ** 0.5 * min(step * w ** -1.5, step ** -0.5)
        assert abs(train_utils.noam_scheme(lr0, gs, w) - ref) < 1e-15
    assert abs(train_utils.noam_scheme(lr0, w - 1, w) - lr0) < 1e-12          # peak = init_lr at the end of warm-up
    assert train_utils.noam_scheme(lr0, 10, w) < train_utils.noam_scheme(lr0, 100, w) < train_utils.noam_scheme(lr0, 399, w)
    assert train_utils.noam_scheme(lr0, 399, w) > train_utils.noam_scheme(lr0, 4000, w)


def test_bert_lr_warmup_then_linear_decay():
    # create_optimizer (tools/train_utils.py:252-274): linear warm-up to init_lr, then polynomial (power 1) decay to 0
    lr0, n, nw = 5e-6, 1000, 100
    v = [train_utils.bert_lr(lr0, s, n, nw) for s in range(n + 1)]
    assert v[0] == 0.0 and abs(v[50] - lr0 * 0.5) < 1e-18
    assert abs(v[100] - lr0 * (1 - 100 / n)) < 1e-12 and abs(v[550] - lr0 * (1 - 550 / n)) < 1e-12
    assert v[n] == 0.0 and all(a >= b for a, b in zip(v[100:], v[101:]))


def test_staircase_decay():
    assert train_utils.lr_decay(1e-3, 0, 100, 0.95) == 1e-3
    assert abs(train_utils.lr_decay(1e-3, 250, 100, 0.95) - 1e-3 * 0.95 ** 2) < 1e-18


def test_deferred_is_evaluated_once_and_only_when_fetched():
    calls = []

    def thunk():
        calls.append(1)
        return torch.tensor([1.0, 3.0])
    ll = variables.Deferred(thunk)
    loss = (-ll).mean()                      # what every plugin builds: mean(-log_likelihood)
    assert calls == []                       # PREDICT never fetches it
    assert float(loss) == -2.0 and calls == [1]
    assert float(loss) == -2.0 and calls == [1]
    assert float((ll + 1.0).mean()) == 3.0 and float((2.0 * ll).mean()) == 4.0


def test_grad_exchange_buckets_cover_the_flat_buffer_in_backward_order(monkeypatch):
    """GradExchange (host logic, no GPU): the flat gradient buffer is cut into one contiguous bucket per encoder layer —
    layer 11 first, each waiting for its own layer's event — plus tail buckets for everything else; together they cover
    every element exactly once."""
    import torch
    from chinesener_b200 import bert, variables
    from chinesener_b200.tools import train_utils as tu

    class FakeEvent(object):
        cuda_event = 0

        def record(self, *a):
            pass

    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "Stream", lambda **kw: object())
    cfg = dict(bert.BERT_BASE_CHINESE, num_hidden_layers=4, vocab_size=50, hidden_size=64, intermediate_size=128,
               max_position_embeddings=16)
    st = variables.VariableStore("cpu")
    bert.create_bert_variables(cfg, st)
    st.get_variable("logits/kernel", (64, 10), variables.zeros)
    st.get_variable("crf_layer/transitions", (10, 10), variables.zeros)
    keys = ["crf", "logit", "lstm"]

    def group_of(name):
        for gi, k in enumerate(keys):
            if k in name:
                return (gi, 0 if tu._decays(name) else 1)
        return (len(keys), 0 if tu._decays(name) else 1)

    fs = tu.FlatState(st, group_of)
    ex = tu.GradExchange(fs)
    kinds = [b[0] for b in ex.buckets]
    assert kinds[:4] == [("layer", 0), ("layer", 1), ("layer", 2), ("layer", 3)]          # readiness order: layer 3 first
    assert all(k == ("tail",) for k in kinds[4:]) and len(kinds) >= 5
    for g, (key, s, e, ev) in enumerate(ex.buckets[:4]):
        names = [n for n in fs.names if s <= fs.slices[n][0] < e]
        assert names and all(f"/layer_{3 - g}/" in n and n.endswith("/kernel") for n in names)
        assert ev is ex.layer_events[3 - g]
    spans = sorted((s, e) for _, s, e, _ in ex.buckets)
    assert spans[0][0] == 0 and spans[-1][1] == fs.grads.numel()
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))                            # contiguous, no overlap
    assert all(ev is ex.tail_event for k, _, _, ev in ex.buckets if k == ("tail",))
