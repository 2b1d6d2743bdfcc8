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
