"""GPU: training-side kernels vs autograd of the float64 oracle (the reference's tf.gradients)."""
import numpy as np
import pytest
import torch

from chinesener_b200 import ops
from oracle import crf, crf_torch, nn as onn, optim

pytestmark = pytest.mark.gpu


def _lstm_w(D, H, seed):
    g = torch.Generator().manual_seed(seed)
    w = {}
    for d in ("fw", "bw"):
        lim = (6.0 / (D + 5 * H)) ** 0.5
        w[f"bilstm_layer/bidirectional_rnn/{d}/multi_rnn_cell/cell_0/lstm_cell/kernel"] = \
            ((torch.rand(D + H, 4 * H, generator=g) * 2 - 1) * lim).double()
        w[f"bilstm_layer/bidirectional_rnn/{d}/multi_rnn_cell/cell_0/lstm_cell/bias"] = (torch.randn(4 * H, generator=g) * 0.1).double()
    return w


@pytest.mark.parametrize("B,L,H,act", [(6, 20, 128, "tanh"), (64, 48, 128, "relu"), (5, 17, 200, "tanh"), (3, 9, 64, "relu"),
                                       (150, 8, 128, "tanh")])
def test_bilstm_bptt_matches_autograd(B, L, H, act):
    D = 24
    g = torch.Generator().manual_seed(B + L + H)
    x = torch.randn(B, L, D, generator=g, dtype=torch.float64, requires_grad=True)
    w = {k: v.clone().requires_grad_(True) for k, v in _lstm_w(D, H, H).items()}
    lens = torch.randint(1, L + 1, (B,), generator=g, dtype=torch.int32)
    lens[0] = L
    if B > 2:
        lens[1] = 1
        lens[2] = 0
    d_out = torch.randn(B, L, 2 * H, generator=g, dtype=torch.float64)
    out_ref = onn.bilstm(x, w, lens, act, 1.0, torch.float64)
    (out_ref * d_out).sum().backward()
    names = [f"bilstm_layer/bidirectional_rnn/{d}/multi_rnn_cell/cell_0/lstm_cell/" for d in ("fw", "bw")]
    ks = [w[n + "kernel"].detach() for n in names]
    bs = [w[n + "bias"].detach() for n in names]
    xproj = torch.cat([x.detach().view(B * L, D) @ k[:D] + b for k, b in zip(ks, bs)], dim=1).float().cuda()
    whf, whb = ks[0][D:].float().contiguous().cuda(), ks[1][D:].float().contiguous().cuda()
    out, gates, cst, hst = ops.bilstm_recurrence(xproj, whf, whb, lens.cuda(), B, L, H, activation=act, save_for_backward=True)
    assert torch.equal(hst, out)                     # keep_prob 1: carried h == emitted output
    torch.testing.assert_close(out.cpu().double(), out_ref.detach(), rtol=1e-4, atol=1e-4)
    dxp = ops.bilstm_recurrence_bwd(d_out.float().cuda(), gates, cst, whf, whb, lens.cuda(), B, L, H, activation=act)
    dxp = dxp.cpu().double()
    x2 = x.detach().view(B * L, D)
    o = out.cpu().double()
    for di, n in enumerate(names):
        dz = dxp[:, di * 4 * H:(di + 1) * 4 * H]
        gk = w[n + "kernel"].grad
        torch.testing.assert_close(x2.t() @ dz, gk[:D], rtol=2e-3, atol=2e-4)            # dW_x
        torch.testing.assert_close(dz.sum(0), w[n + "bias"].grad, rtol=2e-3, atol=2e-4)   # d_bias
        hprev = torch.zeros(B, L, H, dtype=torch.float64)
        if di == 0:
            hprev[:, 1:] = o[:, :-1, :H]
        else:
            hprev[:, :-1] = o[:, 1:, H:]
        torch.testing.assert_close(hprev.view(B * L, H).t() @ dz, gk[D:], rtol=2e-3, atol=2e-4)  # dW_h
    dx = dxp[:, :4 * H] @ ks[0][:D].t() + dxp[:, 4 * H:] @ ks[1][:D].t()
    torch.testing.assert_close(dx.view(B, L, D), x.grad, rtol=2e-3, atol=2e-4)
    # padded steps carry no gradient
    valid = torch.arange(L)[None, :] < lens[:, None]
    assert (dxp.view(B, L, -1)[~valid] == 0).all()


def test_dense_small_n_backward_and_wgrad_gemm():
    g = torch.Generator().manual_seed(0)
    M, F, N = 3000, 256, 10
    x = torch.randn(M, F, generator=g)
    w = torch.randn(F, N, generator=g) * 0.1
    dy = torch.randn(M, N, generator=g)
    dW = torch.zeros(F, N, device="cuda")
    db = torch.zeros(N, device="cuda")
    dx = ops.dense_small_n_bwd(x.cuda(), w.cuda(), dy.cuda(), dW, db)
    torch.testing.assert_close(dW.cpu().double(), x.double().t() @ dy.double(), rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(db.cpu().double(), dy.double().sum(0), rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(dx.cpu().double(), dy.double() @ w.double().t(), rtol=1e-4, atol=1e-4)
    # tensor-core weight gradient (bf16 operands): x^T dy for a [M,50] x [M,1024] pair
    a = torch.randn(M, 50, generator=g)
    b = torch.randn(M, 1024, generator=g) * 0.1
    res = ops.wgrad_gemm(a.cuda(), b.cuda())
    ref = a.double().t() @ b.double()
    assert (res.cpu().double() - ref).abs().max() < 2e-2 * ref.abs().max()
    cs = torch.zeros(1024, device="cuda")
    ops.colsum_add(b.cuda(), cs, 0.5)
    torch.testing.assert_close(cs.cpu().double(), 0.5 * b.double().sum(0), rtol=1e-4, atol=1e-3)


def test_dropout_is_counter_based():
    x = torch.ones(1 << 20, device="cuda")
    y = ops.dropout(x, 0.7, seed=1234)
    kept = (y != 0).float().mean().item()
    assert abs(kept - 0.7) < 5e-3
    assert torch.allclose(y[y != 0], torch.tensor(1 / 0.7, device="cuda"))
    assert torch.equal(y, ops.dropout(x, 0.7, seed=1234))           # same seed -> same mask (backward reuses it)
    assert not torch.equal(y, ops.dropout(x, 0.7, seed=1235))
    g = torch.randn(1 << 20, device="cuda")
    assert torch.equal(ops.dropout(g, 0.7, seed=1234) != 0, (y != 0) & (g != 0))


def test_optimizer_steps_match_reference_formulas():
    rng = np.random.default_rng(0)
    n = 5000
    p, g, m, v = (rng.normal(size=n).astype(np.float32) for _ in range(4))
    v = np.abs(v)
    # tf.train.AdamOptimizer + clip_by_value(5)
    g2 = g * 10
    pr, mr, vr = optim.tf_adam_step(p.astype(np.float64), g2.astype(np.float64), m.astype(np.float64), v.astype(np.float64),
                                    lr=0.005, t=7)
    lr_t = 0.005 * np.sqrt(1 - 0.999 ** 7) / (1 - 0.9 ** 7)
    P, G, M_, V = (torch.from_numpy(a.copy()).cuda() for a in (p, g2, m, v))
    ops.adam_step(P, G, M_, V, lr=lr_t, eps=1e-8, mode=1, clip=5.0)
    np.testing.assert_allclose(P.cpu().numpy(), pr, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(M_.cpu().numpy(), mr, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(V.cpu().numpy(), vr, rtol=1e-5, atol=1e-6)
    # AdamWeightDecayOptimizer after clip_by_global_norm(1.0)
    (gc,), gn = optim.clip_by_global_norm([g.astype(np.float64)], 1.0)
    pr, mr, vr = optim.adam_weight_decay_step(p.astype(np.float64), gc, m.astype(np.float64), v.astype(np.float64), lr=5e-6 * 500,
                                              name="crf_layer/transitions")
    P, G, M_, V = (torch.from_numpy(a.copy()).cuda() for a in (p, g, m, v))
    gsq = torch.zeros(1, device="cuda")
    ops.sumsq_add(G, gsq)
    assert abs(gsq.item() ** 0.5 - gn) < 1e-3 * gn
    ops.adam_step(P, G, M_, V, lr=5e-6 * 500, eps=1e-6, weight_decay=0.01, mode=0, clip=1.0, gnorm_sq=gsq)
    np.testing.assert_allclose(P.cpu().numpy(), pr, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(V.cpu().numpy(), vr, rtol=1e-5, atol=1e-6)


def test_crf_torch_oracle_agrees_with_numpy_oracle():
    rng = np.random.default_rng(1)
    x = rng.normal(size=(7, 9, 5)); tr = rng.normal(size=(5, 5)); lens = np.array([9, 1, 0, 4, 9, 3, 2])
    tags = rng.integers(0, 5, size=(7, 9))
    a = crf.crf_log_likelihood(x, tags, lens, tr)
    b = crf_torch.crf_log_likelihood(torch.from_numpy(x), torch.from_numpy(tags), torch.from_numpy(lens), torch.from_numpy(tr))
    np.testing.assert_allclose(b.numpy(), a, rtol=1e-10, atol=1e-10)


def test_bilstm_dropout_wrapper_masks_are_consistent():
    """keep_prob < 1: output / state masks are independent Bernoulli(keep) scaled by 1/keep, and the
    backward kernel regenerates them (finite-difference check of one loss through the kernels)."""
    B, L, H, D = 4, 12, 128, 16
    g = torch.Generator().manual_seed(3)
    w = _lstm_w(D, H, 1)
    names = [f"bilstm_layer/bidirectional_rnn/{d}/multi_rnn_cell/cell_0/lstm_cell/" for d in ("fw", "bw")]
    ks = [w[n + "kernel"] for n in names]
    lens = torch.full((B,), L, dtype=torch.int32)
    xproj = torch.randn(B * L, 8 * H, generator=g).cuda()
    whf, whb = ks[0][D:].float().contiguous().cuda(), ks[1][D:].float().contiguous().cuda()
    out1, gates, cst, hst = ops.bilstm_recurrence(xproj, whf, whb, lens.cuda(), B, L, H, save_for_backward=True, keep_prob=1.0)
    out, gates, cst, hst = ops.bilstm_recurrence(xproj, whf, whb, lens.cuda(), B, L, H, save_for_backward=True, keep_prob=0.8, seed=77)
    kept = (out != 0).float().mean().item()
    assert abs(kept - 0.8) < 0.03
    assert (out != 0).ne(hst != 0).float().mean().item() > 0.2          # independent masks
    d_out = torch.randn(B, L, 2 * H, generator=g).cuda()
    dxp = ops.bilstm_recurrence_bwd(d_out, gates, cst, whf, whb, lens.cuda(), B, L, H, keep_prob=0.8, seed=77)
    # directional finite difference of  sum(out * d_out)  w.r.t. xproj
    v = torch.randn_like(xproj)
    eps = 1e-2
    f = lambda xp: (ops.bilstm_recurrence(xp, whf, whb, lens.cuda(), B, L, H, keep_prob=0.8, seed=77).double() * d_out.double()).sum().item()
    fd = (f(xproj + eps * v) - f(xproj - eps * v)) / (2 * eps)
    an = (dxp.double() * v.double()).sum().item()
    assert abs(fd - an) < 2e-2 * max(1.0, abs(an)), (fd, an)
