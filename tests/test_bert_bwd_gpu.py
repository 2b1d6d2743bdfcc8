"""GPU: encoder backward kernels vs torch autograd (float64) of the same op."""
import math

import pytest
import torch

from chinesener_b200 import ops

pytestmark = pytest.mark.gpu


def _ln(z, g, b, eps):
    m = z.mean(-1, keepdim=True)
    v = ((z - m) ** 2).mean(-1, keepdim=True)
    return (z - m) * torch.rsqrt(v + eps) * g + b


@pytest.mark.parametrize("M,H,ybf16", [(200, 768, True), (77, 768, False), (33, 160, False)])
def test_layernorm_backward(M, H, ybf16):
    g = torch.Generator().manual_seed(M + H)
    y = torch.randn(M, H, generator=g)
    if ybf16:
        y = y.to(torch.bfloat16).float()
    r = torch.randn(M, H, generator=g)
    gam = 1 + 0.1 * torch.randn(H, generator=g)
    bet = 0.1 * torch.randn(H, generator=g)
    dout = torch.randn(M, H, generator=g)
    yd, rd, gd, bd = (t.double().requires_grad_(True) for t in (y, r, gam, bet))
    (_ln(yd + rd, gd, bd, 1e-12) * dout.double()).sum().backward()
    dgam = torch.zeros(H, device="cuda")
    dbet = torch.zeros(H, device="cuda")
    ycu = y.cuda().to(torch.bfloat16) if ybf16 else y.cuda()
    dz32, dz16 = ops.layernorm_bwd(ycu, gam.cuda(), dout.cuda(), dgam, dbet, residual=r.cuda(), eps=1e-12)
    torch.testing.assert_close(dz32.cpu().double(), yd.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dz32.cpu().double(), rd.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dz16.float().cpu().double(), yd.grad, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(dgam.cpu().double(), gd.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dbet.cpu().double(), bd.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("erf", [False, True])
def test_gelu_forward_backward(erf):
    g = torch.Generator().manual_seed(1)
    pre = (torch.randn(64, 3072, generator=g) * 2).to(torch.bfloat16)
    dact = torch.randn(64, 3072, generator=g).to(torch.bfloat16)
    x = pre.double().requires_grad_(True)
    act = torch.nn.functional.gelu(x, approximate="none" if erf else "tanh")
    (act * dact.double()).sum().backward()
    a = ops.gelu_bf16(pre.cuda(), erf=erf)
    d = ops.gelu_bwd_bf16(pre.cuda(), dact.cuda(), erf=erf)
    torch.testing.assert_close(a.float().cpu().double(), act.detach(), rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(d.float().cpu().double(), x.grad, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("M,N", [(1000, 200), (8192, 768), (333, 3072), (77, 51), (130, 66), (1, 8)])
def test_transpose_and_colsum_shapes(M, N):
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, N, generator=g).to(torch.bfloat16).cuda()
    Mp = (M + 7) // 8 * 8
    t = ops.transpose_bf16(x)
    assert t.shape == (N, Mp)
    assert torch.equal(t[:, :M], x.t().contiguous()) and bool((t[:, M:] == 0).all())
    cs = torch.zeros(N, device="cuda")
    ops.colsum_bf16_add(x, cs)
    torch.testing.assert_close(cs, x.float().sum(0), rtol=1e-4, atol=2e-3 * max(1.0, M ** 0.5))


def test_transpose_colsum_embed_bwd():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1000, 200, generator=g).to(torch.bfloat16).cuda()
    t = ops.transpose_bf16(x)
    assert t.shape == (200, 1000) and torch.equal(t, x.t().contiguous())
    cs = torch.zeros(200, device="cuda")
    ops.colsum_bf16_add(x, cs)
    torch.testing.assert_close(cs, x.float().sum(0), rtol=1e-4, atol=1e-3)
    # weight gradient on tensor cores from bf16 activations
    dy = torch.randn(1000, 96, generator=g).to(torch.bfloat16).cuda()
    dw = ops.wgrad_gemm_bf16(x, dy)
    ref = x.double().t() @ dy.double()
    assert (dw.double() - ref).abs().max() < 1e-3 * ref.abs().max() + 1e-3
    # embedding scatter-add
    B, L, H, V = 3, 7, 64, 50
    ids = torch.randint(0, V, (B, L), generator=g, dtype=torch.int32)
    seg = torch.randint(0, 2, (B, L), generator=g, dtype=torch.int32)
    dx = torch.randn(B * L, H, generator=g)
    dw_, dt_, dp_ = torch.zeros(V, H, device="cuda"), torch.zeros(2, H, device="cuda"), torch.zeros(16, H, device="cuda")
    ops.bert_embed_bwd(dx.cuda(), ids.cuda(), seg.cuda(), dw_, dt_, dp_)
    rw = torch.zeros(V, H).index_add_(0, ids.view(-1).long(), dx)
    rt = torch.zeros(2, H).index_add_(0, seg.view(-1).long(), dx)
    rp = torch.zeros(16, H).index_add_(0, torch.arange(L).repeat(B), dx)
    torch.testing.assert_close(dw_.cpu(), rw, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(dt_.cpu(), rt, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(dp_.cpu(), rp, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,L,NH", [(2, 128, 12), (3, 64, 4), (2, 150, 2), (1, 37, 3)])
def test_attention_backward(B, L, NH):
    D = 64
    g = torch.Generator().manual_seed(B * L + NH)
    qkv = (torch.randn(B * L, 3 * NH * D, generator=g) * 0.7).to(torch.bfloat16)
    dctx = torch.randn(B * L, NH * D, generator=g).to(torch.bfloat16)
    lens = torch.randint(1, L + 1, (B,), generator=g)
    lens[0] = L
    mask = (torch.arange(L)[None, :] < lens[:, None]).to(torch.int32)
    x = qkv.double().requires_grad_(True)
    q, k, v = x.view(B, L, 3, NH, D).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / math.sqrt(D) + (1.0 - mask.double())[:, None, None, :] * -10000.0
    ctx_ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, NH * D)
    (ctx_ref * dctx.double()).sum().backward()
    ctx = ops.bert_attention(qkv.cuda(), mask.cuda(), B, L, NH, D)
    dqkv = ops.bert_attention_bwd(qkv.cuda(), mask.cuda(), ctx, dctx.cuda(), B, L, NH, D)
    ref = x.grad
    err = (dqkv.float().cpu().double() - ref).abs().max().item()
    assert err < 3e-2 * ref.abs().max().item() + 1e-3, (err, ref.abs().max().item())


@pytest.mark.parametrize("B,L,NH", [(2, 128, 12), (2, 70, 3)])
def test_attention_probs_dropout_forward_and_backward(B, L, NH):
    """BertModel(is_training=True): dropout on the softmax output.  The kernels' counter-based mask is
    rebuilt on the host (tests/_masks.py) and fed to the float64 autograd reference."""
    from _masks import attention_keep
    D, keep, seed = 64, 0.9, (77 << 32) | 123456789
    g = torch.Generator().manual_seed(B * L + NH + 1)
    qkv = (torch.randn(B * L, 3 * NH * D, generator=g) * 0.7).to(torch.bfloat16)
    dctx = torch.randn(B * L, NH * D, generator=g).to(torch.bfloat16)
    lens = torch.randint(1, L + 1, (B,), generator=g)
    lens[0] = L
    mask = (torch.arange(L)[None, :] < lens[:, None]).to(torch.int32)
    z = torch.from_numpy(attention_keep(B, NH, L, keep, seed)).double() / keep
    assert 0.85 < float((z > 0).double().mean()) < 0.95
    x = qkv.double().requires_grad_(True)
    q, k, v = x.view(B, L, 3, NH, D).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / math.sqrt(D) + (1.0 - mask.double())[:, None, None, :] * -10000.0
    ctx_ref = ((torch.softmax(s, -1) * z) @ v).permute(0, 2, 1, 3).reshape(B * L, NH * D)
    (ctx_ref * dctx.double()).sum().backward()
    ctx = ops.bert_attention(qkv.cuda(), mask.cuda(), B, L, NH, D, keep_prob=keep, seed=seed)
    err = (ctx.float().cpu().double() - ctx_ref.detach()).abs().max().item()
    assert err < 2e-2 * ctx_ref.abs().max().item() + 1e-3, err
    dqkv = ops.bert_attention_bwd(qkv.cuda(), mask.cuda(), ctx, dctx.cuda(), B, L, NH, D, keep_prob=keep, seed=seed)
    ref = x.grad
    err = (dqkv.float().cpu().double() - ref).abs().max().item()
    assert err < 3e-2 * ref.abs().max().item() + 1e-3, (err, ref.abs().max().item())


def test_bf16_dropout_matches_host_mask_and_is_its_own_backward():
    from _masks import elementwise_keep
    n, keep, seed = 100003, 0.9, 987654321012
    x = torch.randn(n).to(torch.bfloat16)
    y = ops.dropout(x.cuda(), keep, seed)
    m = torch.from_numpy(elementwise_keep(n, keep, seed))
    ref = torch.where(m, (x.float() / keep).to(torch.bfloat16), torch.zeros((), dtype=torch.bfloat16))
    assert torch.equal(y.cpu(), ref)
    y32 = ops.dropout(x.float().cuda(), keep, seed)                # the f32 kernel shares the decisions
    assert torch.equal((y32 != 0).cpu() | (x.float() == 0), m | (x.float() == 0))
    xi = x.cuda().clone()
    assert ops.dropout(xi, keep, seed, inplace=True).data_ptr() == xi.data_ptr() and torch.equal(xi, y)


def test_layernorm_with_fused_hidden_dropout_forward_and_backward():
    """ner_layernorm_dropout / _bwd: LN(dropout(y) + r) with the host-rebuilt mask, against float64 autograd."""
    from _masks import elementwise_keep
    M, H, keep, seed = 150, 768, 0.9, 424242
    g = torch.Generator().manual_seed(7)
    y = torch.randn(M, H, generator=g).to(torch.bfloat16)
    r = torch.randn(M, H, generator=g)
    gam, bet = 1 + 0.1 * torch.randn(H, generator=g), 0.1 * torch.randn(H, generator=g)
    dout = torch.randn(M, H, generator=g)
    z = torch.from_numpy(elementwise_keep(M * H, keep, seed)).view(M, H).double() / keep
    yd, rd, gd, bd = (t.double().requires_grad_(True) for t in (y.float(), r, gam, bet))
    ref = _ln(yd * z + rd, gd, bd, 1e-12)
    (ref * dout.double()).sum().backward()
    o32, o16 = ops.layernorm(y.cuda(), gam.cuda(), bet.cuda(), residual=r.cuda(), eps=1e-12, keep_prob=keep, seed=seed)
    torch.testing.assert_close(o32.cpu().double(), ref.detach(), rtol=1e-4, atol=1e-4)
    dgam, dbet = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
    dz32, dz16 = ops.layernorm_bwd(y.cuda(), gam.cuda(), dout.cuda(), dgam, dbet, residual=r.cuda(), eps=1e-12, keep_prob=keep,
                                   seed=seed)
    torch.testing.assert_close(dz32.cpu().double(), rd.grad, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(dz16.float().cpu().double(), yd.grad, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(dgam.cpu().double(), gd.grad, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(dbet.cpu().double(), bd.grad, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("keep", [1.0, 0.9])
def test_layernorm_bwd_fused_bias_gradient(keep):
    """ner_layernorm_dropout_bwd_bias: d_bias += column sums of the masked dense-branch gradient == the separate
    ner_colsum_bf16_add pass over dz_bf16 it replaces (up to the bf16 rounding of that pass's input)."""
    M, H = 3150, 768
    g = torch.Generator().manual_seed(17)
    y = torch.randn(M, H, generator=g).to(torch.bfloat16).cuda()
    r = torch.randn(M, H, generator=g).cuda()
    gam = (1 + 0.1 * torch.randn(H, generator=g)).cuda()
    dout = torch.randn(M, H, generator=g).cuda()
    dgam, dbet, dbias = (torch.zeros(H, device="cuda") for _ in range(3))
    dbias += 1.0                                                    # accumulated into
    dz32, dz16 = ops.layernorm_bwd(y, gam, dout, dgam, dbet, residual=r, eps=1e-12, keep_prob=keep, seed=99, d_bias=dbias)
    dg2, db2 = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
    e32, e16 = ops.layernorm_bwd(y, gam, dout, dg2, db2, residual=r, eps=1e-12, keep_prob=keep, seed=99)
    assert torch.equal(dz32, e32) and torch.equal(dz16, e16)
    ref = torch.ones(H, device="cuda")
    ops.colsum_bf16_add(dz16, ref)
    scale = ref.abs().max().item()
    assert (dbias - ref).abs().max().item() < 5e-3 * max(1.0, scale)   # the old pass summed bf16-rounded values
    if keep == 1.0:                                                 # no mask: the dense-branch gradient is dz itself
        torch.testing.assert_close(dbias - 1.0, dz32.sum(0), rtol=1e-4, atol=1e-2)


def test_pack_group_equals_separate_packs():
    """ner_pack_weights_group_bf16: one launch writes the [N,K] packs and the TF-layout casts of a group of kernels
    (Q | K | V into one fused operand each) == ner_pack_weight_bf16 / ner_cast_bf16 per kernel."""
    g = torch.Generator().manual_seed(23)
    H, I = 768, 3072
    q, k, v, wi, wd, odd = (torch.randn(*s, generator=g).cuda() for s in ((H, H), (H, H), (H, H), (H, I), (I, H), (100, 72)))
    bf = lambda *s: torch.full(s, 7.0, dtype=torch.bfloat16, device="cuda")
    nk_qkv, kn_qkv = bf(3 * H, H), bf(H, 3 * H)
    nk_wi, kn_wi, nk_wd, kn_wd, nk_odd, kn_odd = bf(I, H), bf(H, I), bf(H, I), bf(I, H), bf(72, 100), bf(100, 72)
    triples = [(src, nk_qkv[j * H:(j + 1) * H], kn_qkv[:, j * H:(j + 1) * H]) for j, src in enumerate((q, k, v))]
    triples += [(wi, nk_wi, kn_wi), (wd, nk_wd, kn_wd), (odd, nk_odd, kn_odd), (wi, None, None)]
    grp = ops.PackGroup(triples)
    grp.run()
    wqkv = torch.cat([q, k, v], dim=1).contiguous()
    assert torch.equal(nk_qkv, ops.pack_weight_bf16(wqkv)) and torch.equal(kn_qkv, ops.cast_bf16(wqkv))
    assert torch.equal(nk_wi, ops.pack_weight_bf16(wi)) and torch.equal(kn_wi, ops.cast_bf16(wi))
    assert torch.equal(nk_wd, ops.pack_weight_bf16(wd)) and torch.equal(kn_wd, ops.cast_bf16(wd))
    assert torch.equal(nk_odd, odd.t().contiguous().to(torch.bfloat16)) and torch.equal(kn_odd, odd.to(torch.bfloat16))
    q.mul_(2.0)                                                      # same table, new values: re-run only
    grp.run()
    assert torch.equal(nk_qkv[:H], ops.pack_weight_bf16(q))


@pytest.mark.parametrize("erf", [False, True])
def test_gelu_bwd_with_fused_bias_gradient(erf):
    M, N = 3150, 3072
    g = torch.Generator().manual_seed(31)
    pre = torch.randn(M, N, generator=g).to(torch.bfloat16).cuda()
    dact = (torch.randn(M, N, generator=g) * 0.1).to(torch.bfloat16).cuda()
    dbias = torch.ones(N, device="cuda")
    got = ops.gelu_bwd_bias_bf16(pre, dact, dbias, erf=erf)
    want = ops.gelu_bwd_bf16(pre, dact, erf=erf)
    assert torch.equal(got, want)
    ref = torch.ones(N, device="cuda")
    ops.colsum_bf16_add(want, ref)
    torch.testing.assert_close(dbias, ref, rtol=1e-4, atol=1e-3)
