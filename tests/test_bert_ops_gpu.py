"""GPU: glue kernels and attention vs plain PyTorch fp32 references of the same op."""
import math

import pytest
import torch

from chinesener_b200 import ops

pytestmark = pytest.mark.gpu


def _ln(x, g, b, eps):
    m = x.mean(-1, keepdim=True)
    v = ((x - m) ** 2).mean(-1, keepdim=True)
    return (x - m) * torch.rsqrt(v + eps) * g + b


@pytest.mark.parametrize("B,L,H", [(4, 16, 768), (3, 7, 160), (2, 128, 1024), (5, 9, 64)])
def test_embed_ln(B, L, H):
    g = torch.Generator().manual_seed(B * L + H)
    V = 1000
    word = torch.randn(V, H, generator=g).cuda()
    typ = torch.randn(2, H, generator=g).cuda()
    pos = torch.randn(512, H, generator=g).cuda()
    gam = (1 + 0.1 * torch.randn(H, generator=g)).cuda()
    bet = (0.1 * torch.randn(H, generator=g)).cuda()
    ids = torch.randint(0, V, (B, L), generator=g, dtype=torch.int32).cuda()
    seg = torch.randint(0, 2, (B, L), generator=g, dtype=torch.int32).cuda()
    f32, b16 = ops.bert_embed_ln(word, typ, pos, gam, bet, ids, seg, eps=1e-12)
    ref = _ln(word[ids.long()] + typ[seg.long()] + pos[:L][None], gam, bet, 1e-12).view(B * L, H)
    torch.testing.assert_close(f32, ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(b16.float(), ref, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("M,H,eps", [(100, 768, 1e-12), (33, 160, 1.1920929e-07), (17, 320, 1e-5)])
def test_layernorm_with_residual(M, H, eps):
    g = torch.Generator().manual_seed(M + H)
    y = torch.randn(M, H, generator=g).cuda() * 3
    r = torch.randn(M, H, generator=g).cuda()
    gam = (1 + 0.1 * torch.randn(H, generator=g)).cuda()
    bet = (0.1 * torch.randn(H, generator=g)).cuda()
    f32, b16 = ops.layernorm(y, gam, bet, residual=r, eps=eps)
    ref = _ln(y + r, gam, bet, eps)
    torch.testing.assert_close(f32, ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(b16.float(), ref, rtol=1e-2, atol=1e-2)
    f32b, _ = ops.layernorm(y, gam, bet, eps=eps, want_bf16=False)
    torch.testing.assert_close(f32b, _ln(y, gam, bet, eps), rtol=1e-5, atol=1e-5)
    y16 = y.to(torch.bfloat16)                                     # bf16 dense output + f32 residual
    f32c, _ = ops.layernorm(y16, gam, bet, residual=r, eps=eps)
    torch.testing.assert_close(f32c, _ln(y16.float() + r, gam, bet, eps), rtol=1e-5, atol=1e-5)


def test_pack_cast_pad_lookup():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(250, 1600, generator=g).cuda()
    assert torch.equal(ops.pack_weight_bf16(w), w.t().contiguous().to(torch.bfloat16))
    x = torch.randn(1000, 37, generator=g).cuda()
    assert torch.equal(ops.cast_bf16(x), x.to(torch.bfloat16))
    p = ops.cast_pad_bf16(x, 40)
    assert torch.equal(p[:, :37], x.to(torch.bfloat16)) and (p[:, 37:] == 0).all()
    table = torch.randn(500, 50, generator=g).cuda()
    ids = torch.randint(0, 500, (6, 11), generator=g, dtype=torch.int32).cuda()
    assert torch.equal(ops.embedding_lookup(table, ids), table[ids.long()])
    buf = torch.zeros(6, 11, 70, device="cuda")
    ops.embedding_lookup(table, ids, out=buf, col_offset=20)
    assert torch.equal(buf[..., 20:], table[ids.long()]) and (buf[..., :20] == 0).all()


@pytest.mark.parametrize("M,F,N,bf16", [(8192, 256, 10, False), (100, 768, 10, True), (33, 400, 7, False),
                                        (50, 160, 20, False), (5, 64, 32, True)])
def test_dense_small_n(M, F, N, bf16):
    g = torch.Generator().manual_seed(M + F + N)
    x = torch.randn(M, F, generator=g).cuda()
    if bf16:
        x = x.to(torch.bfloat16)
    w = (torch.randn(F, N, generator=g) * 0.1).cuda()
    b = torch.randn(N, generator=g).cuda()
    out = ops.dense_small_n(x, w, b)
    ref = x.float() @ w + b
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("B,L,NH", [(2, 128, 12), (3, 150, 4), (1, 64, 2), (2, 37, 3), (1, 256, 2)])
def test_attention(B, L, NH):
    D = 64
    g = torch.Generator().manual_seed(B * L + NH)
    qkv = (torch.randn(B * L, 3 * NH * D, generator=g)).to(torch.bfloat16).cuda()
    lens = torch.randint(1, L + 1, (B,), generator=g)
    lens[0] = L
    mask = (torch.arange(L)[None, :] < lens[:, None]).to(torch.int32).cuda()
    ctx = ops.bert_attention(qkv, mask, B, L, NH, D)
    q, k, v = qkv.float().view(B, L, 3, NH, D).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / math.sqrt(D) + (1.0 - mask.float())[:, None, None, :] * -10000.0
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, NH * D)
    torch.testing.assert_close(ctx.float(), ref, rtol=2e-2, atol=2e-2)


def test_seq_pack_plan_and_packed_attention():
    B, L, NH, D = 5, 70, 3, 64
    g = torch.Generator().manual_seed(5)
    lens = torch.tensor([70, 1, 33, 64, 17])
    mask = (torch.arange(L)[None, :] < lens[:, None]).to(torch.int32).cuda()
    cu, tok_src = ops.seq_pack_plan(mask)
    assert cu.cpu().tolist() == [0, 70, 71, 104, 168, 185]
    exp = torch.cat([b * L + torch.arange(int(n)) for b, n in enumerate(lens)])
    assert torch.equal(tok_src[:185].cpu().long(), exp)
    qkv = torch.randn(B * L, 3 * NH * D, generator=g).to(torch.bfloat16).cuda()
    ref = ops.bert_attention(qkv, mask, B, L, NH, D)                    # padded
    qkv_p = qkv[tok_src[:185].long()].contiguous()
    out = ops.bert_attention(qkv_p, None, B, L, NH, D, cu_seqlens=cu)   # packed
    torch.testing.assert_close(out.float(), ref[tok_src[:185].long()].float(), rtol=2e-2, atol=2e-2)
