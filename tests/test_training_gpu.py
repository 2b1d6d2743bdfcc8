"""GPU: TRAIN mode of the bilstm_crf plugin — gradients vs autograd of the float64 oracle, one
optimizer step vs the reference's Adam formulas, and a short loss-goes-down run."""
import numpy as np
import pytest
import torch

from chinesener_b200 import autodiff, engine, synthetic, variables
from oracle import crf_torch, nn as onn, optim

pytestmark = pytest.mark.gpu


def _setup(B=8, L=64, V=2000, seed=3, dropout=0.0, keep=1.0):
    feats = synthetic.msra_batch(B, L, vocab=V, seed=seed)
    g = torch.Generator().manual_seed(0)
    emb = torch.nn.functional.normalize(torch.randn(V, 50, generator=g), dim=1).numpy()
    params = dict(synthetic.data_params(L), embedding=emb, embedding_dropout=dropout, keep_prob_list=[keep])
    est = engine.Estimator("bilstm_crf", params)
    return est, feats, emb


def _oracle_grads(w, feats, emb, act):
    wd = {k: v.double().clone().requires_grad_(True) for k, v in w.items()}
    x = torch.from_numpy(emb).double()[feats['token_ids'].long()]
    lstm = onn.bilstm(x, wd, feats['seq_len'], act, 1.0, torch.float64)
    logits = lstm @ wd['logits/kernel'] + wd['logits/bias']
    ll = crf_torch.crf_log_likelihood(logits, feats['label_ids'], feats['seq_len'], wd['crf_layer/transitions'])
    loss = (-ll).mean()
    loss.backward()
    return float(loss), {k: v.grad for k, v in wd.items()}


def test_bilstm_crf_gradients_match_oracle_autograd():
    est, feats, emb = _setup()
    est.evaluate(feats)                                    # creates the variables
    w = est.store.state_dict()
    ref_loss, ref = _oracle_grads(w, feats, emb, est.params['rnn_activation'])
    dev = est.to_device(feats)
    with variables.use_store(est.store), autodiff.recording(est.store) as tape:
        loss, _ = est.build_graph(dev, None, est.params, True)
        tape.backward()
    assert abs(float(loss) - ref_loss) < 2e-3 * max(1.0, abs(ref_loss))
    for name, g_ref in ref.items():
        g = est.store.grads[name].cpu().double()
        scale = max(g_ref.abs().max().item(), 1e-6)
        err = (g - g_ref).abs().max().item()
        # bf16 operands in the input-projection / weight-gradient GEMMs: 2e-2 of the gradient scale
        assert err < 2e-2 * scale, (name, err, scale)


def test_one_train_step_equals_reference_adam():
    est, feats, emb = _setup()
    est.evaluate(feats)
    w0 = {k: v.clone() for k, v in est.store.state_dict().items()}
    _, ref = _oracle_grads(w0, feats, emb, est.params['rnn_activation'])
    est.train_step(feats)
    w1 = est.store.state_dict()
    for name in ("crf_layer/transitions", "logits/kernel", "logits/bias"):
        g = ref[name].numpy()
        p, _, _ = optim.tf_adam_step(w0[name].double().numpy(), g, 0.0, 0.0, lr=est.params['lr'], t=1)
        # the first Adam step is lr*sign(g): skip elements whose gradient is within noise of zero
        ok = np.abs(g) > 1e-3 * np.abs(g).max()
        np.testing.assert_allclose(w1[name].numpy()[ok], p[ok], rtol=1e-3, atol=2e-4)
    assert est.store.global_step == 1


def test_training_reduces_the_loss_with_dropout_on():
    est, feats, emb = _setup(dropout=0.3, keep=0.8)      # embedding dropout + DropoutWrapper(0.8)
    losses = [float(est.train_step(feats)) for _ in range(25)]
    assert losses[-1] < 0.7 * losses[0], losses
    ev = est.evaluate(feats)
    assert np.isfinite(ev['loss'])
