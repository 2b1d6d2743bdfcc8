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


def _softlex_setup(dropout=0.0, keep=1.0, B=8, L=32, V=3000, NW=5000):
    feats = synthetic.msra_batch(B, L, vocab=V, seed=4)
    ids, wts = synthetic.softlexicon_features(B, L, NW, seed=4, lens=feats['seq_len'].numpy())
    feats['softlexicon_ids'], feats['softlexicon_weights'] = ids, wts
    g = torch.Generator().manual_seed(1)
    emb = torch.nn.functional.normalize(torch.randn(V, 50, generator=g), dim=1).numpy()
    wemb = torch.nn.functional.normalize(torch.randn(NW, 50, generator=g), dim=1).numpy()
    params = dict(synthetic.data_params(L), embedding=emb, word_embedding=wemb, word_enhance_dim=4, max_lexicon_len=10,
                  embedding_dropout=dropout, keep_prob_list=[keep])
    return engine.Estimator("bilstm_crf_softlexicon", params), feats, emb


def test_softlexicon_gradients_match_oracle_autograd():
    """TRAIN mode of bilstm_crf_softlexicon (BASELINE config 4): the lexicon table's gradient is the scatter-add of
    the pool backward; every variable against float64 autograd of the oracle."""
    est, feats, emb = _softlex_setup()
    est.evaluate(feats)
    w = est.store.state_dict()
    wd = {k: v.double().clone().requires_grad_(True) for k, v in w.items()}
    B, L = feats['token_ids'].shape
    G, S = 4, 10
    x = torch.from_numpy(emb).double()[feats['token_ids'].long()]
    wh = onn.softlexicon_pool(wd['word_enhance/softlexicon_embedding'], feats['softlexicon_ids'].view(B, L, G * S),
                              feats['softlexicon_weights'].view(B, L, G * S).double(), G, S)
    lstm = onn.bilstm(torch.cat([wh, x], -1), wd, feats['seq_len'], est.params['rnn_activation'], 1.0, torch.float64)
    logits = lstm @ wd['logits/kernel'] + wd['logits/bias']
    ll = crf_torch.crf_log_likelihood(logits, feats['label_ids'], feats['seq_len'], wd['crf_layer/transitions'])
    ref_loss = (-ll).mean()
    ref_loss.backward()
    dev = est.to_device(feats)
    with variables.use_store(est.store), autodiff.recording(est.store) as tape:
        loss, _ = est.build_graph(dev, None, est.params, True)
        tape.backward()
    assert abs(float(loss) - float(ref_loss)) < 2e-3 * max(1.0, abs(float(ref_loss)))
    for name, v in wd.items():
        g_ref = v.grad
        g = est.store.grads[name].cpu().double()
        scale = max(g_ref.abs().max().item(), 1e-6)
        assert (g - g_ref).abs().max().item() < 2e-2 * scale, name


def test_softlexicon_training_reduces_the_loss():
    est, feats, _ = _softlex_setup(dropout=0.5, keep=0.9)
    losses = [float(est.train_step(feats)) for _ in range(30)]
    assert np.isfinite(losses).all() and losses[-1] < 0.7 * losses[0], losses
