"""GPU: contracts around the plugins — the checkpoint variable set, checkpoint round trips, the per-sentence
PREDICT stream that evaluation.py consumes, and InferHelper over the multi-task / word-enhance plugins."""
import json
import os

import numpy as np
import pytest
import torch

from chinesener_b200 import checkpoint, engine, synthetic, tf_checkpoint
from chinesener_b200.config import BERT_BASE_CHINESE

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
INDEX = json.load(open(os.path.join(GOLD, "variables_index.json")))


def _names_and_shapes(est):
    return {k: list(v.shape) for k, v in est.store.vars.items()}


def _expect(model):
    g = {k: v["shape"] for k, v in INDEX[model]["variables"].items()}
    assert g.pop("global_step") == []             # the store keeps it as an attribute, the checkpoint as an int64 scalar
    return g


def _char_table(V=11329, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(V, 50, generator=g), dim=1).numpy()


def test_bert_bilstm_crf_variable_set_is_the_serving_checkpoints(tmp_path):
    """names + shapes of every variable the plugin creates == serving_model/bert_bilstm_crf/1/variables/variables.index
    (207 entries, 412 755 392 bytes of fp32)."""
    (tmp_path / "bert_config.json").write_text(json.dumps(BERT_BASE_CHINESE))
    L = 32
    est = engine.Estimator("bert_bilstm_crf", dict(synthetic.data_params(L), pretrain_dir=str(tmp_path)))
    est.evaluate(synthetic.msra_batch(2, L, seed=1))
    assert _names_and_shapes(est) == _expect("bert_bilstm_crf")
    assert sum(v.numel() * 4 for v in est.store.vars.values()) + 8 == INDEX["bert_bilstm_crf"]["total_bytes"]
    # 204 variables receive gradients (the 2 pooler variables do not)
    assert len(est.store.trainable_names()) == 204


def test_bilstm_crf_and_softlexicon_variable_sets(tmp_path):
    L = 32
    est = engine.Estimator("bilstm_crf", dict(synthetic.data_params(L), embedding=_char_table()))
    est.evaluate(synthetic.msra_batch(2, L, vocab=11329, seed=1))
    assert _names_and_shapes(est) == _expect("bilstm_crf")          # the embedding is a constant, not a checkpoint variable
    NW = 704370
    g = torch.Generator().manual_seed(2)
    wemb = torch.randn(NW, 50, generator=g).numpy()
    feats = synthetic.msra_batch(2, L, vocab=11329, seed=1)
    feats['softlexicon_ids'], feats['softlexicon_weights'] = synthetic.softlexicon_features(2, L, NW, seed=1, lens=feats['seq_len'].numpy())
    est = engine.Estimator("bilstm_crf_softlexicon", dict(synthetic.data_params(L), embedding=_char_table(), word_embedding=wemb,
                                                          word_enhance_dim=4, max_lexicon_len=10))
    est.evaluate(feats)
    assert _names_and_shapes(est) == _expect("bilstm_crf_softlexicon")


def test_mtl_variable_set(tmp_path):
    """serving_model/bert_bilstm_crf_mtl: task_list = [msr (7 CWS labels), msra (10)], asymmetry -> msra/logits/kernel [512, 10]."""
    (tmp_path / "bert_config.json").write_text(json.dumps(BERT_BASE_CHINESE))
    L = 32
    base = synthetic.data_params(L)
    msr = dict(label_size=7, idx2tag={i: str(i) for i in range(7)})
    params = dict(base, pretrain_dir=str(tmp_path), task_list=['msr', 'msra'], msr=msr,
                  msra=dict(label_size=10, idx2tag=base['idx2tag']), asymmetry=True)
    feats = synthetic.msra_batch(2, L, seed=3)
    feats['task_ids'] = torch.tensor([0, 1], dtype=torch.int32)
    feats['label_ids'][0] = feats['label_ids'][0] % 7
    est = engine.Estimator("bert_bilstm_crf_mtl", params)
    est.evaluate(feats)
    assert _names_and_shapes(est) == _expect("bert_bilstm_crf_mtl")


def test_checkpoint_round_trips_npz_and_tf_bundle(tmp_path):
    """train a few steps -> save -> fresh Estimator restores variables, Adam moments and global_step (same next step);
    the variables also survive a TensorFlow tensor-bundle export / import (the format of the reference's checkpoints)."""
    L = 48
    params = dict(synthetic.data_params(L), embedding=_char_table(2000), embedding_dropout=0.0)
    feats = [synthetic.msra_batch(8, L, vocab=2000, seed=s) for s in range(4)]
    a = engine.Estimator("bilstm_crf", dict(params))
    for f in feats[:3]:
        a.train_step(f)
    path = checkpoint.save_checkpoint(a.store, str(tmp_path / "ckpt"))
    torch.save(a.store.state_dict(), str(tmp_path / "vars.pt"))
    assert path.endswith("model.ckpt-3.npz")
    b = engine.Estimator("bilstm_crf", dict(params))
    b.evaluate(feats[0])                                  # creates (differently initialised) variables
    assert checkpoint.restore_checkpoint(b.store, checkpoint.latest_checkpoint(str(tmp_path / "ckpt"))) == 3
    la, lb = a.train_step(feats[3]), b.train_step(feats[3])
    assert abs(float(la) - float(lb)) < 1e-5 * max(1.0, abs(float(la))) and b.store.global_step == 4
    for k in a.store.vars:        # the same 4th update (atomic float accumulation orders differ between two runs: not bit-equal):
        assert torch.allclose(a.store.vars[k], b.store.vars[k], rtol=1e-4, atol=1e-6), k      # the moments and the step came along
    c0 = engine.Estimator("bilstm_crf", dict(params))      # control: variables only, zero moments, step 0 -> a different update
    c0.evaluate(feats[0])
    c0.store.load_state_dict({k: v for k, v in torch.load(str(tmp_path / "vars.pt")).items()}, strict=True)
    c0.train_step(feats[3])
    assert not torch.allclose(a.store.vars["logits/kernel"], c0.store.vars["logits/kernel"], rtol=1e-4, atol=1e-6)
    prefix = str(tmp_path / "export" / "model.ckpt")
    os.makedirs(os.path.dirname(prefix))
    tensors = {k: v.numpy() for k, v in a.store.state_dict().items()}
    tensors["global_step"] = np.asarray(a.store.global_step, np.int64)
    tf_checkpoint.save_tf_checkpoint(prefix, tensors)
    back = tf_checkpoint.load_tf_checkpoint(prefix, verify=True)
    c = engine.Estimator("bilstm_crf", dict(params))
    c.evaluate(feats[0])
    c.store.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in back.items() if k != "global_step"}, strict=True)
    assert torch.equal(c.predict(feats[1])['pred_ids'], a.predict(feats[1])['pred_ids']) and int(back["global_step"]) == 4


def test_predict_sentences_feeds_single_eval(tmp_path):
    """Estimator.predict_sentences -> the per-sentence dicts of `<model>_predict.pkl` -> evaluation.SingleEval."""
    from chinesener_b200.evaluation import SingleEval
    from chinesener_b200.main import predict_to_list
    L, B = 40, 6
    params = dict(synthetic.data_params(L), embedding=_char_table(500))
    est = engine.Estimator("bilstm_crf", params)
    batches = []
    for s in range(3):
        f = synthetic.msra_batch(B, L, vocab=500, seed=s)
        f['label_ids'] = torch.where(f['label_ids'] >= 8, torch.ones_like(f['label_ids']), f['label_ids'])   # giga: no [CLS]/[SEP]
        f['tokens'] = [["字"] * int(n) + ["[PAD]"] * (L - int(n)) for n in f['seq_len']]
        batches.append(f)
    pred = predict_to_list(est, lambda: iter(batches))
    assert len(pred) == 3 * B and pred[0]['pred_ids'].shape == (L,) and pred[0]['pred_ids'].dtype == np.int32
    assert pred[0]['tokens'].dtype == object and isinstance(pred[0]['tokens'][0], bytes)
    whole = torch.cat([est.predict(f)['pred_ids'] for f in batches]).numpy()
    assert np.array_equal(np.stack([p['pred_ids'] for p in pred]), whole)
    tag_rep, ent_rep = SingleEval(pred, params['idx2tag']).gen_report()
    assert ent_rep['micro avg']['support'] > 0 and 0.0 <= ent_rep['micro avg']['f1-score'] <= 1.0
    # gold labels scored against themselves -> F1 1.0 through the same path
    gold = [dict(p, pred_ids=p['label_ids']) for p in pred]
    assert SingleEval(gold, params['idx2tag']).entity_eval()['micro avg']['f1-score'] == 1.0


def test_infer_helper_on_mtl_and_softlexicon_plugins(tmp_path):
    """InferHelper picks the processor from the model name and carries task_ids / softlexicon features."""
    from chinesener_b200.data.tokenizer import FullTokenizer, TokenizerAdapter
    from chinesener_b200.data.word_enhance import WordVocab
    from chinesener_b200.inference import InferHelper, TAG2IDX
    from chinesener_b200.tools.infer_utils import extract_entity
    gold = json.load(open(os.path.join(GOLD, "warmup_features.json"), encoding="utf8"))
    idx2tag = {v: k for k, v in TAG2IDX.items()}
    L = 64
    # --- multi-task plugin: task_ids = 1 rides along
    small = dict(BERT_BASE_CHINESE, num_hidden_layers=2)
    (tmp_path / "bert_config.json").write_text(json.dumps(small))
    vocab = dict(gold["bert_vocab_subset"])
    vocab.setdefault("[UNK]", 100)
    base = synthetic.data_params(L)
    cws = dict(label_size=5, idx2tag={i: t for i, t in enumerate(['[PAD]', 'B', 'M', 'E', 'S'])})
    params = dict(base, pretrain_dir=str(tmp_path), task_list=['cws', 'msra'], cws=cws,
                  msra=dict(label_size=10, idx2tag=base['idx2tag']), asymmetry=True)
    est = engine.Estimator("bert_bilstm_crf_mtl", params)
    helper = InferHelper(L, TAG2IDX, "bert_bilstm_crf_mtl", FullTokenizer(vocab), estimator=est)
    assert helper.mtl == 1 and helper.word_enhance is None
    ent = helper.infer(gold["text"])
    assert helper.feature['task_ids'] == 1
    from chinesener_b200.data.base_preprocess import features_to_batch
    pred = est.predict(features_to_batch([helper.feature]))['pred_ids'].numpy()[0]
    assert dict(ent) == dict(extract_entity(helper.feature['tokens'], [int(i) for i in pred], idx2tag))
    # --- word-enhance plugin: SoftLexiconProc builds softlexicon_ids / weights
    chars = sorted(set(gold["text"]))
    words = chars + [gold["text"][i:i + 2] for i in range(0, len(gold["text"]) - 1, 3)]
    wv = WordVocab(words, {w: 3 + i for i, w in enumerate(words)})
    g = torch.Generator().manual_seed(0)
    params = dict(synthetic.data_params(L), embedding=torch.randn(len(chars) + 2, 50, generator=g).numpy(),
                  word_embedding=torch.randn(wv.n_word + 3, 50, generator=g).numpy(), word_enhance_dim=4, max_lexicon_len=10)
    est = engine.Estimator("bilstm_crf_softlexicon", params)
    helper = InferHelper(L, TAG2IDX, "bilstm_crf_softlexicon", TokenizerAdapter(chars), estimator=est, vocab=wv)
    assert helper.word_enhance == 'softlexicon' and type(helper.proc).__name__ == 'SoftLexiconProc'
    ent = helper.infer(gold["text"])
    assert len(helper.feature['softlexicon_ids']) == L * 40
    pred = est.predict(features_to_batch([helper.feature]))['pred_ids'].numpy()[0]
    assert dict(ent) == dict(extract_entity(helper.feature['tokens'], [int(i) for i in pred], idx2tag))


def test_gpu_span_extractor_equals_extract_entity():
    """ner_extract_spans + host string join == reference-shaped extract_entity (tools/infer_utils.py:76-99) on random —
    mostly ill-formed — tag sequences: I after O / [PAD], type switches inside a span, B runs, spans touching both ends."""
    from chinesener_b200.inference import TAG2IDX
    from chinesener_b200.tools.infer_utils import extract_entity, extract_entity_device
    idx2tag = {v: k for k, v in TAG2IDX.items()}
    rng = np.random.default_rng(0)
    for B, L, p_o in ((64, 150, 0.5), (7, 33, 0.1), (3, 1, 0.3), (130, 64, 0.8), (5, 400, 0.0)):
        probs = np.array([0.05, p_o, .1, .15, .1, .15, .1, .15, .02, .02])
        pred = rng.choice(10, size=(B, L), p=probs / probs.sum()).astype(np.int32)
        pred[0, :] = 3                                          # all I-ORG: one span over the whole sentence
        if B > 1:
            pred[1, :] = 2                                      # all B-ORG: L single-token spans
        tokens = [[chr(0x4E00 + int(rng.integers(0, 3000))) for _ in range(L)] for _ in range(B)]
        got = extract_entity_device(tokens, torch.from_numpy(pred).cuda(), idx2tag)
        for b in range(B):
            want = extract_entity(tokens[b], [int(i) for i in pred[b]], idx2tag)
            assert dict(got[b]) == dict(want), (B, L, b)


def test_infer_batch_uses_the_device_span_path(tmp_path):
    from chinesener_b200.data.tokenizer import TokenizerAdapter
    from chinesener_b200.inference import InferHelper, TAG2IDX
    chars = list("中共中央致中国致公党十一大的贺词各位代表同志们")
    params = dict(synthetic.data_params(32), embedding=_char_table(len(chars) + 2))
    est = engine.Estimator("bilstm_crf", params)
    helper = InferHelper(32, TAG2IDX, "bilstm_crf", TokenizerAdapter(chars), estimator=est)
    texts = ["中共中央致中国致公党十一大的贺词", "各位代表、各位同志", "中"]
    helper.infer(texts[0])
    est.store.vars["logits/kernel"].mul_(30.0)
    est.store.touch()
    assert [dict(e) for e in helper.infer_batch(texts)] == [dict(helper.infer(t)) for t in texts]
