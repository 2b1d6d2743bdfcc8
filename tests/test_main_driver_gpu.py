"""GPU: the command-line driver (chinesener_b200/main.py; reference main.py:14-140) end to end on tiny record sets and a
2-layer BERT: single-dataset run of the hot-path plugin, and the two-dataset run of the multi-task plugin (`--data a,b`)."""
import json
import os
import pickle

import numpy as np
import pytest

from chinesener_b200 import main as driver
from chinesener_b200.data.tokenizer import TokenizerBert

from test_dataset_pipeline import _prepare_two_tasks

pytestmark = pytest.mark.gpu

L = 64


def _setup(tmp_path):
    root, tok = _prepare_two_tasks(tmp_path, TokenizerBert, L)
    cfg = {'vocab_size': len(tok.vocab2idx), 'hidden_size': 768, 'num_hidden_layers': 2, 'num_attention_heads': 12,
           'intermediate_size': 3072, 'max_position_embeddings': 512, 'type_vocab_size': 2, 'initializer_range': 0.02,
           'hidden_dropout_prob': 0.1, 'attention_probs_dropout_prob': 0.1}
    pre = tmp_path / "pretrain"
    pre.mkdir()
    (pre / "bert_config.json").write_text(json.dumps(cfg))
    return root, str(pre)


def test_single_dataset_run_writes_checkpoint_and_prediction_pickle(tmp_path):
    root, pre = _setup(tmp_path)
    report = tmp_path / "rep.json"
    with pytest.warns(UserWarning):                    # no BERT checkpoint in pretrain_dir: random init, said loudly
        s = driver.main(['--model_name', 'bert_bilstm_crf', '--data', 'msra', '--data_dir', os.path.join(root, 'msra'),
                         '--checkpoint_root', str(tmp_path / 'ckpt'), '--pretrain_dir', pre, '--epoch_size', '2', '--batch_size', '4',
                         '--report', str(report)])
    assert s['n_predict'] == 24 and s['history']['final_step'] == 16 * 2 // 4
    pred = pickle.load(open(os.path.join(root, 'msra', 'bert_bilstm_crf_predict.pkl'), 'rb'))
    assert len(pred) == 24 and pred[0]['pred_ids'].shape == (L,) and pred[0]['pred_ids'].dtype == np.int32
    assert pred[0]['tokens'][0] == b'[CLS]'
    assert os.path.isdir(tmp_path / 'ckpt' / 'ner_msra_bert_bilstm_crf') and json.load(open(report))['model'] == 'bert_bilstm_crf'


def test_two_dataset_run_of_the_multitask_plugin(tmp_path):
    root, pre = _setup(tmp_path)
    with pytest.warns(UserWarning):
        s = driver.main(['--model_name', 'bert_bilstm_crf_mtl', '--data', 'msra,msr', '--data_dir', root,
                         '--checkpoint_root', str(tmp_path / 'ckpt'), '--pretrain_dir', pre, '--epoch_size', '2', '--batch_size', '4'])
    assert s['history']['final_step'] == 26 * 2 // 4 and np.isfinite(s['history']['evals'][-1]['loss'])
    assert os.path.isdir(tmp_path / 'ckpt' / 'ner_msra_msr_bert_bilstm_crf_mtl')
    n = {'msra': 24, 'msr': 5}
    for data in ('msra', 'msr'):
        pred = pickle.load(open(os.path.join(root, data, 'bert_bilstm_crf_mtl_msra_msr_predict.pkl'), 'rb'))
        assert len(pred) == n[data] == s['tasks'][data]['n_predict']
        size = 10 if data == 'msra' else 7
        assert all(int(p['pred_ids'].max()) < size for p in pred)            # each dataset decoded by its own tower
    # a second call warm-starts from the checkpoint and only predicts
    s2 = driver.main(['--model_name', 'bert_bilstm_crf_mtl', '--data', 'msra,msr', '--data_dir', root,
                      '--checkpoint_root', str(tmp_path / 'ckpt'), '--pretrain_dir', pre, '--predict_only', '1', '--batch_size', '4'])
    assert s2['history'] is None and s2['tasks']['msr']['n_predict'] == 5
