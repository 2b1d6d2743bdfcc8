# -*-coding:utf-8 -*-
"""In-process counterpart of reference inference.py:33-99: `InferHelper.infer(text)` featurises one sentence exactly
like the reference's serving client and runs PREDICT on the local engine instead of a TF-Serving gRPC round trip."""
import re

import numpy as np

from .data.base_preprocess import extract_prefix_surfix, features_to_batch, get_instance
from .data.tokenizer import TokenizerBert
from .tools.infer_utils import extract_entity, fix_tokens

MAX_SEQ_LEN = 150
TAG2IDX = {'[PAD]': 0, 'O': 1, 'B-ORG': 2, 'I-ORG': 3, 'B-PER': 4, 'I-PER': 5, 'B-LOC': 6, 'I-LOC': 7, '[CLS]': 8, '[SEP]': 9}


class InferHelper(object):
    """`tokenizer` (and, for the word-enhance models, the keyword arguments their processor needs — the SoftLexicon word
    vocabulary) are passed in where the reference looks them up by module name; the processor class is picked from the
    model name exactly as the reference does (extract_prefix_surfix + get_instance, inference.py:36-39)."""

    def __init__(self, max_seq_len, tag2idx, model_name, tokenizer, estimator=None, **proc_kwargs):
        self.model_name = model_name
        self.word_enhance, self.tokenizer_type = extract_prefix_surfix(model_name)
        self.mtl = 1 if re.search('(mtl)|(adv)', model_name) else 0            # whether is multitask
        self.proc = get_instance(self.tokenizer_type, max_seq_len, tag2idx, tokenizer, word_enhance=self.word_enhance, **proc_kwargs)
        self.max_seq_len, self.tag2idx = max_seq_len, tag2idx
        self.idx2tag = dict((v, k) for k, v in tag2idx.items())
        self.estimator = estimator
        self.feature = None

    def make_feature(self, sentence):
        """reference inference.py:64-82: sequence features + fake labels ('0.0' strings / zero ids), task id for the
        multi-task models, WordPiece tokens mapped back to the sentence's characters."""
        self.feature = self.proc.build_seq_feature(sentence)
        self.feature['labels'] = np.zeros(shape=(self.max_seq_len,)).astype(str).tolist()
        self.feature['label_ids'] = np.zeros(shape=(self.max_seq_len,)).astype(int).tolist()
        if self.mtl:
            self.feature['task_ids'] = 1
        if self.tokenizer_type == TokenizerBert:
            self.feature['tokens'] = fix_tokens(sentence, self.feature['tokens'])
        return self.feature

    def decode_prediction(self, pred_ids):
        return extract_entity(self.feature['tokens'], [int(i) for i in np.squeeze(pred_ids)], self.idx2tag)

    def infer(self, text):
        feature = self.make_feature(text)
        out = self.estimator.predict(features_to_batch([feature]))
        return self.decode_prediction(out['pred_ids'].numpy())

    def infer_batch(self, texts):
        """Many sentences per call: one PREDICT batch, the tag scan of extract_entity on the GPU (ner_extract_spans) —
        the tag tensor stays on the device, only the spans come back.  -> one entity dict per text, as infer() gives."""
        from .tools.infer_utils import extract_entity_device
        feats = [dict(self.make_feature(t)) for t in texts]
        dev = self.estimator.to_device(features_to_batch(feats, pin_memory=True))
        pred = self.estimator.predict_device(dev)
        return extract_entity_device([f['tokens'] for f in feats], pred, self.idx2tag)
