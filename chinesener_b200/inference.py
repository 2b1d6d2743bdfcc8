# -*-coding:utf-8 -*-
"""In-process counterpart of reference inference.py:33-99: `InferHelper.infer(text)` featurises one sentence exactly
like the reference's serving client and runs PREDICT on the local engine instead of a TF-Serving gRPC round trip."""
import re

import numpy as np

from .data.base_preprocess import BasicProc, features_to_batch
from .data.tokenizer import TokenizerBert, TokenizerGiga
from .tools.infer_utils import extract_entity, fix_tokens

MAX_SEQ_LEN = 150
TAG2IDX = {'[PAD]': 0, 'O': 1, 'B-ORG': 2, 'I-ORG': 3, 'B-PER': 4, 'I-PER': 5, 'B-LOC': 6, 'I-LOC': 7, '[CLS]': 8, '[SEP]': 9}


class InferHelper(object):
    def __init__(self, max_seq_len, tag2idx, model_name, tokenizer, estimator=None):
        self.model_name = model_name
        self.tokenizer_type = TokenizerBert if re.search('bert', model_name) else TokenizerGiga
        self.mtl = 1 if re.search('(mtl)|(adv)', model_name) else 0
        self.proc = BasicProc(self.tokenizer_type, max_seq_len, tag2idx, tokenizer)
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
