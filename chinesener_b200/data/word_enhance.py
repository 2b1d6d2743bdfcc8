# -*-coding:utf-8 -*-
"""SoftLexicon host builder — the step just before the gather-and-pool kernel
(ner_softlexicon_pool_fwd).  Restates reference data/word_enhance.py:36-70 (VocabModel add-on tokens and frequencies),
:302-337 (build_soft_lexicon: every substring of <= 10 characters found in the word vocabulary lands in the B/M/E/S
sets of the characters it covers), :163-205 (postproc_soft_lexicon: pad to / keep the 10 most frequent per set, weights
= frequency normalised over all four sets of the token) and data/base_preprocess.py:376-438 (SoftLexiconProc: CLS/SEP/PAD
rows are all-zero ids and weights, everything flattened to [max_seq_len * 4 * 10])."""
from collections import OrderedDict
from itertools import chain

from .base_preprocess import BasicProc
from .tokenizer import TokenizerBert

MaxWordLen = 10
MaxLexiconLen = 10      # only keep top-n words per B/M/E/S set
SoftKeys = ('B', 'M', 'E', 'S')


class WordVocab(object):
    """The word side of reference VocabModel (:36-70): vocabulary order = embedding row, add-on tokens <None> = n_word,
    <PAD> = n_word + 1, <eos> = n_word + 2; frequencies from the pretrained model, <None> -> 1, <PAD> -> 0."""
    none_token, pad_token, eos_token = '<None>', '<PAD>', '<eos>'

    def __init__(self, index2word, counts):
        self.vocab2idx = dict((w, i) for i, w in enumerate(index2word))
        self.vocab_freq = dict((self.vocab2idx[w], counts[w]) for w in index2word)
        self.n_word = len(self.vocab_freq)
        self.vocab2idx.update({self.none_token: self.n_word, self.pad_token: self.n_word + 1, self.eos_token: self.n_word + 2})
        self.vocab_freq.update({self.vocab2idx[self.none_token]: 1, self.vocab2idx[self.pad_token]: 0})


def build_soft_lexicon(sentence, vocab):
    """-> per character {'B': [ids], 'M': [...], 'E': [...], 'S': [...]}; an empty set holds the <None> token."""
    sentence = sentence.replace(' ', '')
    soft_lexicon = [OrderedDict((k, set()) for k in SoftKeys) for _ in range(len(sentence))]
    for i in range(len(sentence)):
        for j in range(i, min(i + MaxWordLen, len(sentence))):
            word = sentence[i:(j + 1)]
            if word in vocab.vocab2idx:
                if j - i == 0:
                    soft_lexicon[i]['S'].add(word)
                else:
                    soft_lexicon[i]['B'].add(word)
                    soft_lexicon[j]['E'].add(word)
                    for k in range(i + 1, j):
                        soft_lexicon[k]['M'].add(word)
        for key, val in soft_lexicon[i].items():
            if not val:
                soft_lexicon[i][key].add(vocab.none_token)
    return [OrderedDict((k, [vocab.vocab2idx[w] for w in v]) for k, v in lex.items()) for lex in soft_lexicon]


def combine_soft_lexicon(idx_list):
    """Union, set by set, of the lexicons of the characters one word piece swallowed (reference data/word_enhance.py:
    150-160).  The reference walks the keys of Soft2Idx, which include 'None' — a key build_soft_lexicon never creates, so
    its loop raises KeyError — and would emit the sets in S/M/B/E order where unmerged rows are B/M/E/S; this restatement
    keeps the B/M/E/S row layout the pooling kernel is fed everywhere else."""
    merged = OrderedDict((k, []) for k in SoftKeys)
    for lexicon in idx_list:
        for key in SoftKeys:
            for i in lexicon[key]:
                if i not in merged[key]:
                    merged[key].append(i)
    return merged


def align_with_token(idx_list, tokens, combine_func=combine_soft_lexicon):
    """Reference data/word_enhance.py:89-119 — the BERT tokenizer can put several characters into one word piece
    ('1994' -> '19', '##94'); per-character features are merged so there is one row per token.  combine_func: the reference
    picks it by word-enhance method (soft lexicon: union; softword: max; bichar: min)."""
    token_len = [len(t.replace('##', '')) if t != '[UNK]' else 1 for t in tokens if t not in ('[CLS]', '[SEP]', '[PAD]')]
    if len(idx_list) == len(token_len):
        return idx_list                       # no mismatch between the word pieces and the characters
    pos, output_list = 0, []
    for tl in token_len:
        output_list.append(idx_list[pos] if tl == 1 else combine_func(idx_list[pos:pos + tl]))
        pos += tl
    assert len(output_list) == len(token_len)
    return output_list


def postproc_soft_lexicon(output_list, vocab, vocabfreq=None):
    """-> (ids, weights), each seq_len x (4 * MaxLexiconLen)."""
    vocabfreq = vocab.vocab_freq if vocabfreq is None else vocabfreq
    pad_id = vocab.vocab2idx[vocab.pad_token]

    def helper(ids):
        n = len(ids)
        if n <= MaxLexiconLen:
            ids = list(ids) + [pad_id] * (MaxLexiconLen - n)
            return ids, [vocabfreq.get(i, 1) for i in ids]
        tmp = sorted([(i, vocabfreq.get(i, 1)) for i in ids], key=lambda x: x[1], reverse=True)[:MaxLexiconLen]
        return [t[0] for t in tmp], [t[1] for t in tmp]

    seq_ids, seq_weights = [], []
    for lexicon in output_list:
        ids, weights, total_weight = [], [], 0
        for key in SoftKeys:
            i, w = helper(lexicon[key])
            ids += i
            weights += w
            total_weight += sum(w)
        seq_ids.append(ids)
        seq_weights.append([w / total_weight for w in weights])
    return seq_ids, seq_weights


class SoftLexiconProc(BasicProc):
    """BasicProc + softlexicon_ids / softlexicon_weights (one lexicon row per token; word pieces merge their characters)."""

    def __init__(self, tokenizer_type, max_seq_len, tag2idx, tokenizer, vocab, vocabfreq=None):
        super(SoftLexiconProc, self).__init__(tokenizer_type, max_seq_len, tag2idx, tokenizer)
        self.vocab, self.vocabfreq = vocab, vocabfreq

    def format_soft_seq(self, seq, type='ids'):
        default_encoding = [0.0 if type == 'weight' else 0] * (len(SoftKeys) * MaxLexiconLen)
        if self.tokenizer_type == TokenizerBert:
            seq = [default_encoding] + seq[:(self.max_seq_len - 2)] + [default_encoding]
        else:
            seq = seq[:self.max_seq_len]
        seq = seq + [default_encoding] * (self.max_seq_len - len(seq))
        return list(chain(*seq))

    def build_seq_feature(self, sentence):
        f_seq = super(SoftLexiconProc, self).build_seq_feature(sentence)
        soft_lexicon = build_soft_lexicon(sentence, self.vocab)
        if self.tokenizer_type == TokenizerBert:
            soft_lexicon = align_with_token(soft_lexicon, f_seq['tokens'])
        ids, weights = postproc_soft_lexicon(soft_lexicon, self.vocab, self.vocabfreq)
        f_seq['softlexicon_ids'] = self.format_soft_seq(ids)
        f_seq['softlexicon_weights'] = self.format_soft_seq(weights, type='weight')
        return f_seq

    def build_data_params(self, n_sample):
        params = super(SoftLexiconProc, self).build_data_params(n_sample)
        params.update({'word_enhance_dim': len(SoftKeys), 'max_lexicon_len': MaxLexiconLen, 'vocab2idx': self.vocab.vocab2idx})
        return params
