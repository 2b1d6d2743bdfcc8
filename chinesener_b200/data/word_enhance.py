# -*-coding:utf-8 -*-
"""SoftLexicon host side — the step just before the gather-and-pool kernel (ner_softlexicon_pool_fwd).

The matching itself (reference data/word_enhance.py:302-337 build_soft_lexicon, :89-119 align_with_token, :163-205
postproc_soft_lexicon, data/base_preprocess.py:397-412 format_soft_seq) runs in C++ behind the C-ABI
(`ner_lexicon_create / ner_lexicon_build`, chinesener_b200/csrc/lexicon_host.cu): a code-point trie over the word
vocabulary, whole lists of sentences per call on a pool of host threads, output already in the kernel's
[n, max_seq_len * 40] id / weight layout.  This module keeps the reference's Python surface around it: `WordVocab` (the word
side of VocabModel, :36-70) and `SoftLexiconProc` (data/base_preprocess.py:376-438).  The Python restatement of the reference
loop lives in oracle/lexicon.py and is only the tests' checker.
"""
import numpy as np

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
        self.index2word = list(index2word)
        self.vocab2idx = dict((w, i) for i, w in enumerate(self.index2word))
        self.vocab_freq = dict((self.vocab2idx[w], counts[w]) for w in self.index2word)
        self.n_word = len(self.vocab_freq)
        self.vocab2idx.update({self.none_token: self.n_word, self.pad_token: self.n_word + 1, self.eos_token: self.n_word + 2})
        self.vocab_freq.update({self.vocab2idx[self.none_token]: 1, self.vocab2idx[self.pad_token]: 0})


def _utf32(strings):
    """list of str -> (uint32 code points back to back, int64 offsets [n + 1])."""
    offs = np.zeros(len(strings) + 1, np.int64)
    np.cumsum([len(s) for s in strings], out=offs[1:])
    cps = np.frombuffer(''.join(strings).encode('utf-32-le'), dtype=np.uint32)
    assert cps.size == offs[-1]
    return np.ascontiguousarray(cps), offs


def token_char_lens(tokens):
    """characters each WordPiece token covers (reference align_with_token :94): '##' stripped, [UNK] = 1, specials skipped."""
    return [len(t.replace('##', '')) if t != '[UNK]' else 1 for t in tokens if t not in ('[CLS]', '[SEP]', '[PAD]')]


class NativeLexicon(object):
    """The vocabulary as a C++ trie.  `vocabfreq` (id -> frequency, default the vocabulary's own counts) is what
    postproc_soft_lexicon weighs by; ids it does not hold weigh 1 (`vocabfreq.get(i, 1)`, reference :182)."""

    def __init__(self, vocab, vocabfreq=None):
        from .. import _lib
        self._lib = _lib.lib()
        self.vocab, self.n_word = vocab, vocab.n_word
        words = getattr(vocab, 'index2word', None)
        if words is None:
            words = [w for w, i in sorted(vocab.vocab2idx.items(), key=lambda kv: kv[1]) if i < vocab.n_word]
        freq_of = vocab.vocab_freq if vocabfreq is None else vocabfreq
        freq = np.asarray([freq_of.get(i, 1) for i in range(self.n_word + 2)], dtype=np.float64)
        cps, offs = _utf32(words)
        self._h = self._lib.ner_lexicon_create(cps.ctypes.data, offs.ctypes.data, freq.ctypes.data, self.n_word)
        if not self._h:
            raise _lib.NerB200Error('ner_lexicon_create failed (bad vocabulary input)')

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h:
            self._lib.ner_lexicon_destroy(h)

    def num_nodes(self):
        return int(self._lib.ner_lexicon_num_nodes(self._h))

    def build(self, sentences, max_seq_len, bert=False, tokens=None, n_threads=0):
        """sentences: list of raw strings; tokens (bert only): per sentence the tokenizer's tokens, so that rows follow the
        word pieces.  -> (ids int32 [n, max_seq_len * 40], weights float32 [n, max_seq_len * 40])."""
        from .. import _lib
        n = len(sentences)
        cps, offs = _utf32([s.replace(' ', '') for s in sentences])
        tl = toff = None
        if bert and tokens is not None:
            lens = [token_char_lens(t) for t in tokens]
            toff = np.zeros(n + 1, np.int64)
            np.cumsum([len(x) for x in lens], out=toff[1:])
            tl = np.asarray([v for x in lens for v in x], dtype=np.int32)
        ids = np.empty((n, max_seq_len * len(SoftKeys) * MaxLexiconLen), np.int32)
        wts = np.empty(ids.shape, np.float32)
        if tl is not None and tl.size == 0:
            tl = np.zeros(1, np.int32)          # keep the pointer non-NULL: "token lengths given, all sentences empty"
        _lib.check(self._lib.ner_lexicon_build(
            self._h, cps.ctypes.data if cps.size else None, offs.ctypes.data, n, None if tl is None else tl.ctypes.data,
            None if toff is None else toff.ctypes.data, max_seq_len, 1 if bert else 0, ids.ctypes.data, wts.ctypes.data, n_threads))
        _lib.LAUNCHES -= 1          # host call, not a kernel launch
        return ids, wts


class SoftLexiconProc(BasicProc):
    """BasicProc + softlexicon_ids / softlexicon_weights (one lexicon row per token; word pieces merge their characters).
    `build_seq_features(sentences)` featurises a list in one native call; `build_seq_feature` is the reference's
    one-sentence surface over it."""

    def __init__(self, tokenizer_type, max_seq_len, tag2idx, tokenizer, vocab, vocabfreq=None):
        super(SoftLexiconProc, self).__init__(tokenizer_type, max_seq_len, tag2idx, tokenizer)
        self.vocab, self.vocabfreq = vocab, vocabfreq
        self.word_enhance = 'softlexicon'
        self.lexicon = NativeLexicon(vocab, vocabfreq)

    def build_seq_features(self, sentences, n_threads=0):
        feats = [super(SoftLexiconProc, self).build_seq_feature(s) for s in sentences]
        bert = self.tokenizer_type == TokenizerBert
        ids, wts = self.lexicon.build(sentences, self.max_seq_len, bert, [f['tokens'] for f in feats] if bert else None, n_threads)
        for f, i, w in zip(feats, ids, wts):
            f['softlexicon_ids'], f['softlexicon_weights'] = i.tolist(), w.tolist()
        return feats

    def build_seq_feature(self, sentence):
        return self.build_seq_features([sentence], n_threads=1)[0]

    def build_data_params(self, n_sample):
        params = super(SoftLexiconProc, self).build_data_params(n_sample)
        params.update({'word_enhance_dim': len(SoftKeys), 'max_lexicon_len': MaxLexiconLen, 'vocab2idx': self.vocab.vocab2idx})
        return params
