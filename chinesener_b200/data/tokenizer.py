# -*-coding:utf-8 -*-
"""Tokenizers with the reference's surface (reference data/tokenizer.py:14-100).

`get_bert_tokenizer` returns the WordPiece `FullTokenizer` the reference takes from bert-base 0.0.9
(`bert_base.bert.tokenization`, a copy of google-research/bert tokenization.py — third party, not under
/root/reference): its published algorithm is restated here — BasicTokenizer (text cleaning, whitespace around every
CJK codepoint, lower-casing + accent stripping, punctuation splitting) followed by greedy longest-match-first
WordPiece with the `##` continuation prefix.  `TokenizerAdapter` is the reference's own character tokenizer for the
giga embedding (full-width -> half-width folding, [UNK] for out-of-vocabulary characters, [PAD]/[UNK] appended to the
vocabulary; data/tokenizer.py:47-100).  Pinned by the featurised warm-up sentence of the reference's serving models
(tests/golden/warmup_features.json, tests/test_host_pipeline.py).
"""
import collections
import os
import unicodedata

TokenizerBert = 'bert'
TokenizerGiga = 'giga'


def load_vocab(vocab_file):
    vocab = collections.OrderedDict()
    with open(vocab_file, encoding='utf-8') as f:
        for index, line in enumerate(f):
            vocab[line.strip()] = index
    return vocab


def _is_whitespace(ch):
    if ch in (' ', '\t', '\n', '\r'):
        return True
    return unicodedata.category(ch) == 'Zs'


def _is_control(ch):
    if ch in ('\t', '\n', '\r'):
        return False
    return unicodedata.category(ch) in ('Cc', 'Cf')


def _is_punctuation(ch):
    cp = ord(ch)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
        return True
    return unicodedata.category(ch).startswith('P')


def _is_chinese_char(cp):
    return (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2A6DF or 0x2A700 <= cp <= 0x2B73F
            or 0x2B740 <= cp <= 0x2B81F or 0x2B820 <= cp <= 0x2CEAF or 0xF900 <= cp <= 0xFAFF or 0x2F800 <= cp <= 0x2FA1F)


class BasicTokenizer(object):
    def __init__(self, do_lower_case=True):
        self.do_lower_case = do_lower_case

    def tokenize(self, text):
        out = []
        for ch in text:                                   # _clean_text + _tokenize_chinese_chars
            cp = ord(ch)
            if cp == 0 or cp == 0xFFFD or _is_control(ch):
                continue
            if _is_whitespace(ch):
                out.append(' ')
            elif _is_chinese_char(cp):
                out.extend((' ', ch, ' '))
            else:
                out.append(ch)
        split_tokens = []
        for token in ''.join(out).strip().split():
            if self.do_lower_case:
                token = ''.join(c for c in unicodedata.normalize('NFD', token.lower()) if unicodedata.category(c) != 'Mn')
            word = []                                     # _run_split_on_punc
            for ch in token:
                if _is_punctuation(ch):
                    if word:
                        split_tokens.append(''.join(word))
                        word = []
                    split_tokens.append(ch)
                else:
                    word.append(ch)
            if word:
                split_tokens.append(''.join(word))
        return split_tokens


class WordpieceTokenizer(object):
    def __init__(self, vocab, unk_token='[UNK]', max_input_chars_per_word=200):
        self.vocab, self.unk_token, self.max_input_chars_per_word = vocab, unk_token, max_input_chars_per_word

    def tokenize(self, token):
        if len(token) > self.max_input_chars_per_word:
            return [self.unk_token]
        pieces, start = [], 0
        while start < len(token):
            end, cur = len(token), None
            while start < end:                            # longest match first
                sub = token[start:end] if start == 0 else '##' + token[start:end]
                if sub in self.vocab:
                    cur = sub
                    break
                end -= 1
            if cur is None:
                return [self.unk_token]
            pieces.append(cur)
            start = end
        return pieces


class FullTokenizer(object):
    """bert tokenization.FullTokenizer(vocab_file, do_lower_case): tokenize / convert_tokens_to_ids / convert_ids_to_tokens.
    `vocab_file` may also be a ready {token: id} mapping."""

    def __init__(self, vocab_file, do_lower_case=True):
        self.vocab = vocab_file if isinstance(vocab_file, dict) else load_vocab(vocab_file)
        self.inv_vocab = {v: k for k, v in self.vocab.items()}
        self.basic_tokenizer = BasicTokenizer(do_lower_case=do_lower_case)
        self.wordpiece_tokenizer = WordpieceTokenizer(vocab=self.vocab)

    def tokenize(self, text):
        return [p for t in self.basic_tokenizer.tokenize(text) for p in self.wordpiece_tokenizer.tokenize(t)]

    def convert_tokens_to_ids(self, tokens):
        return [self.vocab[t] for t in tokens]

    def convert_ids_to_tokens(self, ids):
        return [self.inv_vocab[i] for i in ids]


def get_bert_tokenizer(model_dir='./pretrain_model/ch_google/'):
    """reference data/tokenizer.py:14-21 (either google_bert or wwm_bert: same vocabulary)."""
    return FullTokenizer(os.path.join(model_dir, 'vocab.txt'), do_lower_case=True)


class TextVectors(object):
    """The three attributes of gensim's KeyedVectors the reference reads (`index2word`, `vectors`, `vector_size`),
    loaded from a GloVe-style text file — one `word v_1 ... v_d` line per entry, no header — which is what
    pretrain_model/glove_2_wv.py:11-23 converts and loads (pretrain_model/giga/__init__.py).  A repeated word keeps its
    first vector, as load_word2vec_format does."""

    def __init__(self, path):
        import numpy as np
        words, rows, seen = [], [], set()
        with open(path, encoding='utf-8') as f:
            for line in f:
                parts = line.rstrip().split(' ')
                if len(parts) < 3 or parts[0] in seen:
                    continue
                seen.add(parts[0])
                words.append(parts[0])
                rows.append(np.asarray(parts[1:], dtype=np.float32))
        self.index2word, self.vectors = words, np.stack(rows)
        self.vector_size = self.vectors.shape[1]


def get_giga_tokenizer(vec_file='./pretrain_model/giga/gigaword_chn.all.a2b.uni.ite50.vec'):
    """reference data/tokenizer.py:24-33 — the character tokenizer of every non-BERT model."""
    return TokenizerAdapter(TextVectors(vec_file))


class TokenizerAdapter(object):
    """reference data/tokenizer.py:47-100 — character tokenizer over a word2vec-style model (`index2word`, optional
    `vectors`); here `model` may also be a plain list of vocabulary entries."""

    def __init__(self, model):
        self.model = model
        self.vocab2idx = self.get_vocab2idx()

    def get_vocab2idx(self):
        words = self.model if isinstance(self.model, (list, tuple)) else self.model.index2word
        vocab2idx = dict((word, idx) for idx, word in enumerate(words))
        n_vocab = len(vocab2idx)
        vocab2idx.update({'[PAD]': n_vocab, '[UNK]': n_vocab + 1})
        return vocab2idx

    def embedding(self, seed=None):
        """reference :68-74 (a property there): pretrained vectors + two N(0,1) rows for [PAD] / [UNK], every row scaled
        to unit L2 norm (tools/utils.py:10-14).  The reference draws the add-on rows from numpy's unseeded global
        stream; `seed` makes the draw repeatable."""
        import numpy as np
        rng = np.random.default_rng(seed) if seed is not None else np.random
        emb = np.vstack((np.asarray(self.model.vectors), rng.normal(0, 1, size=(2, self.model.vector_size)))).astype(np.float32)
        norm = np.linalg.norm(emb, axis=1, keepdims=True)
        norm[norm == 0] = np.finfo(np.float32).eps
        return emb / norm

    @staticmethod
    def full2half(text):
        num = ord(text)
        if num == 0x3000:
            num = 0x20
        elif 0xFF01 <= num <= 0xFF5E:
            num = num - 0xFEE0
        return chr(num)

    def tokenize(self, text):
        tokens = []
        for i in text:
            if i.strip():
                i = self.full2half(i)
                tokens.append(i if i in self.vocab2idx else '[UNK]')
        return tokens

    def convert_tokens_to_ids(self, tokens):
        return [self.vocab2idx[i] for i in tokens]
