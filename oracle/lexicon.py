# -*-coding:utf-8 -*-
"""CPU restatement of the reference's SoftLexicon feature builder — TEST INFRASTRUCTURE (the checker of the C++ trie builder
chinesener_b200/csrc/lexicon_host.cu); nothing under chinesener_b200/ imports it.

  build_soft_lexicon      <- data/word_enhance.py:302-337   (the O(L * 10) substring-in-dict loop, kept as written)
  combine_soft_lexicon    <- data/word_enhance.py:150-160
  align_with_token        <- data/word_enhance.py:89-119
  postproc_soft_lexicon   <- data/word_enhance.py:163-205
  format_soft_seq         <- data/base_preprocess.py:397-412

One deliberate difference: the reference collects matches in Python `set`s, whose iteration order depends on the process's
string hash seed — the order of ids inside a B/M/E/S group (and which of several equally frequent words survives the
top-10 cut) is therefore not defined by the reference.  Here the sets are insertion-ordered (dict keys), i.e. first-seen
order, which is one of the orders the reference can produce and the one the C++ builder emits.
"""
from collections import OrderedDict
from itertools import chain

MaxWordLen = 10
MaxLexiconLen = 10
SoftKeys = ('B', 'M', 'E', 'S')


def build_soft_lexicon(sentence, vocab):
    """vocab: object with vocab2idx / none_token (chinesener_b200.data.word_enhance.WordVocab).
    -> per character OrderedDict{'B': [ids], 'M': [...], 'E': [...], 'S': [...]}; an empty set holds the <None> token."""
    sentence = sentence.replace(' ', '')
    soft_lexicon = [OrderedDict((k, OrderedDict()) for k in SoftKeys) for _ in range(len(sentence))]
    for i in range(len(sentence)):
        for j in range(i, min(i + MaxWordLen, len(sentence))):
            word = sentence[i:(j + 1)]
            if word in vocab.vocab2idx:
                if j - i == 0:
                    soft_lexicon[i]['S'][word] = None
                else:
                    soft_lexicon[i]['B'][word] = None
                    soft_lexicon[j]['E'][word] = None
                    for k in range(i + 1, j):
                        soft_lexicon[k]['M'][word] = None
        for key, val in soft_lexicon[i].items():
            if not val:
                soft_lexicon[i][key][vocab.none_token] = None
    return [OrderedDict((k, [vocab.vocab2idx[w] for w in v]) for k, v in lex.items()) for lex in soft_lexicon]


def combine_soft_lexicon(idx_list):
    """Union, set by set, of the lexicons of the characters one word piece swallowed.  The reference walks the keys of
    Soft2Idx, which include 'None' — a key build_soft_lexicon never creates, so its loop raises KeyError — and would emit
    the sets in S/M/B/E order where unmerged rows are B/M/E/S; this restatement keeps the B/M/E/S row layout."""
    merged = OrderedDict((k, []) for k in SoftKeys)
    for lexicon in idx_list:
        for key in SoftKeys:
            for i in lexicon[key]:
                if i not in merged[key]:
                    merged[key].append(i)
    return merged


def token_char_lens(tokens):
    """characters each token covers (reference :94): '##' stripped, [UNK] = 1, special tokens skipped."""
    return [len(t.replace('##', '')) if t != '[UNK]' else 1 for t in tokens if t not in ('[CLS]', '[SEP]', '[PAD]')]


def align_with_token(idx_list, tokens, combine_func=combine_soft_lexicon):
    token_len = token_char_lens(tokens)
    if len(idx_list) == len(token_len):
        return idx_list
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


def format_soft_seq(seq, max_seq_len, bert, type='ids'):
    default_encoding = [0.0 if type == 'weight' else 0] * (len(SoftKeys) * MaxLexiconLen)
    seq = ([default_encoding] + seq[:(max_seq_len - 2)] + [default_encoding]) if bert else seq[:max_seq_len]
    seq = seq + [default_encoding] * (max_seq_len - len(seq))
    return list(chain(*seq))


def soft_lexicon_features(sentence, vocab, max_seq_len, bert=False, tokens=None, vocabfreq=None):
    """SoftLexiconProc.build_seq_feature's two extra features for one sentence -> (ids [L*40], weights [L*40])."""
    lex = build_soft_lexicon(sentence, vocab)
    if bert and tokens is not None:
        lex = align_with_token(lex, tokens)
    ids, weights = postproc_soft_lexicon(lex, vocab, vocabfreq)
    return format_soft_seq(ids, max_seq_len, bert), format_soft_seq(weights, max_seq_len, bert, type='weight')
