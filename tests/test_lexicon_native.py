"""CPU: the C++ SoftLexicon builder (ner_lexicon_*, host code inside libner_b200.so) against the Python restatement of the
reference loop (oracle/lexicon.py) — same ids, same float32 weights, on random text over random vocabularies, with
WordPiece merging, truncation to the ten most frequent words, [CLS]/[SEP] rows, empty and over-long sentences."""
import random
import time

import numpy as np
import pytest

from chinesener_b200 import _lib
from chinesener_b200.data.word_enhance import NativeLexicon, SoftLexiconProc, WordVocab
from oracle import lexicon as ol

ALPHABET = "的一是在不了有和人这中大为上个国我以要他时来用们生到作地于出就分对成会可主发年动同工也能下过子说产种面而方后多定行学法所民得经十三"


def _random_vocab(rng, n_words, alphabet=ALPHABET, max_len=6, distinct_freq=True):
    words = set(alphabet[: len(alphabet) // 2])                       # half of the single characters are words themselves
    while len(words) < n_words:
        words.add(''.join(rng.choice(alphabet) for _ in range(rng.randint(2, max_len))))
    words = sorted(words)
    rng.shuffle(words)
    freqs = rng.sample(range(1, 50 * len(words)), len(words)) if distinct_freq else [rng.randint(1, 5) for _ in words]
    return WordVocab(words, dict(zip(words, freqs)))


def _oracle(sentences, vocab, L, bert=False, tokens=None, vocabfreq=None):
    ids, wts = [], []
    for k, s in enumerate(sentences):
        i, w = ol.soft_lexicon_features(s, vocab, L, bert, None if tokens is None else tokens[k], vocabfreq)
        ids.append(i)
        wts.append(w)
    return np.asarray(ids, np.int32), np.asarray(wts, np.float64).astype(np.float32)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_native_builder_equals_the_reference_loop_on_random_text(seed):
    rng = random.Random(seed)
    vocab = _random_vocab(rng, 400)
    lex = NativeLexicon(vocab)
    L = 40
    sents = [''.join(rng.choice(ALPHABET + " ") for _ in range(rng.choice([0, 1, 2, 7, 25, 39, 40, 41, 70]))) for _ in range(120)]
    ids, wts = lex.build(sents, L, n_threads=3)
    rid, rw = _oracle(sents, vocab, L)
    np.testing.assert_array_equal(ids, rid)
    np.testing.assert_array_equal(wts, rw)                         # same float64 division, same float32 rounding
    assert ids.dtype == np.int32 and wts.dtype == np.float32 and ids.shape == (120, L * 40)
    real = wts.reshape(120, L, 40).sum(-1)
    n_chars = np.minimum([len(s.replace(' ', '')) for s in sents], L)
    for r, n in zip(real, n_chars):
        np.testing.assert_allclose(r[:n], 1.0, atol=1e-6)          # every character row is normalised
        assert (r[n:] == 0).all()                                  # padding rows: all-zero ids and weights
    one, _ = lex.build(sents[:1], L, n_threads=1)
    np.testing.assert_array_equal(one[0], ids[0])                  # thread count does not change results


def test_truncation_keeps_the_ten_most_frequent_including_ties():
    rng = random.Random(5)
    # a vocabulary dense enough that the M sets overflow 10 entries: every 1..7-gram over two letters is a word
    import itertools
    letters = "中国"
    words = [''.join(p) for n in range(1, 8) for p in itertools.product(letters, repeat=n)]
    words = sorted(set(words))
    for distinct in (True, False):                                 # ties: the stable sort keeps first-seen order, as Python's does
        freqs = rng.sample(range(1, 10000), len(words)) if distinct else [rng.randint(1, 3) for _ in words]
        vocab = WordVocab(words, dict(zip(words, freqs)))
        lex = NativeLexicon(vocab)
        sents = [''.join(rng.choice(letters) for _ in range(30)) for _ in range(20)]
        ids, wts = lex.build(sents, 32)
        assert (ids.reshape(20, 32, 4, 10)[:, :30] != vocab.vocab2idx['<PAD>']).all(-1).any()      # some set is full
        rid, rw = _oracle(sents, vocab, 32)
        np.testing.assert_array_equal(ids, rid)
        np.testing.assert_array_equal(wts, rw)


def test_word_piece_rows_and_bert_layout():
    from chinesener_b200.data.tokenizer import FullTokenizer, TokenizerBert
    from chinesener_b200.inference import TAG2IDX
    tok = FullTokenizer({t: i for i, t in enumerate(['[PAD]', '[UNK]', '[CLS]', '[SEP]', '19', '##94', '年', '夏', '20', '##08', '##年'])})
    vocab = WordVocab(['1994', '19', '94年', '年', '夏', '2008', '08年夏'], {'1994': 5, '19': 4, '94年': 3, '年': 2, '夏': 1, '2008': 9, '08年夏': 7})
    sents = ['1994年夏', '2008年 夏', '夏', '']
    tokens = [['[CLS]'] + tok.tokenize(s) + ['[SEP]'] for s in sents]
    lex = NativeLexicon(vocab)
    for L in (8, 5):                                               # 5: more tokens than max_seq_len - 2 rows
        ids, wts = lex.build(sents, L, bert=True, tokens=tokens)
        rid, rw = _oracle(sents, vocab, L, bert=True, tokens=tokens)
        np.testing.assert_array_equal(ids, rid)
        np.testing.assert_array_equal(wts, rw)
        assert (ids[:, :40] == 0).all() and (wts[:, :40] == 0).all()            # [CLS] row
    proc = SoftLexiconProc(TokenizerBert, 8, TAG2IDX, tok, vocab)
    f = proc.build_seq_feature(sents[0])
    assert f['softlexicon_ids'] == rid_row(sents[0], vocab, tokens[0], 8)
    many = proc.build_seq_features(sents)
    assert [m['softlexicon_ids'] for m in many] == [proc.build_seq_feature(s)['softlexicon_ids'] for s in sents]


def rid_row(sentence, vocab, tokens, L):
    return ol.soft_lexicon_features(sentence, vocab, L, True, tokens)[0]


def test_custom_vocabfreq_and_missing_ids_weigh_one():
    rng = random.Random(9)
    vocab = _random_vocab(rng, 100)
    custom = {i: rng.randint(1, 50) for i in range(0, vocab.n_word, 2)}          # odd ids missing -> frequency 1; <PAD> missing -> 1 (!)
    lex = NativeLexicon(vocab, vocabfreq=custom)
    sents = [''.join(rng.choice(ALPHABET) for _ in range(20)) for _ in range(10)]
    ids, wts = lex.build(sents, 24)
    rid, rw = _oracle(sents, vocab, 24, vocabfreq=custom)
    np.testing.assert_array_equal(ids, rid)
    np.testing.assert_array_equal(wts, rw)


def test_bad_input_is_rejected():
    h = _lib.lib()
    assert not h.ner_lexicon_create(None, None, None, 3)
    assert h.ner_lexicon_build(None, None, None, 1, None, None, 8, 0, None, None, 1) == -1
    assert h.ner_lexicon_num_nodes(None) == -1
    vocab = WordVocab(['中', '中国'], {'中': 1, '国': 1, '中国': 2})
    lex = NativeLexicon(vocab)
    assert lex.num_nodes() == 3                                     # root, 中, 中国
    ids, wts = lex.build([], 8)
    assert ids.shape == (0, 320)


def test_throughput_against_the_python_loop():
    """Not a pass/fail benchmark: asserts only that the native builder is not slower, prints both rates (pytest -s)."""
    rng = random.Random(3)
    vocab = _random_vocab(rng, 20000, max_len=4)
    sents = [''.join(rng.choice(ALPHABET) for _ in range(rng.randint(20, 120))) for _ in range(400)]
    lex = NativeLexicon(vocab)
    lex.build(sents, 150, n_threads=1)                 # first call pays the page faults of the fresh output arrays
    t0 = time.perf_counter()
    lex.build(sents, 150, n_threads=1)
    t_native = time.perf_counter() - t0
    t0 = time.perf_counter()
    _oracle(sents[:100], vocab, 150)
    t_python = (time.perf_counter() - t0) * 4
    print(f"SoftLexicon builder: native {len(sents) / t_native:.0f} sentences/s (1 thread), python loop {len(sents) / t_python:.0f} sentences/s")
    assert t_native < t_python
