# -*-coding:utf-8 -*-
"""Dataset preparation: raw `<split>/sentences.txt` + `<split>/tags.txt` -> `.nerrec` column files + `data_params.pkl`.

Counterpart of the reference's per-dataset scripts (data/msra/preprocess.py:1-52) and of BasicProc.init_data /
dump_tfrecord (data/base_preprocess.py:152-156, 229-253): same split renaming (train/val/test -> train/valid/predict),
same file naming `<tokenizer>_<split>[_<word_enhance>]`, same `data_params.pkl` contents, a sentence whose token and tag
counts differ is dropped and counted — but the output is the memory-mappable format of data/records.py instead of
TFRecords.

    python -m chinesener_b200.data.preprocess --src <dir with train/ val/ test/> --out datasets/msra \
        --tokenizer giga --giga_vec <gigaword .vec>  [--bert_vocab <vocab.txt>]
"""
import argparse
import os
import pickle

from .base_preprocess import get_instance
from .records import write_records
from .tokenizer import TokenizerBert, TokenizerGiga, get_bert_tokenizer, get_giga_tokenizer

# data/msra/preprocess.py:7-27 (people_daily uses the same tag set and length)
MSRA_TAG2IDX = {'[PAD]': 0, 'O': 1, 'B-ORG': 2, 'I-ORG': 3, 'B-PER': 4, 'I-PER': 5, 'B-LOC': 6, 'I-LOC': 7, '[CLS]': 8, '[SEP]': 9}
MSRA_MAX_SEQ_LEN = 150
MAPPING = {'train': 'train', 'val': 'valid', 'test': 'predict'}


# data/msr/preprocess.py:7-22 (Chinese word segmentation as a tagging task: the auxiliary task of the multi-task plugins)
MSR_TAG2IDX = {'[PAD]': 0, 'B': 1, 'I': 2, 'E': 3, 'S': 4, '[CLS]': 5, '[SEP]': 6}
MSR_MAPPING = {'training': 'train', 'test_gold': 'valid', 'test': 'predict'}


def msr_gen_tag(length):
    """data/msr/preprocess.py:27-33"""
    if length == 1:
        return 'S'
    if length == 2:
        return 'B E'
    return ' '.join(['B'] + ['I'] * (length - 2) + ['E'])


def load_msr_data(data_dir, file_name):
    """data/msr/preprocess.py:36-52 — `msr_<split>.utf8`: words separated by spaces -> characters + B/I/E/S tags."""
    sentences, tags = [], []
    for line in read_text(data_dir, 'msr_{}.utf8'.format(file_name)):
        if line == '':
            continue
        words = [t for t in line.split(' ') if t not in ['', '"']]
        tags.append(' '.join(msr_gen_tag(len(t)) for t in words))
        sentences.append(' '.join(c for t in words for c in t))
    return sentences, tags


def read_text(data_dir, filename):
    with open(os.path.join(data_dir, filename), 'r', encoding='utf-8') as f:
        return [line.strip() for line in f]


def load_data(data_dir, file_name):
    sentences = read_text(data_dir, os.path.join(file_name, 'sentences.txt'))
    tags = read_text(data_dir, os.path.join(file_name, 'tags.txt'))
    assert len(sentences) == len(tags)
    return sentences, tags


def dump_records(proc, src_dir, out_dir, file_name, mapping=MAPPING, word_enhance=None, embedding=None, verbose=True,
                 load_data=None):
    """One split through `proc.build_feature` -> `<out_dir>/<tokenizer>_<renamed>[_<enhance>].nerrec`; the train split
    also writes `<tokenizer>[_<enhance>]_data_params.pkl` (data/base_preprocess.py:206-227, 247-253)."""
    sentences, tags = (load_data or globals()['load_data'])(src_dir, file_name)
    feats, n_invalid = [], 0
    for sentence, tag in zip(sentences, tags):
        try:
            feats.append(proc.build_feature(sentence, tag))
        except Exception as e:          # the reference prints and skips (n_token != n_tag after tokenisation, unknown tag)
            n_invalid += 1
            if verbose:
                print(e)
    os.makedirs(out_dir, exist_ok=True)
    stem = '_'.join(filter(None, [proc.tokenizer_type, mapping[file_name], word_enhance]))
    write_records(os.path.join(out_dir, stem + '.nerrec'), feats, proc.max_seq_len)
    if verbose:
        print('Dump {} sample, invalid_sample = {}'.format(len(feats), n_invalid))
    if 'train' in file_name:
        params = proc.build_data_params(len(feats))
        if proc.tokenizer_type == TokenizerGiga and embedding is not None:
            params['embedding'] = embedding
        with open(os.path.join(out_dir, '_'.join(filter(None, [proc.tokenizer_type, word_enhance, 'data_params.pkl']))), 'wb') as f:
            pickle.dump(params, f)
    return len(feats), n_invalid


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--src', required=True, help='directory holding train/ val/ test/ with sentences.txt + tags.txt')
    ap.add_argument('--out', required=True)
    ap.add_argument('--tokenizer', default=TokenizerGiga, choices=[TokenizerGiga, TokenizerBert])
    ap.add_argument('--giga_vec', default='./pretrain_model/giga/gigaword_chn.all.a2b.uni.ite50.vec')
    ap.add_argument('--bert_dir', default='./pretrain_model/ch_google/')
    ap.add_argument('--max_seq_len', type=int, default=MSRA_MAX_SEQ_LEN)
    ap.add_argument('--seed', type=int, default=1234, help='seed of the two add-on embedding rows ([PAD], [UNK])')
    ap.add_argument('--format', default='ner', choices=['ner', 'msr'], help="'msr': word-segmented msr_<split>.utf8 files (CWS tags)")
    args = ap.parse_args()
    if args.tokenizer == TokenizerGiga:
        tok = get_giga_tokenizer(args.giga_vec)
        emb = tok.embedding(args.seed)
    else:
        tok, emb = get_bert_tokenizer(args.bert_dir), None
    msr = args.format == 'msr'
    proc = get_instance(args.tokenizer, args.max_seq_len, MSR_TAG2IDX if msr else MSRA_TAG2IDX, tok)
    for file in (MSR_MAPPING if msr else MAPPING):
        print('Dumping records for {} tokenizer = {}'.format(file, args.tokenizer))
        dump_records(proc, args.src, args.out, file, mapping=MSR_MAPPING if msr else MAPPING, embedding=emb,
                     load_data=load_msr_data if msr else None)


if __name__ == '__main__':
    main()
