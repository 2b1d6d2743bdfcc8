# -*-coding:utf-8 -*-
"""Entity extraction for the inference path (reference tools/infer_utils.py:76-118)."""
from collections import defaultdict


def extract_entity(tokens, pred_ids, idx2tag):
    """Collect the entity strings of every type from a BIO tag sequence (reference :76-101)."""
    assert len(tokens) == len(pred_ids), '{}!={} tokens and pred_ids must have same length'.format(len(tokens), len(pred_ids))
    ngram = ''
    entity = defaultdict(set)
    prev_tag = idx2tag[pred_ids[0]]
    for t, i in zip(tokens, pred_ids):
        tag = idx2tag[i]
        if tag.split('-')[0] == 'I':
            if prev_tag[0].split('-')[0] in ['B', 'I']:
                ngram += t
        else:
            if ngram != '':
                entity[prev_tag.split('-')[1]].add(ngram)
            ngram = t if tag.split('-')[0] == 'B' else ''
        prev_tag = tag
    if ngram != '':
        entity[prev_tag.split('-')[1]].add(ngram)
    return entity


def fix_tokens(sentence, tokens):
    """Put the original characters back where WordPiece produced [UNK] or a ## continuation (reference :104-118)."""
    j = 0
    for i in range(len(tokens)):
        if tokens[i] == '[UNK]':
            tokens[i] = sentence[j]
            j += 1
        elif tokens[i][:2] == '##':
            tokens[i] = tokens[i].replace('##', '')
            j += len(tokens[i])
        elif tokens[i] in ['[PAD]', '[CLS]', '[SEP]']:
            continue
        else:
            j += len(tokens[i])
    return tokens
