# -*-coding:utf-8 -*-
"""Prediction post-processing with the reference's surface (reference tools/predict_utils.py)."""
from collections import defaultdict


def decode_prediction(tokens, tags):
    """reference tools/predict_utils.py:6-36 — tokens (bytes) + tag strings -> {'PER': {...}, ...}.

    B-x starts an entity, I-x of the same type extends it, an I-y of another type appends
    '[ERROR]', anything else closes it; '##' word-piece markers are stripped.
    """
    if len(tokens) != len(tags):
        raise AssertionError('NER Decode {}!={}: token and pred_ids must have same len'.format(len(tokens), tags))
    spans = defaultdict(set)
    state = {'text': '', 'type': ''}          # the open entity

    def flush(new_text='', new_type=None):
        if state['text']:
            spans[state['type']].add(state['text'])
        state['text'] = new_text
        if new_type is not None:
            state['type'] = new_type

    for token, tag in zip(tokens, tags):
        piece = (token.decode() if isinstance(token, bytes) else token).replace('##', '')
        if 'B' in tag:
            flush(piece, tag.split('-')[1])
        elif 'I' in tag:
            state['text'] += piece if tag.split('-')[1] == state['type'] else '[ERROR]'
        else:
            flush()
    flush()
    return spans


def process_prediction(pred_dict, idx2tag):
    """reference tools/predict_utils.py:39-60 — drop [CLS]/[SEP]/[PAD] positions, map ids -> tags."""
    rm_tag = ['[CLS]', '[PAD]', '[SEP]']
    tok = [t.decode() if isinstance(t, bytes) else t for t in pred_dict['tokens']]
    keep = [t not in rm_tag for t in tok]
    out = {k: [v for v, m in zip(val, keep) if m] for k, val in pred_dict.items() if val is not None}
    keep = [idx2tag[int(i)] not in rm_tag for i in out['pred_ids']]        # corner case only
    out = {k: [v for v, m in zip(val, keep) if m] for k, val in out.items()}
    out['sentence'] = ''.join(t.decode() if isinstance(t, bytes) else t for t in out['tokens'])
    out['preds'] = [idx2tag[int(i)] for i in out['pred_ids']]
    out['labels'] = [idx2tag[int(i)] for i in out['label_ids']]
    out['label_entity'] = decode_prediction(out['tokens'], out['labels'])
    out['pred_entity'] = decode_prediction(out['tokens'], out['preds'])
    return out
