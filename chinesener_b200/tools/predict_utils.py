# -*-coding:utf-8 -*-
"""Prediction post-processing with the reference's surface (reference tools/predict_utils.py)."""
from collections import defaultdict


def decode_prediction(tokens, tags):
    """reference tools/predict_utils.py:6-36 — tokens (bytes) + tag strings -> {'PER': {...}, ...}.

    B-x starts an entity, I-x of the same type extends it, an I-y of another type appends
    '[ERROR]', anything else closes it; '##' word-piece markers are stripped.
    """
    assert len(tokens) == len(tags), \
        'NER Decode {}!={}: token and pred_ids must have same len'.format(len(tokens), tags)
    result = defaultdict(set)
    entity, type1 = '', ''
    for token, tag in zip(tokens, tags):
        text = (token.decode() if isinstance(token, bytes) else token).replace('##', '')
        if 'B' in tag:
            if entity:
                result[type1].add(entity)
            entity, type1 = text, tag.split('-')[1]
        elif 'I' in tag:
            entity += text if tag.split('-')[1] == type1 else '[ERROR]'
        else:
            if entity:
                result[type1].add(entity)
            entity = ''
    if entity:
        result[type1].add(entity)
    return result


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
