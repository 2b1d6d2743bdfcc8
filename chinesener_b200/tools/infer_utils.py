"""Entity extraction for the in-process inference path — behaviour of reference tools/infer_utils.py:76-118
(`extract_entity`, `fix_tokens`), which `inference.InferHelper.infer` applies to the PREDICT output."""
from collections import defaultdict

_SPECIAL_TOKENS = frozenset(('[PAD]', '[CLS]', '[SEP]'))


def extract_entity(tokens, pred_ids, idx2tag):
    """{entity type: set of surface strings} read off a BIO tag sequence.

    Rules kept from the reference (:76-101): an `I-*` tag extends the open span only when the previous tag starts with
    B or I; any other tag closes the span, whose type is the type of the tag just before; `B-*` opens a new one."""
    if len(tokens) != len(pred_ids):
        raise AssertionError('{}!={} tokens and pred_ids must have same length'.format(len(tokens), len(pred_ids)))
    found = defaultdict(set)
    pieces, last = [], idx2tag[pred_ids[0]]

    def close():
        text = ''.join(pieces)
        if text != '':
            found[last.split('-')[1]].add(text)

    for tok, idx in zip(tokens, pred_ids):
        tag = idx2tag[idx]
        kind = tag.split('-')[0]
        if kind == 'I':
            if last[:1] in ('B', 'I'):
                pieces.append(tok)
        else:
            close()
            pieces = [tok] if kind == 'B' else []
        last = tag
    close()
    return found


def fix_tokens(sentence, tokens):
    """Restore the raw characters WordPiece replaced: `[UNK]` becomes the character under the cursor, `##xx` loses its
    continuation mark (reference :104-118).  Edits `tokens` in place and returns it."""
    cursor = 0
    for k, tok in enumerate(tokens):
        if tok in _SPECIAL_TOKENS:
            continue
        if tok == '[UNK]':
            tokens[k] = sentence[cursor]
            cursor += 1
            continue
        if tok.startswith('##'):
            tok = tokens[k] = tok.replace('##', '')
        cursor += len(tok)
    return tokens


def extract_entity_device(tokens_batch, pred_ids, idx2tag, _cache={}):
    """Batched extract_entity with the tag scan on the GPU (ner_extract_spans): pred_ids [B, L] int32 on the device,
    tokens_batch B lists of L token strings -> list of {type: set of surface strings}, equal to
    [extract_entity(tokens, pred_ids[b], idx2tag) for b ...].  Only the spans (4 bytes each) cross to the host."""
    from .. import ops
    key = tuple(sorted(idx2tag.items()))
    ent = _cache.get(key)
    if ent is None:
        table, types = ops.tag_classes(idx2tag)
        ent = _cache[key] = (table.to(pred_ids.device), types)
    table, types = ent
    if table.device != pred_ids.device:
        table = table.to(pred_ids.device)
    spans, counts = ops.extract_spans(pred_ids, table)
    counts = counts.cpu().numpy()
    spans = spans[:, :max(int(counts.max()), 1)].cpu().numpy() if len(counts) else spans.cpu().numpy()
    out = []
    for b, toks in enumerate(tokens_batch):
        found = defaultdict(set)
        for w in spans[b, :counts[b]]:
            s, e, t = int(w) & 0xFFF, (int(w) >> 12) & 0xFFF, int(w) >> 24
            text = ''.join(toks[s:e])
            if text != '':
                found[types[t]].add(text)
        out.append(found)
    return out
