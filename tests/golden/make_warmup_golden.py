"""Extracts the featurised warm-up sentence of the reference's exported serving models
(serving_model/<model>/1/assets.extra/tf_serving_warmup_requests, written by reference warmup.py:11-26 through
inference.InferHelper.make_feature) into tests/golden/warmup_features.json.  The records are TFRecord-framed
PredictionLog protos whose single input is a serialized tf.train.Example; both are decoded with a minimal protobuf
wire-format reader (no TensorFlow needed).  Run in the authoring container: `python tests/golden/make_warmup_golden.py`."""
import json
import os
import struct

REF = "/root/reference"
TEXT = '给中央军委委员、总参谋长傅全有上将致唁函的有:美国太平洋总部司令布鲁赫海军上将。'   # warmup.py:15


def varint(b, i):
    r = s = 0
    while True:
        c = b[i]
        i += 1
        r |= (c & 0x7F) << s
        s += 7
        if not c & 0x80:
            return r, i


def fields(b):
    i = 0
    while i < len(b):
        key, i = varint(b, i)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, i = varint(b, i)
        elif wt == 2:
            n, i = varint(b, i)
            v = b[i:i + n]
            i += n
        elif wt == 1:
            v = b[i:i + 8]
            i += 8
        else:
            v = b[i:i + 4]
            i += 4
        yield f, wt, v


def sub(b, path):
    for f, wt, v in fields(b):
        if f == path[0] and wt == 2:
            return v if len(path) == 1 else sub(v, path[1:])


def first_record(path):
    b = open(path, 'rb').read()
    n = struct.unpack('<Q', b[:8])[0]
    return b[12:12 + n]


def example_features(ex):
    out = {}
    for f, _, entry in fields(sub(ex, [1])):
        kv = {ff: vv for ff, _, vv in fields(entry)}
        name = kv[1].decode()
        for ff, _, vv in fields(kv[2]):
            if ff == 1:       # bytes_list
                out[name] = [x.decode('utf8') for _, _, x in fields(vv)]
            elif ff == 2:     # float_list (packed)
                vals = []
                for _, wt, c in fields(vv):
                    vals.extend(struct.unpack('<%df' % (len(c) // 4), c))
                out[name] = vals
            elif ff == 3:     # int64_list (packed)
                vals = []
                for _, wt, c in fields(vv):
                    if wt == 2:
                        j = 0
                        while j < len(c):
                            x, j = varint(c, j)
                            vals.append(x)
                    else:
                        vals.append(c)
                out[name] = vals
    return out


def main():
    golden = {"text": TEXT, "max_seq_len": 150, "models": {}}
    for model in ("bert_bilstm_crf", "bilstm_crf", "bilstm_crf_softlexicon"):
        rec = first_record(os.path.join(REF, "serving_model", model, "1", "assets.extra", "tf_serving_warmup_requests"))
        req = sub(rec, [6, 1])                                   # PredictionLog.predict_log.request
        for f, _, entry in fields(req):
            if f == 2:                                           # inputs map: 'example' -> TensorProto(string_val)
                kv = {ff: vv for ff, _, vv in fields(entry)}
                serialized = [v for ff, _, v in fields(kv[2]) if ff == 8][0]
                golden["models"][model] = example_features(serialized)
    # the vocabulary entries the sentence touches (token -> id), so the tokenizer test runs without vocab.txt
    b = golden["models"]["bert_bilstm_crf"]
    golden["bert_vocab_subset"] = {t: i for t, i in zip(b["tokens"], b["token_ids"])}
    g = golden["models"]["bilstm_crf"]
    golden["giga_vocab_subset"] = {t: i for t, i in zip(g["tokens"], g["token_ids"]) if t != '[PAD]'}
    golden["giga_n_vocab"] = g["token_ids"][-1]                   # '[PAD]' = n_vocab
    # SoftLexicon record: keep the layout facts and the first rows (6000 ids / float weights per sentence)
    sl = golden["models"].pop("bilstm_crf_softlexicon")
    ids, wts = sl["softlexicon_ids"], sl["softlexicon_weights"]
    golden["softlexicon"] = {"seq_len": sl["seq_len"], "n_ids": len(ids), "ids_first_rows": ids[:120],
                             "weights_first_rows": wts[:120], "ids_row_40": ids[40 * 40:41 * 40],
                             "weights_row_40": wts[40 * 40:41 * 40],
                             "row_weight_sums": [sum(wts[r * 40:(r + 1) * 40]) for r in range(150)],
                             "none_id": max(ids) - 1, "pad_id": max(ids), "token_ids": sl["token_ids"]}
    with open(os.path.join(os.path.dirname(__file__), "warmup_features.json"), "w", encoding="utf8") as f:
        json.dump(golden, f, ensure_ascii=False, indent=0)


if __name__ == "__main__":
    main()
