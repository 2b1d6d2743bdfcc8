"""Copies the reference's serving checkpoints' `variables.index` files (SURVEY 8c golden (3): names, shapes, dtypes of every
checkpoint variable — 8.6 KB of table data each; the tensor data files are LFS pointers) into tests/golden/variables_index/
and writes their decoded listing beside them.  Run in the authoring container (needs /root/reference):

    python tests/golden/make_variables_index_golden.py
"""
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from chinesener_b200 import tf_checkpoint  # noqa: E402

REF = "/root/reference/serving_model"
MODELS = ("bert_bilstm_crf", "bilstm_crf", "bilstm_crf_softlexicon", "bert_bilstm_crf_mtl")

if __name__ == "__main__":
    out = {}
    for m in MODELS:
        src = os.path.join(REF, m, "1", "variables", "variables.index")
        dst = os.path.join(HERE, "variables_index", m + ".index")
        shutil.copyfile(src, dst)
        header, entries = tf_checkpoint.read_bundle_index(dst)
        out[m] = {"num_shards": header["num_shards"], "total_bytes": sum(e["size"] for e in entries.values()),
                  "variables": {k: {"dtype": e["dtype"], "shape": e["shape"]} for k, e in entries.items()}}
    with open(os.path.join(HERE, "variables_index.json"), "w") as f:
        json.dump(out, f, indent=1)
    print({m: (len(v["variables"]), v["total_bytes"]) for m, v in out.items()})
