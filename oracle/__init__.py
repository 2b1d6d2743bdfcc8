"""CPU restatement of the reference's bert_bilstm_crf hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package, and only as the checker or the timed CPU baseline.  Nothing under
chinesener_b200/ imports it; the product path fails loudly when libner_b200.so is missing.

PARITY STATUS: the arithmetic of this path lives in un-vendored third-party code
(tensorflow==1.14.0: tf.contrib.crf, tf.nn.rnn_cell.LSTMCell, bidirectional_dynamic_rnn;
bert-base==0.0.9: modeling.BertModel — reference requirement.txt:6,44) that cannot be
installed here (Python 3.12, no wheel, no network).  Each function restates the published
algorithm and cites the reference call site it follows.  What the reference's own artefacts
pin is checked in tests/test_golden.py (prediction pickles: crf_decode zero-fill + F1 tables;
tener shift docstring; decode_prediction example; warm-up request feature layout).  Numeric
values of logits / log-likelihoods are NOT pinned by any reference fixture ("parity unpinned"
for those) and are cross-checked against brute-force enumeration instead.

crf_c.c / native.py: the same CRF decode and log-likelihood in plain C (gcc, OpenMP), pinned to crf.py bit for bit
(tests/test_oracle_native.py); bench.py uses it to re-run every row of the roofline-sized CRF launches.
"""
