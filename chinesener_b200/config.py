"""Defaults every plugin's TRAIN_PARAMS starts from (keys and values of reference config.py:5-16, the contract the
plugins and train ops read), and the BERT-base-Chinese architecture used when no bert_config.json is given."""

TRAIN_PARAMS = dict(
    dtype='float32',
    lr=5e-6,
    batch_size=32,
    epoch_size=50,
    log_steps=100,
    warmup_ratio=0.1,
    embedding_dropout=0.1,
    early_stop_ratio=1,                              # stop after ratio * steps_per_epoch without improvement
    pretrain_dir='./pretrain_model/ch_google',       # BertModel checkpoint + bert_config.json + vocab.txt
)

# pretrain_model/ch_google/bert_config.json of the reference (Google chinese_L-12_H-768_A-12);
# used when params['pretrain_dir'] holds no bert_config.json.
BERT_BASE_CHINESE = {
    'vocab_size': 21128, 'hidden_size': 768, 'num_hidden_layers': 12, 'num_attention_heads': 12,
    'intermediate_size': 3072, 'max_position_embeddings': 512, 'type_vocab_size': 2,
    'hidden_act': 'gelu', 'initializer_range': 0.02,
    'hidden_dropout_prob': 0.1, 'attention_probs_dropout_prob': 0.1,   # applied by BertModel(is_training=True) only
}
