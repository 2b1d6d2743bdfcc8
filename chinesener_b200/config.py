"""Run / train parameters — same keys and values as reference config.py:5-29."""

TRAIN_PARAMS = {
    'dtype': 'float32',
    'lr': 5e-6,
    'log_steps': 100,
    'pretrain_dir': './pretrain_model/ch_google',  # pretrain Bert-Model
    'batch_size': 32,
    'epoch_size': 50,
    'embedding_dropout': 0.1,
    'warmup_ratio': 0.1,
    'early_stop_ratio': 1  # stop after ratio * steps_per_epoch
}

RUN_CONFIG = {
    'summary_steps': 10,
    'log_steps': 100,
    'save_steps': 500,
    'keep_checkpoint_max': 3,
    'allow_growth': True,
    'pre_process_gpu_fraction': 0.8,
    'log_device_placement': True,
    'allow_soft_placement': True,
    'inter_op_parallel': 2,
    'intra_op_parallel': 2
}

# pretrain_model/ch_google/bert_config.json of the reference (Google chinese_L-12_H-768_A-12);
# used when params['pretrain_dir'] holds no bert_config.json.
BERT_BASE_CHINESE = {
    'vocab_size': 21128, 'hidden_size': 768, 'num_hidden_layers': 12, 'num_attention_heads': 12,
    'intermediate_size': 3072, 'max_position_embeddings': 512, 'type_vocab_size': 2,
    'hidden_act': 'gelu', 'initializer_range': 0.02,
    'hidden_dropout_prob': 0.1, 'attention_probs_dropout_prob': 0.1,   # applied by BertModel(is_training=True) only
}
