"""numpy restatement of the reference's two optimizer steps (tools/train_utils.py:246-390).

  adam_weight_decay_step  <- bert optimization.AdamWeightDecayOptimizer as configured at
                             tools/train_utils.py:276-282 (no bias correction, decoupled decay,
                             exclude LayerNorm/layer_norm/bias) after clip_by_global_norm(1.0) (:315)
  tf_adam_step            <- tf.train.AdamOptimizer after clip_by_value(-5, 5)
                             (custom_train_op / gradient_clipping :340-350, 378-390)
  bert_lr / decayed_lr    <- create_optimizer's warm-up + linear decay (:252-274) and
                             tf.train.exponential_decay(staircase=True) (:365-376)
"""
import numpy as np


def clip_by_global_norm(grads, clip_norm=1.0):
    gn = np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads))
    scale = clip_norm / max(gn, clip_norm)
    return [g * scale for g in grads], gn


def adam_weight_decay_step(p, g, m, v, lr, name, b1=0.9, b2=0.999, eps=1e-6, wd=0.01):
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    upd = m / (np.sqrt(v) + eps)
    if not any(tok in name for tok in ("LayerNorm", "layer_norm", "bias")):
        upd = upd + wd * p
    return p - lr * upd, m, v


def tf_adam_step(p, g, m, v, lr, t, b1=0.9, b2=0.999, eps=1e-8, clip_value=5.0):
    g = np.clip(g, -clip_value, clip_value)
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    return p - lr_t * m / (np.sqrt(v) + eps), m, v


def bert_lr(init_lr, global_step, num_train_steps, num_warmup_steps):
    lr = init_lr * max(0.0, 1.0 - min(global_step, num_train_steps) / num_train_steps)   # polynomial_decay power 1, end 0
    if num_warmup_steps and global_step < num_warmup_steps:
        lr = init_lr * global_step / num_warmup_steps
    return lr


def decayed_lr(init_lr, global_step, step_per_epoch, decay_rate):
    return init_lr * decay_rate ** (global_step // step_per_epoch)
