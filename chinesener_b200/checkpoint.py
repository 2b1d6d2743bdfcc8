# -*-coding:utf-8 -*-
"""Checkpoints of a VariableStore: variables under their TF names, the Adam slots and `global_step` — what the
reference's tf.estimator checkpoints hold (serving_model/*/variables/variables.index lists `global_step`; training
checkpoints add `<var>/adam_m`, `<var>/adam_v` / `<var>/Adam`, `<var>/Adam_1`), so a resumed run continues its LR
schedule, bias correction and moments instead of restarting them (reference tools/utils.py:52-66 warm-starts from the
latest `*ckpt*` in model_dir).

File format: one uncompressed `.npz` per checkpoint, `model.ckpt-<global_step>.npz`; keys are TF variable names,
slots are stored as `<name>/adam_m` and `<name>/adam_v`, the step as `global_step` (int64 scalar).
"""
import glob
import os
import re

import numpy as np
import torch

SLOT_M, SLOT_V = '/adam_m', '/adam_v'


def checkpoint_path(model_dir, step):
    return os.path.join(model_dir, 'model.ckpt-{}.npz'.format(int(step)))


def all_checkpoints(model_dir):
    out = []
    for p in glob.glob(os.path.join(model_dir, 'model.ckpt-*.npz')):
        m = re.search(r'model\.ckpt-(\d+)\.npz$', p)
        if m:
            out.append((int(m.group(1)), p))
    return [p for _, p in sorted(out)]


def latest_checkpoint(model_dir):
    ck = all_checkpoints(model_dir) if model_dir and os.path.isdir(model_dir) else []
    return ck[-1] if ck else None


def save_checkpoint(store, model_dir, keep_checkpoint_max=3):
    """-> path.  Keeps the newest `keep_checkpoint_max` files (RUN_CONFIG['keep_checkpoint_max'], reference config.py:24)."""
    os.makedirs(model_dir, exist_ok=True)
    arrays = {k: v.detach().cpu().numpy() for k, v in store.vars.items()}
    fs = getattr(store, '_flat_state', None)
    if fs is not None:
        for n, (m, v) in fs.slot_dict().items():
            arrays[n + SLOT_M] = m.detach().cpu().numpy()
            arrays[n + SLOT_V] = v.detach().cpu().numpy()
    arrays['global_step'] = np.asarray(store.global_step, np.int64)
    path = checkpoint_path(model_dir, store.global_step)
    tmp = path + '.tmp.npz'
    np.savez(tmp, **arrays)
    os.replace(tmp, path)
    for old in all_checkpoints(model_dir)[:-keep_checkpoint_max]:
        os.remove(old)
    return path


def restore_checkpoint(store, path, strict=False):
    """Variables by name, then slots and global_step.  Slots wait in `store._slot_init` until the train op builds its
    flat optimizer state (tools/train_utils.FlatState picks them up)."""
    with np.load(path) as z:
        names = list(z.files)
        variables = {k: z[k] for k in names if k != 'global_step' and not k.endswith(SLOT_M) and not k.endswith(SLOT_V)}
        slots = {k[:-len(SLOT_M)]: (z[k], z[k[:-len(SLOT_M)] + SLOT_V]) for k in names if k.endswith(SLOT_M)}
        step = int(z['global_step']) if 'global_step' in names else 0
    store.load_state_dict({k: torch.from_numpy(v) for k, v in variables.items()}, strict=strict)
    store.global_step = step
    fs = getattr(store, '_flat_state', None)
    if fs is not None:
        fs.load_slots(slots)
    else:
        store._slot_init = slots
    return step
