"""model_fn-equivalent glue: picks a plugin by name and runs PREDICT / EVAL steps.

Mirrors reference tools/train_utils.py:145-187 (build_model_fn): the plugin is looked up by
`importlib` under `model.<name>`, called as build_graph(features, labels, params, is_training)
and its PREDICT output dict carries the keys 'pred_ids', 'label_ids', 'tokens'
(tools/train_utils.py:181-185).  Host batches come in as (pinned) CPU tensors and are copied
to the device inside `predict` — that copy is part of the end-to-end number bench.py reports.
"""
import contextlib
import importlib

import torch

from . import autodiff, variables

_DEVICE_KEYS = ('token_ids', 'mask', 'segment_ids', 'label_ids', 'seq_len', 'softlexicon_ids', 'softlexicon_weights',
                'bichar_ids', 'softword_ids', 'ex_softword_ids', 'task_ids')


def load_plugin(model_name):
    mod = importlib.import_module(f'{__package__}.model.{model_name}')
    return getattr(mod, 'build_graph'), getattr(mod, 'TRAIN_PARAMS')


class Estimator:
    """Minimal stand-in for tf.estimator.Estimator over one plugin + one variable store."""

    def __init__(self, model_name, params, store=None, device='cuda'):
        self.model_name = model_name
        self.build_graph, train_params = load_plugin(model_name)
        self.params = dict(train_params)
        self.params.update(params)
        self.device = torch.device(device)
        self.store = store or variables.VariableStore(self.device)
        # data-parallel gradient exchange (N > 1): 'overlap' = bucketed all-reduces behind the backward pass, 'overlap_bf16' =
        # the same with bf16 buckets, 'single' = one all-reduce of the flat buffer after the backward pass
        self.store.grad_exchange = self.params.get('grad_exchange', 'overlap')

    def to_device(self, features):
        out = {}
        for k, v in features.items():
            if k in _DEVICE_KEYS and torch.is_tensor(v):
                out[k] = v.to(self.device, non_blocking=True)
            else:
                out[k] = v
        m = features.get('mask')
        if torch.is_tensor(m) and not m.is_cuda and 'mask' in out:
            # host-side token count rides along so sequence packing needs no device sync
            out['mask'].total_tokens = int(m.sum())
        return out

    def predict_device(self, dev_features):
        """PREDICT on device-resident features -> pred_ids (device).  One fused C call when the plugin has an
        executor in fastpath.FUSED_PREDICT (same kernels as build_graph, see fastpath.py), else build_graph."""
        if self.params.get('fused_predict', True):
            from . import fastpath
            fn = fastpath.FUSED_PREDICT.get(self.model_name)
            if fn is not None:
                with variables.use_store(self.store):
                    pred = fn(self, dev_features)
                if pred is not None:
                    return pred
        return self.forward_device(dev_features, False)[1]      # multi-task plugins return (loss, pred_ids, task_ids)

    def forward_device(self, dev_features, is_training=False):
        from .tools import layer
        prec0 = layer.BERT_PRECISION
        layer.BERT_PRECISION = self.params.get('bert_precision', prec0)
        try:
            with variables.use_store(self.store):
                return self.build_graph(dev_features, None, self.params, is_training)
        finally:
            layer.BERT_PRECISION = prec0

    def predict(self, features):
        """PREDICT mode on one host batch -> dict(pred_ids int32 [B,L] on host, label_ids, tokens)."""
        dev = self.to_device(features)
        pred_ids = self.predict_device(dev)
        return {'pred_ids': pred_ids.cpu(), 'label_ids': features.get('label_ids'), 'tokens': features.get('tokens')}

    def stack_to_device(self, feature_list):
        """Several host batches -> ONE device batch (rows concatenated in order): every device feature is allocated once
        at the summed batch size and each host part is copied straight into its row slice (pinned -> non-blocking)."""
        if len(feature_list) == 1:
            return self.to_device(feature_list[0])
        out, total = {}, 0
        first = feature_list[0]
        for k, v in first.items():
            if k in _DEVICE_KEYS and torch.is_tensor(v):
                rows = sum(f[k].shape[0] for f in feature_list)
                dst = torch.empty((rows,) + tuple(v.shape[1:]), dtype=v.dtype, device=self.device)
                r = 0
                for f in feature_list:
                    n = f[k].shape[0]
                    dst[r:r + n].copy_(f[k], non_blocking=True)
                    r += n
                out[k] = dst
        for f in feature_list:
            m = f.get('mask')
            if not (torch.is_tensor(m) and not m.is_cuda):
                total = None
                break
            total += int(m.sum())
        if total is not None and 'mask' in out:
            out['mask'].total_tokens = total
        return out

    def predict_iter(self, batches, depth=2, streams=1, group=1):
        """Generator form of PREDICT — the shape of tf.estimator.Estimator.predict(input_fn), which the
        reference drives at main.py:52-55: yields one result dict per host batch, in order.  The device work
        of up to `depth` calls is in flight before the oldest result is awaited, so the host->device copy
        of batch i+1 and the enqueue of its kernels overlap batch i on the GPU; results come back through
        a small ring of pinned host buffers.
        streams > 1: consecutive calls run on different CUDA streams (sentences are independent, SURVEY
        8(e)), so the SMs a kernel of one call leaves idle are taken by the other call's kernels.
        group > 1: `group` consecutive host batches are stacked into one device batch per call (sentences are
        independent, so the tags are those of the separate calls): the packed token count of a 64-sentence MSRA batch
        (~3.2 k rows) fills 0.5 / 1.5 / 2.0 waves of 128x256 GEMM tiles on 148 SMs, two batches (~6.3 k rows) fill
        1.0 / 3.0 / 4.0 — the tensor-core tiles stop idling in partial waves without relying on stream overlap."""
        from . import ops
        ring, inflight = {}, []
        depth = max(depth, streams + 1) if streams > 1 else depth
        side = [torch.cuda.Stream() for _ in range(streams)] if streams > 1 else None
        if side is not None:
            torch.cuda.synchronize()          # weight packs / caches built on the caller's stream are complete

        def finish(item):
            ev, buf, feats_list = item
            ev.synchronize()
            r = 0
            for feats in feats_list:
                n = feats['token_ids'].shape[0] if torch.is_tensor(feats.get('token_ids')) else buf.shape[0]
                yield {'pred_ids': buf[r:r + n].clone(), 'label_ids': feats.get('label_ids'), 'tokens': feats.get('tokens')}
                r += n

        def grouped(it):
            cur = []
            for f in it:
                cur.append(f)
                if len(cur) == group:
                    yield cur
                    cur = []
            if cur:
                yield cur

        # streams > 1: the host->device copies of a call run on their own stream, so the inputs of call k+1 travel while the
        # compute streams are busy with calls k-1 / k (a compute stream that copies its own inputs idles for the ~20 small
        # transfers of a stacked call)
        copy_stream = torch.cuda.Stream() if side is not None else None
        k = 0
        for feats_list in grouped(batches):
            ctx = torch.cuda.stream(side[k % streams]) if side is not None else contextlib.nullcontext()
            if copy_stream is not None:
                with torch.cuda.stream(copy_stream):
                    dev = self.stack_to_device(feats_list)
                    copied = torch.cuda.Event()
                    copied.record()
            with ctx:
                if copy_stream is not None:
                    st = side[k % streams]
                    st.wait_event(copied)
                    for t in dev.values():
                        if torch.is_tensor(t) and t.is_cuda:
                            t.record_stream(st)          # allocated on the copy stream, consumed here
                else:
                    dev = self.stack_to_device(feats_list)
                tile0 = ops.DEFAULT_TILE
                if side is not None or group > 1:   # partial waves are filled (other streams / 2x rows): take the fastest tile
                    ops.DEFAULT_TILE = ops.TILE_AUTO_THROUGHPUT
                try:
                    pred_ids = self.predict_device(dev)
                finally:
                    ops.DEFAULT_TILE = tile0
                key = (tuple(pred_ids.shape), k % (depth + 1))
                buf = ring.get(key)
                if buf is None:
                    buf = ring[key] = torch.empty(pred_ids.shape, dtype=pred_ids.dtype, pin_memory=True)
                buf.copy_(pred_ids, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            inflight.append((ev, buf, feats_list))
            k += 1
            if len(inflight) >= depth:
                yield from finish(inflight.pop(0))
        while inflight:
            yield from finish(inflight.pop(0))
        if side is not None:
            for st in side:
                torch.cuda.current_stream().wait_stream(st)

    def predict_sentences(self, batches, depth=2, streams=1, group=1):
        """tf.estimator.Estimator.predict(input_fn) as the reference consumes it (main.py:52-55, evaluation.py:16-24):
        one dict PER SENTENCE — {'pred_ids': int32 [L], 'label_ids': int32 [L], 'tokens': [L]} as numpy arrays / lists —
        the element type of `<model>_predict.pkl` and the input of evaluation.SingleEval.  (predict / predict_iter yield
        one dict per batch.)"""
        for out in self.predict_iter(batches, depth=depth, streams=streams, group=group):
            pred = out['pred_ids'].numpy()
            lab = out['label_ids'].numpy() if torch.is_tensor(out['label_ids']) else out['label_ids']
            tok = out['tokens']
            for b in range(pred.shape[0]):
                yield {'pred_ids': pred[b], 'label_ids': None if lab is None else lab[b], 'tokens': None if tok is None else tok[b]}

    def train_step(self, features):
        """TRAIN mode of model_fn (reference tools/train_utils.py:151-168): forward with the tape,
        backward, then the train op the reference picks by model name (:156-164).  -> loss (float)."""
        from .tools import train_utils
        dev = features if all(not torch.is_tensor(v) or v.is_cuda for v in features.values()) else self.to_device(features)
        self.store.dropout_calls = 0
        with variables.use_store(self.store), autodiff.recording(self.store) as tape:
            loss = self.build_graph(dev, None, self.params, True)[0]
            tape.backward()
            p = self.params
            if 'bert' in self.model_name:
                train_utils.bert_train_op(loss, p['lr'], p['num_train_steps'], p['warmup_ratio'], p['diff_lr_times'])
            elif 'transformer' in self.model_name:
                train_utils.transformer_train_op(loss, p['lr'], p['num_train_steps'], p['warmup_ratio'])
            else:
                train_utils.custom_train_op(loss, p['lr'], p['step_per_epoch'], p['decay_rate'])
        return loss

    def evaluate(self, features):
        dev = self.to_device(features)
        loss, pred_ids = self.forward_device(dev, False)[:2]
        return {'loss': float(loss), 'pred_ids': pred_ids.cpu()}
