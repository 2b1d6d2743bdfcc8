# -*-coding:utf-8 -*-
"""Train / evaluate / predict driver with the reference's command line (reference main.py:14-140).

    python -m chinesener_b200.main --model_name bilstm_crf --data msra [--clear_model 1] [--data_dir datasets/msra]

What `singletask_train` does, in the reference's order (main.py:14-62):
  * TRAIN_PARAMS of the plugin + the dataset's params (NerDataset -> label_size, max_seq_len, step_per_epoch,
    num_train_steps, embedding ...);
  * an Estimator over `./checkpoint/ner_<data>_<model>`, warm-started from its latest checkpoint (tools/utils.py:52-66);
  * tf.estimator.train_and_evaluate: TRAIN over shuffle(64).repeat(epoch_size).batch(batch_size); a checkpoint every
    RUN_CONFIG['save_steps'] steps, each followed by an EVAL pass over the `valid` split; the
    stop_if_no_decrease_hook(metric 'loss', max_steps_without_decrease = step_per_epoch * early_stop_ratio) ends
    training when the best eval loss is that many steps old;
  * PREDICT over the `predict` split with the LAST checkpoint's weights; the list of per-sentence dicts
    {'pred_ids' int32[L], 'label_ids' int32[L], 'tokens' bytes[L]} is pickled to `<data_dir>/<model>_predict.pkl`
    (main.py:52-55) — the file evaluation.py scores.
The Estimator's own evaluation cadence is time based (throttle_secs=60, the hook polls every 60 s); on a B200 a whole
epoch takes seconds, so the cadence here is the step-based one those timers converge to on the reference's hardware:
evaluate at every checkpoint.  Exporting a SavedModel (main.py:57-60) has no counterpart: the in-process InferHelper
serves from the checkpoint.
"""
import argparse
import json
import os
import pickle
import shutil
import time

import numpy as np

RUN_CONFIG = {'summary_steps': 10, 'log_steps': 100, 'save_steps': 500, 'keep_checkpoint_max': 3}     # config.py:19-30


def clear_model(model_dir):
    """tools/utils.py:17-23"""
    try:
        shutil.rmtree(model_dir)
    except Exception as e:
        print('Error! {} occured at model cleaning'.format(e))
    else:
        print('{} model cleaned'.format(model_dir))


def evaluate(estimator, input_fn):
    """EVAL pass: mean of the per-batch losses (tf.estimator averages the `loss` metric over batches) + tag accuracy over
    the real tokens (tools/train_utils.py:107-127 weights by the non-[PAD] mask)."""
    losses, right, total = [], 0, 0
    for feats in input_fn():
        out = estimator.evaluate(feats)
        losses.append(out['loss'])
        lab, pred = feats['label_ids'].numpy(), out['pred_ids'].numpy()
        real = lab > 0
        right += int(((lab == pred) & real).sum())
        total += int(real.sum())
    return {'loss': float(np.mean(losses)), 'accuracy': right / max(total, 1), 'batches': len(losses)}


def train_and_evaluate(estimator, input_pipe, model_dir, run_config=RUN_CONFIG, log=print, max_steps=None):
    """-> history dict.  See the module docstring for the correspondence with tf.estimator.train_and_evaluate."""
    from . import checkpoint
    p = estimator.params
    max_no_decrease = int(p['step_per_epoch'] * p['early_stop_ratio'])
    evals, best = [], (None, None)               # best = (loss, step)
    t0 = time.time()
    loss_sum, loss_n = None, 0
    stopped = 'input exhausted'
    store = estimator.store

    def checkpoint_and_eval():
        nonlocal best
        path = checkpoint.save_checkpoint(store, model_dir, run_config['keep_checkpoint_max'])
        ev = evaluate(estimator, input_pipe.build_input_fn('valid', is_predict=True, with_strings=False))
        ev.update(step=store.global_step, seconds=round(time.time() - t0, 1))
        evals.append(ev)
        if best[0] is None or ev['loss'] < best[0]:
            best = (ev['loss'], store.global_step)
        log('eval @ step {step}: loss = {loss:.4f} accuracy = {accuracy:.4f} ({seconds}s) -> {0}'.format(os.path.basename(path), **ev))
        # stop_if_no_decrease_hook: the best (lowest) eval loss is at least max_steps_without_decrease steps old
        return store.global_step - best[1] >= max_no_decrease

    for feats in input_pipe.build_input_fn('train')():
        loss = estimator.train_step(feats)
        loss_sum = loss.detach() if loss_sum is None else loss_sum + loss.detach()
        loss_n += 1
        step = store.global_step
        if step % run_config['log_steps'] == 0:
            log('step {}: loss = {:.4f} ({:.1f}s)'.format(step, float(loss_sum) / loss_n, time.time() - t0))
            loss_sum, loss_n = None, 0
        if step % run_config['save_steps'] == 0 and checkpoint_and_eval():
            stopped = 'no decrease of the eval loss for {} steps (best {:.4f} @ {})'.format(max_no_decrease, *best)
            break
        if max_steps is not None and step >= max_steps:
            stopped = 'max_steps'
            break
    if not evals or evals[-1]['step'] != store.global_step:
        checkpoint_and_eval()                     # the Estimator always saves and evaluates at the end of training
    log('training stopped at step {}: {}'.format(store.global_step, stopped))
    return {'evals': evals, 'best_eval_loss': best[0], 'best_eval_step': best[1], 'final_step': store.global_step,
            'stopped': stopped, 'train_seconds': round(time.time() - t0, 1)}


def predict_to_list(estimator, input_fn):
    """estimator.predict(input_fn) of the reference: one dict per SENTENCE, numpy values as tf.estimator yields them."""
    out = []
    for res in estimator.predict_sentences(input_fn()):
        out.append({'pred_ids': res['pred_ids'].astype(np.int32), 'label_ids': res['label_ids'].astype(np.int32),
                    'tokens': np.array([t.encode('utf-8') if isinstance(t, str) else t for t in res['tokens']], dtype=object)})
    return out


def singletask_train(args):
    from . import checkpoint, engine
    from .data.records import NerDataset
    model_name = args.rename if args.rename else args.model_name
    model_dir = os.path.join(args.checkpoint_root, 'ner_{}_{}'.format(args.data, model_name))
    data_dir = args.data_dir or './data/{}'.format(args.data)
    if args.clear_model:
        clear_model(model_dir)

    _, TRAIN_PARAMS = engine.load_plugin(args.model_name)
    TRAIN_PARAMS = dict(TRAIN_PARAMS)
    if args.epoch_size:
        TRAIN_PARAMS['epoch_size'] = args.epoch_size
    if args.pretrain_dir:
        TRAIN_PARAMS['pretrain_dir'] = args.pretrain_dir
    if args.batch_size:
        TRAIN_PARAMS['batch_size'] = args.batch_size
    input_pipe = NerDataset(data_dir, TRAIN_PARAMS['batch_size'], TRAIN_PARAMS['epoch_size'], model_name, seed=args.seed)
    TRAIN_PARAMS.update(input_pipe.params)       # label_size, max_seq_len, num_train_steps ... (main.py:25)
    print('=' * 10 + 'TRAIN PARAMS' + '=' * 10)
    print(dict((i, j) for i, j in TRAIN_PARAMS.items() if ('emb' not in i) and ('vocab' not in i)))
    print('=' * 10 + 'RUN PARAMS' + '=' * 10)
    print(RUN_CONFIG)

    estimator = engine.Estimator(args.model_name, TRAIN_PARAMS)
    estimator.store.gen.manual_seed(args.seed)
    warm = checkpoint.latest_checkpoint(model_dir)
    if warm:
        # variables exist only after the first build_graph: run one EVAL batch, then overwrite them from the checkpoint
        first = next(iter(input_pipe.build_input_fn('valid', is_predict=True, with_strings=False)()))
        estimator.evaluate(first)
        print('warm start from {} (step {})'.format(warm, checkpoint.restore_checkpoint(estimator.store, warm)))

    history = None
    if not args.predict_only:
        history = train_and_evaluate(estimator, input_pipe, model_dir, max_steps=args.max_steps)

    prediction = predict_to_list(estimator, input_pipe.build_input_fn('predict', is_predict=True))
    out_pkl = os.path.join(data_dir, '{}_predict.pkl'.format(model_name))
    with open(out_pkl, 'wb') as f:
        pickle.dump(prediction, f)
    print('{} sentences -> {}'.format(len(prediction), out_pkl))

    from .evaluation import SingleEval
    tag_rep, ent_rep = SingleEval(prediction, TRAIN_PARAMS['idx2tag']).gen_report()
    summary = {'model': model_name, 'data': args.data, 'history': history, 'n_predict': len(prediction), 'seed': args.seed,
               'entity_micro_f1': ent_rep['micro avg']['f1-score'], 'entity_weighted_f1': ent_rep['weighted avg']['f1-score'],
               'entity_report': ent_rep, 'tag_weighted_f1': tag_rep['weighted avg']['f1-score']}
    print('entity micro-F1 = {:.4f}  weighted-F1 = {:.4f}'.format(summary['entity_micro_f1'], summary['entity_weighted_f1']))
    if args.report:
        with open(args.report, 'w') as f:
            json.dump(summary, f, indent=1, default=float)
    return summary


def multitask_train(args):
    """reference main.py:65-118: `--data a,b` with an mtl / adv plugin.  One Estimator over
    `./checkpoint/ner_<a_b>_<model>`, TRAIN / EVAL over the sample-by-sample mix of the datasets (MultiDataset), then one
    PREDICT pass per dataset with that dataset's task id -> `<data_root>/<data>/<model>_<a_b>_predict.pkl`."""
    from . import checkpoint, engine
    from .data.records import MultiDataset
    model_name = args.rename if args.rename else args.model_name
    data_list = args.data.split(',')
    joined = '_'.join(data_list)
    model_dir = os.path.join(args.checkpoint_root, 'ner_{}_{}'.format(joined, model_name))
    data_root = args.data_dir or './data'
    if args.clear_model:
        clear_model(model_dir)

    _, TRAIN_PARAMS = engine.load_plugin(args.model_name)
    TRAIN_PARAMS = dict(TRAIN_PARAMS)
    if args.epoch_size:
        TRAIN_PARAMS['epoch_size'] = args.epoch_size
    if args.pretrain_dir:
        TRAIN_PARAMS['pretrain_dir'] = args.pretrain_dir
    if args.batch_size:
        TRAIN_PARAMS['batch_size'] = args.batch_size
    input_pipe = MultiDataset(data_root, data_list, TRAIN_PARAMS['batch_size'], TRAIN_PARAMS['epoch_size'], model_name, seed=args.seed)
    TRAIN_PARAMS.update(input_pipe.params)       # per-dataset params, task_list, step_per_epoch, num_train_steps, max_seq_len
    print('=' * 10 + 'TRAIN PARAMS' + '=' * 10)
    print(dict((i, j) for i, j in TRAIN_PARAMS.items() if i not in data_list))
    print('=' * 10 + 'RUN PARAMS' + '=' * 10)
    print(RUN_CONFIG)

    estimator = engine.Estimator(args.model_name, TRAIN_PARAMS)
    estimator.store.gen.manual_seed(args.seed)
    warm = checkpoint.latest_checkpoint(model_dir)
    if warm:
        first = next(iter(input_pipe.build_input_fn('valid', is_predict=True)()))
        estimator.evaluate(first)
        print('warm start from {} (step {})'.format(warm, checkpoint.restore_checkpoint(estimator.store, warm)))

    history = None
    if not args.predict_only:
        history = train_and_evaluate(estimator, input_pipe, model_dir, max_steps=args.max_steps)

    summary = {'model': model_name, 'data': data_list, 'history': history, 'seed': args.seed, 'tasks': {}}
    for data in data_list:
        print('Prediction for {}'.format(data))
        prediction = predict_to_list(estimator, input_pipe.build_predict_fn(data))
        out_pkl = os.path.join(data_root, data, '{}_{}_predict.pkl'.format(model_name, joined))
        with open(out_pkl, 'wb') as f:
            pickle.dump(prediction, f)
        # the entity report of evaluation.py is NER specific (a B/I/E/S segmentation tag has no entity type): tag accuracy here
        real = [(int(a), int(b)) for i in prediction for a, b in zip(i['label_ids'], i['pred_ids']) if a > 0]
        acc = sum(a == b for a, b in real) / max(len(real), 1)
        summary['tasks'][data] = {'n_predict': len(prediction), 'file': out_pkl, 'tag_accuracy': acc}
        print('{} sentences -> {} (tag accuracy {:.4f})'.format(len(prediction), out_pkl, acc))
    if args.report:
        with open(args.report, 'w') as f:
            json.dump(summary, f, indent=1, default=float)
    return summary


def build_parser():
    parser = argparse.ArgumentParser()
    # the reference's flags (main.py:121-136); argparse accepts unambiguous prefixes, so `--model` works as it does there
    parser.add_argument('--model_name', type=str, help='model_name[bert_bilstm_crf, bert_crf, bilstm_crf ...]', required=True)
    parser.add_argument('--clear_model', type=int, help='Whether to clear existing model', required=False, default=0)
    parser.add_argument('--data', type=str, help='which data to use[msra, people_daily]', required=False, default='msra')
    parser.add_argument('--gpu', type=int, help='kept for compatibility: the sm_100a path always runs on the GPU', required=False, default=1)
    parser.add_argument('--device', type=int, help='which gpu to use', required=False, default=-1)
    parser.add_argument('--rename', type=str, help='Allow rename model with special parameter', required=False, default='')
    parser.add_argument('--export_only', type=int, help='kept for compatibility (no SavedModel export: InferHelper serves in-process)',
                        required=False, default=0)
    # additions
    parser.add_argument('--data_dir', type=str, default='', help='directory of the .nerrec files (default ./data/<data>); with --data a,b: the root '
                        'holding one directory per dataset (default ./data)')
    parser.add_argument('--checkpoint_root', type=str, default='./checkpoint')
    parser.add_argument('--predict_only', type=int, default=0)
    parser.add_argument('--epoch_size', type=int, default=0, help='override TRAIN_PARAMS["epoch_size"]')
    parser.add_argument('--pretrain_dir', type=str, default='', help='override TRAIN_PARAMS["pretrain_dir"] (bert_config.json [+ checkpoint])')
    parser.add_argument('--batch_size', type=int, default=0, help='override TRAIN_PARAMS["batch_size"]')
    parser.add_argument('--max_steps', type=int, default=None)
    parser.add_argument('--seed', type=int, default=1234)
    parser.add_argument('--report', type=str, default='', help='write a JSON summary (eval history + test F1) here')
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.device >= 0:
        os.environ['CUDA_VISIBLE_DEVICES'] = '{}'.format(args.device)
    if len(args.data.split(',')) > 1:
        return multitask_train(args)
    return singletask_train(args)


if __name__ == '__main__':
    main()
