"""Synthetic MSRA-shaped batches: the bench/test workload generator.

Feature names, dtypes and shapes follow the reference's TFRecord schema
(data/base_preprocess.py:36-71, dataset.py:21-37): token_ids / mask / segment_ids / label_ids
[B,L] int32, seq_len [B] int32 (counts [CLS]+[SEP] for BERT tokenisation,
data/base_preprocess.py:166-174), softlexicon_ids [B,L*40] int32, softlexicon_weights
[B,L*40] f32.  Tag ids: [PAD]=0, O=1 ... [CLS]=8, [SEP]=9 (data/msra/preprocess.py:7-18).
"""
import numpy as np
import torch

MSRA_IDX2TAG = {0: '[PAD]', 1: 'O', 2: 'B-ORG', 3: 'I-ORG', 4: 'B-PER', 5: 'I-PER', 6: 'B-LOC', 7: 'I-LOC',
                8: '[CLS]', 9: '[SEP]'}
# tag marginals of data/msra/train/tags.txt
_TAG_P = {1: .889, 3: .038, 7: .023, 6: .017, 5: .016, 2: .0095, 4: .0081}

# character-length quantiles of data/msra/train/sentences.txt: mean 48.3, median 41, p90 81, p99 147
_LEN_Q = np.array([0.0, 0.05, 0.25, 0.5, 0.75, 0.9, 0.99, 1.0])
_LEN_V = np.array([4.0, 11.0, 25.0, 41.0, 63.0, 81.0, 147.0, 148.0])


def msra_lengths(B, L, rng, full=False):
    """seq_len including [CLS]/[SEP], clipped to L."""
    if full:
        return np.full(B, L, dtype=np.int32)
    u = rng.random(B)
    chars = np.interp(u, _LEN_Q, _LEN_V)
    return (np.clip(np.round(chars), 1, L - 2) + 2).astype(np.int32)


def msra_batch(B, L, vocab=21128, seed=1234, full=False, label_size=10):
    rng = np.random.default_rng(seed)
    lens = msra_lengths(B, L, rng, full)
    tags = np.array(list(_TAG_P.keys()))
    p = np.array(list(_TAG_P.values()))
    p = p / p.sum()
    token_ids = np.zeros((B, L), np.int32)
    label_ids = np.zeros((B, L), np.int32)
    mask = np.zeros((B, L), np.int32)
    for b in range(B):
        n = int(lens[b])
        token_ids[b, :n] = rng.integers(106, vocab, size=n)
        token_ids[b, 0], token_ids[b, n - 1] = 101, 102
        lab = rng.choice(tags, size=n, p=p)
        lab[0], lab[n - 1] = 8, 9
        label_ids[b, :n] = np.minimum(lab, label_size - 1)
        mask[b, :n] = 1
    return {
        'token_ids': torch.from_numpy(token_ids), 'mask': torch.from_numpy(mask),
        'segment_ids': torch.zeros((B, L), dtype=torch.int32), 'label_ids': torch.from_numpy(label_ids),
        'seq_len': torch.from_numpy(lens),
    }


def softlexicon_features(B, L, n_word, seed=1234, realistic=True, G=4, S=10, lens=None):
    """ids/weights [B, L*G*S]; pad id = n_word-1 (weight 0), none id = n_word-2; per-token weights sum to 1."""
    rng = np.random.default_rng(seed + 7)
    pad_id, none_id = n_word - 1, n_word - 2
    ids = np.full((B, L, G, S), pad_id, np.int32)
    w = np.zeros((B, L, G, S), np.float32)
    for b in range(B):
        n = L if lens is None else int(lens[b])
        for t in range(n):
            if realistic:
                cnt = rng.choice([0, 1, 2, 3], size=G, p=[0.35, 0.4, 0.17, 0.08])
                if cnt.sum() == 0:
                    cnt[3] = 1
            else:
                cnt = np.full(G, S)
            freq = []
            for g in range(G):
                if cnt[g] == 0:
                    ids[b, t, g, 0] = none_id
                else:
                    ids[b, t, g, :cnt[g]] = rng.integers(0, n_word - 2, size=cnt[g])
                    f = rng.integers(1, 1000, size=cnt[g]).astype(np.float32)
                    w[b, t, g, :cnt[g]] = f
                    freq.append(f.sum())
            tot = w[b, t].sum()
            if tot > 0:
                w[b, t] /= tot
    return torch.from_numpy(ids.reshape(B, L * G * S)), torch.from_numpy(w.reshape(B, L * G * S))


def softlexicon_features_device(n_tok, n_word, realistic=True, seed=1234, G=4, S=10, device="cuda"):
    """Vectorised variant for large token counts: ids int32 / weights f32 [n_tok, G*S] on `device`.
    Layout as the reference's postproc_soft_lexicon (data/word_enhance.py:163-205): per group the matched words first, then
    <PAD> (= n_word-1, frequency 0 -> weight 0); an empty group holds <None> (= n_word-2, frequency 1) in slot 0; weights
    = frequency / sum of the token's 40 frequencies.  realistic: 0-3 words per group (p = .35/.4/.17/.08, the slot
    statistics of the serving warm-up record); dense: all 10 slots of every group are words."""
    g = torch.Generator(device=device).manual_seed(seed + 7)
    pad_id, none_id = n_word - 1, n_word - 2
    if realistic:
        p = torch.tensor([0.35, 0.4, 0.17, 0.08], device=device)
        cnt = torch.multinomial(p, n_tok * G, replacement=True, generator=g).view(n_tok, G)
    else:
        cnt = torch.full((n_tok, G), S, device=device, dtype=torch.long)
    slot = torch.arange(S, device=device).view(1, 1, S)
    is_word = slot < cnt.unsqueeze(-1)
    is_none = (cnt.unsqueeze(-1) == 0) & (slot == 0)
    words = torch.randint(0, n_word - 2, (n_tok, G, S), device=device, generator=g, dtype=torch.int32)
    freq = torch.randint(1, 1000, (n_tok, G, S), device=device, generator=g).to(torch.float32)
    ids = torch.where(is_word, words, torch.full_like(words, pad_id))
    ids = torch.where(is_none, torch.full_like(words, none_id), ids)
    w = torch.where(is_word, freq, torch.zeros_like(freq))
    w = torch.where(is_none, torch.ones_like(freq), w)
    w = w / w.sum(dim=(1, 2), keepdim=True)
    return ids.view(n_tok, G * S).contiguous(), w.view(n_tok, G * S).contiguous()


def data_params(L, label_size=10, n_sample=45000, batch_size=32, epoch_size=50):
    step_per_epoch = n_sample // batch_size
    return {'label_size': label_size, 'max_seq_len': L, 'idx2tag': dict(MSRA_IDX2TAG), 'n_sample': n_sample,
            'tag2idx': {v: k for k, v in MSRA_IDX2TAG.items()}, 'step_per_epoch': step_per_epoch,
            'num_train_steps': step_per_epoch * epoch_size}
