"""CPU: a host model of `crf_viterbi_tma_kernel` (chinesener_b200/csrc/crf_viterbi.cu) against the oracle.

The model follows the kernel's control flow and index arithmetic statement by statement — byte-accurate shared-memory
carve-up (ring / decoded tags overlay, backpointer words, high-nibble bytes, lengths, barriers) inside a buffer of exactly
`viterbi_tma_smem_bytes`, TMA boxes with zero fill past the row / column extent, the two-stage ring and its refill rule,
the peeled chunk 0, the check-free path for chunks every live row covers, the max + first-equal-index formulation with the
4-bit backpointer packing, the 8-step prefetched backtrace and the division-free output sweep — with one numpy lane vector
per warp.  It cannot see synchronisation or hardware behaviour (the GPU suite does: tests/test_crf_gpu.py); it pins the
LOGIC of the kernel to oracle/crf.py (tf.contrib.crf.crf_decode, reference tools/layer.py:140-142) on ragged lengths,
partial tail warps, ties, every K parity / high-nibble width, L not a multiple of the chunk or of 4.
"""
import numpy as np
import pytest

from oracle import crf

NT, S = 32, 2           # launch_viterbi_tma<K, 32, 2, 2, 8>: one-warp CTAs, two-stage ring, TT = 2


def _geom(K, TT=2):
    T = 4 if K % 2 else TT
    NQ = T * K // 4
    return T, 4 * (NQ | 1), (K + 1) // 2, (0 if K <= 8 else 1 if K <= 10 else 2 if K <= 12 else 4)


def _smem_bytes(L, K):
    T, PW, KP, HB = _geom(K)
    ring = S * NT * PW * 4
    dec = NT * (((L + 3) & ~3) + 4)
    lo = L * NT * 4
    hi = (L * NT * HB + 15) & ~15
    return max(ring, dec) + lo + hi + NT * 4 + 8 + (NT // 32) * S * 8


def viterbi_tma_model(x, lens, trans):
    """-> (tags [B, L] int32, best [B] float32) as the kernel writes them (tags_out starts poisoned: every element the
    kernel is responsible for must be written)."""
    x = np.ascontiguousarray(x, np.float32)
    trans = np.ascontiguousarray(trans, np.float32)
    B, L, K = x.shape
    LK = L * K
    assert LK % 4 == 0 and K <= 16, "the launcher returns UNSUPPORTED otherwise"
    T, PW, KP, HB = _geom(K)
    assert T * K % 4 == 0
    xflat = x.reshape(B, LK)
    tags_out = np.full(B * L, -77, np.int32)
    best_out = np.full(B, np.nan, np.float32)
    vec_out = (L & 3) == 0
    Lp = ((L + 3) & ~3) + 4
    ring_b, dec_b = S * NT * PW * 4, NT * Lp
    off_lo = max(ring_b, dec_b)
    off_hi = off_lo + L * NT * 4
    off_len = off_hi + ((L * NT * HB + 15) & ~15)
    off_bar = off_len + NT * 4
    total = _smem_bytes(L, K)
    assert off_bar % 8 == 0 and off_bar + S * 8 <= total
    lane = np.arange(32)

    def tma_box(stage_view, c0, r0):          # box = 32 rows x PW floats at (column c0, row r0), zero fill out of range
        box = np.zeros((32, PW), np.float32)
        rows = np.arange(r0, r0 + 32)
        cols = np.arange(c0, c0 + PW)
        rv, cv = rows < B, cols < LK
        box[np.ix_(rv, cv)] = xflat[np.ix_(rows[rv], cols[cv])]
        stage_view[:] = box.reshape(-1).view(np.uint8)

    for blk in range((B + NT - 1) // NT):
        smem = np.zeros(total, np.uint8)
        smem[:] = 0xA5                                            # stale shared memory
        row0 = blk * NT
        nv = min(NT, B - row0)
        wrow0 = row0
        live = lane < nv
        mylen = np.ones(32, np.int64)
        mylen[live] = np.minimum(np.maximum(lens[row0:row0 + nv].astype(np.int64), 1), L)
        smem[off_len:off_len + NT * 4] = mylen.astype(np.int32).view(np.uint8)
        wmax = int(np.max(np.where(live, mylen, 0)))
        nchunk = (wmax + T - 1) // T
        stage = [smem[s * 32 * PW * 4:(s + 1) * 32 * PW * 4] for s in range(S)]
        pending = [None] * S                                      # chunk whose bytes are in flight / landed in the stage
        for s in range(S):
            if s < nchunk:
                tma_box(stage[s], s * T * K, wrow0)
                pending[s] = s
        tr2 = np.zeros((K, 2 * KP), np.float32)                   # tr2[j][i] = trans[i][j], pad column 0
        tr2[:, :K] = trans.T
        s2 = np.zeros((32, 2 * KP), np.float32)
        wnv = max(0, min(32, nv))
        wmin = int(np.min(np.where(live, mylen, L)))
        lo = smem[off_lo:off_lo + L * NT * 4].view(np.uint32).reshape(L, NT)
        hi_bytes = smem[off_hi:off_hi + max(L * NT * HB, 1)]
        hi = hi_bytes[:L * NT * HB].view({1: np.uint8, 2: np.uint16, 4: np.uint32}[HB]).reshape(L, NT) if HB else None

        def dp_step(xs, g, t, who):
            nonlocal s2
            m = np.zeros((32, 2 * KP), np.float32)
            wlo = np.zeros(32, np.uint64)
            whi = np.zeros(32, np.uint64)
            for J in range(K):
                v = s2 + tr2[J][None, :]                          # add2 over the pairs: (s[i] + trans[i][J])
                mj = np.max(v[:, :K], axis=1)
                ix = np.full(32, K - 1, np.uint64)
                for I in range(K - 2, -1, -1):                    # walked downwards: the lowest equal index remains
                    ix = np.where(v[:, I] == mj, np.uint64(I), ix)
                m[:, J] = mj
                if J < 8:
                    wlo |= ix << np.uint64(4 * J)
                else:
                    whi |= ix << np.uint64(4 * (J - 8))
            if 2 * KP > K:
                m[:, 2 * KP - 1] = 0.0
            xg = np.zeros((32, 2 * KP), np.float32)
            xg[:, :K] = xs[:, g * K:(g + 1) * K]
            new = m + xg
            s2 = np.where(who[:, None], new, s2)
            lo[t, who] = wlo[who].astype(np.uint32)
            if HB:
                hi[t, who] = whi[who].astype(hi.dtype)            # the store truncates to HB bytes

        def refill(c, st):
            if c + S < nchunk:
                tma_box(stage[st], (c + S) * T * K, wrow0)
                pending[st] = c + S

        def do_chunk(c, first):
            st = c % S
            assert pending[st] == c, "the stage must hold this chunk (ring / parity logic)"
            t0 = c * T
            rows = stage[st].view(np.float32).reshape(32, PW)
            if not first and t0 + T <= wmin:                      # check-free path: every lane, dead ones included
                xs = rows[:, :T * K].copy()
                refill(c, st)                                     # the stage may be overwritten: xs is a copy
                everyone = np.ones(32, bool)
                for g in range(T):
                    assert t0 + g < L
                    dp_step(xs, g, t0 + g, everyone)
                return
            who0 = live & (t0 < mylen)
            xs = rows[:, :T * K].copy()
            for g in range(T):
                t = t0 + g
                who = who0 & (t < mylen)
                if not who.any():
                    continue
                if first and t == 0:
                    nonlocal_s2_init(xs, who)
                else:
                    dp_step(xs, g, t, who)
            refill(c, st)

        def nonlocal_s2_init(xs, who):
            nonlocal s2
            init = np.zeros((32, 2 * KP), np.float32)
            init[:, :K] = xs[:, :K]
            s2 = np.where(who[:, None], init, s2)

        if nchunk > 0:
            do_chunk(0, True)
        for c in range(1, nchunk):
            do_chunk(c, False)

        dec = smem[:NT * Lp].reshape(NT, Lp)                       # decoded tags over the dead ring
        for ln in range(32):
            if not live[ln]:
                continue
            sv = s2[ln]
            bst, y = sv[0], 0
            for j in range(1, K):
                if sv[j] > bst:
                    bst, y = sv[j], j
            best_out[row0 + ln] = bst
            t = int(mylen[ln]) - 1
            while t >= 1:
                wa, wb = [0] * 8, [0] * 8
                for u in range(8):
                    inside = t - u >= 1
                    wa[u] = int(lo[t - u, ln]) if inside else 0
                    wb[u] = int(hi[t - u, ln]) if (inside and HB) else 0
                for u in range(8):
                    tt = t - u
                    if tt >= 1:
                        dec[ln, tt] = y
                        w = wb[u] if (HB > 0 and y >= 8) else wa[u]
                        y = (w >> (4 * (y & 7))) & 15
                t -= 8
            dec[ln, 0] = y

        wlen = smem[off_len:off_len + NT * 4].view(np.int32)
        obase = wrow0 * L
        if vec_out:
            L4, total4 = L >> 2, wnv * (L >> 2)
            for ln in range(32):
                r, q = 0, ln
                while q >= L4:
                    q -= L4
                    r += 1
                idx = ln
                while idx < total4:
                    p = 4 * q
                    n = int(wlen[r])
                    for e in range(4):
                        tags_out[obase + 4 * idx + e] = int(dec[r, p + e]) if p + e < n else 0
                    q += 32
                    while q >= L4:
                        q -= L4
                        r += 1
                    idx += 32
        else:
            for idx in range(wnv * L):
                r, p = idx // L, idx % L
                tags_out[obase + idx] = int(dec[r, p]) if p < wlen[r] else 0
    return tags_out.reshape(B, L), best_out


def _case(B, L, K, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, L, K)).astype(np.float32)
    tr = (rng.standard_normal((K, K)) * 0.5).astype(np.float32)
    lens = rng.integers(1, L + 1, size=B).astype(np.int32)
    return x, tr, lens


@pytest.mark.parametrize("B,L,K", [(77, 128, 10), (96, 50, 16), (65, 21, 12), (40, 12, 5), (70, 16, 9), (64, 20, 8), (33, 8, 11),
                                   (45, 256, 10), (100, 6, 2), (37, 4, 13), (50, 10, 14), (31, 2, 10), (5, 1, 4), (64, 3, 4)])
def test_kernel_model_matches_the_oracle_on_ragged_batches(B, L, K):
    x, tr, lens = _case(B, L, K, seed=100 + K + L)
    lens[0] = L
    if B > 2:
        lens[1], lens[2] = 1, 0
    if B > 5:
        x[5] = np.round(x[5])                                      # a row with many exact ties
    ref_tags, ref_best = crf.crf_decode(x, tr, lens, dtype=np.float32)
    tags, best = viterbi_tma_model(x, lens, tr)
    np.testing.assert_array_equal(tags, ref_tags)
    np.testing.assert_array_equal(best, ref_best.astype(np.float32))


@pytest.mark.parametrize("B,L,K", [(70, 128, 10), (64, 24, 10), (40, 16, 9), (33, 30, 12)])
def test_kernel_model_full_lengths_take_the_check_free_path(B, L, K):
    """Every row at full length: all chunks but the peeled first one (and a trailing partial chunk) run without per-row checks,
    the dead lanes of the tail warp included — the roofline configuration."""
    x, tr, _ = _case(B, L, K, seed=7 + K)
    lens = np.full(B, L, np.int32)
    ref_tags, ref_best = crf.crf_decode(x, tr, lens, dtype=np.float32)
    tags, best = viterbi_tma_model(x, lens, tr)
    np.testing.assert_array_equal(tags, ref_tags)
    np.testing.assert_array_equal(best, ref_best.astype(np.float32))


def test_kernel_model_warp_with_a_short_row_and_ties_everywhere():
    rng = np.random.default_rng(3)
    B, L, K = 64, 32, 10
    x = rng.integers(-1, 2, size=(B, L, K)).astype(np.float32)
    tr = rng.integers(-1, 2, size=(K, K)).astype(np.float32)
    lens = np.full(B, L, np.int32)
    lens[40] = 7                                                   # second warp: wmin = 7, chunks beyond it take the checked path
    ref_tags, ref_best = crf.crf_decode(x, tr, lens, dtype=np.float32)
    tags, best = viterbi_tma_model(x, lens, tr)
    np.testing.assert_array_equal(tags, ref_tags)
    np.testing.assert_array_equal(best, ref_best)
