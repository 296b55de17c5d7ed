"""Oracle: Philox4x32-10 (Salmon et al., SC'11; Random123) and the negative-sampling stream defined in
recommendation_b200/csrc/sampler.cu.  numpy uint32/uint64 arithmetic, bit-exact.  Test infrastructure only.

The reference samplers (ncl.py:91-114, selfcf.py:188-211, directau.py:14-32, lightgcn.py:91-94) use unseeded
Python / numpy / torch RNGs, so only "uniform over items, training positives rejected" is reference behaviour.
Known-answer vectors for the generator itself are Random123's kat_vectors (checked in tests/test_oracle_golden.py).
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: 4 arrays (uint32) broadcastable; key: 2 arrays (uint32).  Returns 4 uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint32).copy() for x in counter]
    c = list(np.broadcast_arrays(*c))
    k0 = np.broadcast_to(np.asarray(key[0], dtype=np.uint32), c[0].shape).copy()
    k1 = np.broadcast_to(np.asarray(key[1], dtype=np.uint32), c[0].shape).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK32).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return c


def sample_negatives(seed, offset, n, n_negs, n_items, users=None, pos_row_ptr=None, pos_col_idx=None, max_trials=1,
                     slot_base=0):
    """int64 [n * n_negs]; stream definition in csrc/sampler.cu's header comment.  slot_base: the window
    [slot_base, slot_base + n * n_negs) of the global slot numbering (gcf_sample_negatives_at)."""
    seed, offset = int(seed) & (2**64 - 1), int(offset) & (2**64 - 1)
    k0 = np.uint32(seed & 0xFFFFFFFF)
    k1 = np.uint32(((seed >> 32) ^ (offset >> 32)) & 0xFFFFFFFF)
    off_lo = np.uint32(offset & 0xFFFFFFFF)
    slots = np.arange(n * n_negs, dtype=np.uint64) + np.uint64(slot_base)
    s_lo, s_hi = (slots & MASK32).astype(np.uint32), (slots >> np.uint64(32)).astype(np.uint32)
    reject = pos_row_ptr is not None
    if not reject:
        max_trials = 1
    out = np.zeros(n * n_negs, dtype=np.int64)
    pending = np.ones(n * n_negs, dtype=bool)
    words = None
    for trial in range(max(1, max_trials)):
        if trial % 4 == 0:
            words = philox4x32_10((s_lo, s_hi, off_lo, np.uint32(trial // 4)), (k0, k1))
        cand = ((words[trial % 4].astype(np.uint64) * np.uint64(n_items)) >> np.uint64(32)).astype(np.int64)
        out[pending] = cand[pending]
        if not reject:
            break
        u = np.asarray(users, dtype=np.int64)[np.arange(n * n_negs) // n_negs]
        is_pos = np.zeros(n * n_negs, dtype=bool)
        for s in np.nonzero(pending)[0]:
            lo, hi = pos_row_ptr[u[s]], pos_row_ptr[u[s] + 1]
            seg = pos_col_idx[lo:hi]
            j = np.searchsorted(seg, cand[s])
            is_pos[s] = j < seg.shape[0] and seg[j] == cand[s]
        pending = pending & is_pos
        if not pending.any():
            break
    return out
