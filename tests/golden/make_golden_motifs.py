"""Golden fixture for the motif-induced adjacency matrices: RUNS THE REFERENCE's own MHCN.build_hyper_adj_mats
(univariate/mhcn.py:340-368, scipy) on a fixed-seed social + interaction graph and stores its inputs (S, Y) and outputs
(H_s, H_j, H_p) as tests/golden/mhcn_motifs.npz.

    python tests/golden/make_golden_motifs.py     # needs /root/reference (build container only)
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import OUT, load_ref  # noqa: E402


def main():
    rng = np.random.default_rng(23)
    n_users, n_items = 90, 70
    # interactions: Zipf-popular items so that co-purchase counts above the reference's H_p threshold (> 3) exist
    u = rng.integers(0, n_users, 1400)
    i = (rng.zipf(1.3, 1400) - 1) % n_items
    pairs = sorted(set(zip(u.tolist(), i.tolist())))
    train = [(f"u{a:03d}", f"i{b:03d}", 1.0) for a, b in pairs]
    test = [(f"u{a:03d}", f"i{b:03d}", 1.0) for a, b in zip(rng.integers(0, n_users, 30).tolist(), rng.integers(0, n_items, 30).tolist())]
    users = sorted({r[0] for r in train})
    social = set()
    while len(social) < 700:                       # directed follow edges, no self loops
        a, b = rng.choice(len(users), 2, replace=False)
        social.add((users[a], users[b]))
    social = sorted(social)
    for k in range(0, 400, 2):                     # reciprocal edges: bidirectional motifs must not be empty
        social.append((social[k][1], social[k][0]))
    social = [[a, b, 1.0] for a, b in sorted(set(social))]
    mh = load_ref("mhcn", "univariate/mhcn.py", stubs=("tensorflow",))
    conf = {"model": {"name": "MHCN"}, "MHCN": {"n_layer": 2, "ss_rate": 0.01}, "emb_size": 8, "batch_size": 64, "lr": 0.001,
            "reg_lambda": 1e-4, "max.epoch": 1}
    torch.manual_seed(5)
    m = mh.MHCN(conf, train, test, social)
    S = m.social_data.get_social_mat().tocsr()
    Y = m.data.interaction_mat.tocsr()
    Hs, Hj, Hp = [x.tocsr() for x in m.build_hyper_adj_mats()]
    for M in (S, Y, Hs, Hj, Hp):
        M.sort_indices()
    csr = lambda M: dict(indptr=M.indptr.astype(np.int64), indices=M.indices.astype(np.int64), data=M.data.astype(np.float32),
                         shape=np.array(M.shape))
    np.savez(OUT / "mhcn_motifs.npz", **{f"{n}_{k}": v for n, M in (("S", S), ("Y", Y), ("Hs", Hs), ("Hj", Hj), ("Hp", Hp))
                                         for k, v in csr(M).items()})
    print("S", S.shape, S.nnz, "Y", Y.shape, Y.nnz, "Hs", Hs.nnz, "Hj", Hj.nnz, "Hp", Hp.nnz)


if __name__ == "__main__":
    main()
