"""Golden fixture for univariate/sept_social.py: RUNS THE REFERENCE's own SEPT class (social variant) on a fixed-seed data set
and stores what one training iteration computes (sept_social.py:361-420, the loop body of 431-461) as
tests/golden/sept_social.npz.

    python tests/golden/make_golden_sept_social.py     # needs /root/reference (build container only)

The reference's augmented branch calls `self.data.convert_to_laplacian_mat`, which its Interaction class does not define
(sept_social.py:427 raises AttributeError for epoch > maxEpoch // 3), so the iteration is replayed with aug_mat = norm_adj --
the value the reference itself uses in the epochs that run -- and the tri-training methods are called exactly as lines 445-456 do.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import OUT, grads_of, load_ref, t2n  # noqa: E402


def main():
    torch.set_num_threads(1)
    rng = np.random.default_rng(41)
    n_users, n_items = 70, 55
    u = rng.integers(0, n_users, 900)
    i = (rng.zipf(1.4, 900) - 1) % n_items
    pairs = list(zip(u.tolist(), i.tolist()))
    pairs += [pairs[k] for k in rng.integers(0, len(pairs), 25)]          # duplicate interactions stay duplicates in norm_adj
    train = [(f"u{a:03d}", f"i{b:03d}", 1.0) for a, b in pairs]
    test = [(f"u{a:03d}", f"i{b:03d}", 1.0) for a, b in zip(rng.integers(0, n_users, 20).tolist(), rng.integers(0, n_items, 20).tolist())]
    users = sorted({r[0] for r in train})
    social = set()
    while len(social) < 520:
        a, b = rng.choice(len(users), 2, replace=False)
        social.add((users[a], users[b]))
    social = [[a, b, 1.0] for a, b in sorted(social)]
    mod = load_ref("sept_social", "univariate/sept_social.py", stubs=("tensorflow",))
    conf = {"model": {"name": "SEPT"}, "SEPT": {"n_layer": 2, "ss_rate": 0.005, "drop_rate": 0.3, "ins_cnt": 5},
            "emb_size": 16, "batch_size": 64, "lr": 0.001, "reg_lambda": 1e-4, "max.epoch": 3}
    torch.manual_seed(9)
    m = mod.SEPT(conf, train, test, social)
    m.build()
    csr = lambda M: dict(indptr=M.indptr.astype(np.int64), indices=M.indices.astype(np.int64), data=M.data.astype(np.float32),
                         shape=np.array(M.shape))
    bi = m.social_data.get_birectional_social_mat().tocsr()
    soc, sha = [x.tocsr() for x in m.get_social_related_views(m.bi_social_mat, m.data.interaction_mat)]
    for M in (bi, soc, sha):
        M.sort_indices()
    Y = m.data.interaction_mat.tocsr(); Y.sort_indices()
    adj = m.data.norm_adj.tocoo()

    user_idx = torch.tensor(rng.integers(0, m.data.user_num, 48))
    pos_idx = torch.tensor(rng.integers(0, m.data.item_num, 48))
    neg_idx = torch.tensor(rng.integers(0, m.data.item_num, 48))
    aug_mat = m.norm_adj
    ego = torch.cat([m.user_embeddings, m.item_embeddings], dim=0)
    m.rec_user_embeddings, m.rec_item_embeddings = m.encoder(ego, m.norm_adj, m.n_layers)
    m.aug_user_embeddings, m.aug_item_embeddings = m.encoder(ego, aug_mat, m.n_layers)
    m.sharing_view_embeddings = m.social_encoder(m.user_embeddings, m.sharing_mat, m.n_layers)
    m.friend_view_embeddings = m.social_encoder(m.user_embeddings, m.social_mat, m.n_layers)
    bu, bp, bn = m.rec_user_embeddings[user_idx], m.rec_item_embeddings[pos_idx], m.rec_item_embeddings[neg_idx]
    rec_loss = mod.bpr_loss(bu, bp, bn)
    rec_loss = rec_loss + m.reg * (m.user_embeddings.norm(2).pow(2) + m.item_embeddings.norm(2).pow(2))
    social_prediction = m.label_prediction(m.friend_view_embeddings, user_idx)
    sharing_prediction = m.label_prediction(m.sharing_view_embeddings, user_idx)
    rec_prediction = m.label_prediction(m.rec_user_embeddings, user_idx)
    f_pos = m.generate_pesudo_labels(sharing_prediction, rec_prediction)
    sh_pos = m.generate_pesudo_labels(social_prediction, rec_prediction)
    r_pos = m.generate_pesudo_labels(social_prediction, sharing_prediction)
    nd_f = m.neighbor_discrimination(f_pos, m.friend_view_embeddings, user_idx)
    nd_s = m.neighbor_discrimination(sh_pos, m.sharing_view_embeddings, user_idx)
    nd_r = m.neighbor_discrimination(r_pos, m.rec_user_embeddings, user_idx)
    total = rec_loss + m.ss_rate * (nd_f + nd_s + nd_r)
    g_u, g_i = grads_of(total, m.user_embeddings, m.item_embeddings)
    # margin between the K-th and (K+1)-th mean probability: tells the parity test which label rows are numerically unambiguous
    def margin(p1, p2, k):
        s = torch.sort((p1 + p2) / 2, dim=1, descending=True).values
        return t2n(s[:, k - 1] - s[:, k])
    K = m.instance_cnt
    np.savez(OUT / "sept_social.npz", n_layers=m.n_layers, ss_rate=m.ss_rate, ins_cnt=K, reg=m.reg,
             user_num=m.data.user_num, item_num=m.data.item_num,
             **{f"{n}_{k}": v for n, M in (("bi", bi), ("Y", Y), ("social", soc), ("sharing", sha)) for k, v in csr(M).items()},
             adj_row=adj.row.astype(np.int64), adj_col=adj.col.astype(np.int64), adj_data=adj.data.astype(np.float32),
             user_w=t2n(m.user_embeddings), item_w=t2n(m.item_embeddings),
             user_idx=t2n(user_idx), pos_idx=t2n(pos_idx), neg_idx=t2n(neg_idx),
             rec_user=t2n(m.rec_user_embeddings), rec_item=t2n(m.rec_item_embeddings),
             sharing_view=t2n(m.sharing_view_embeddings), friend_view=t2n(m.friend_view_embeddings),
             social_prediction=t2n(social_prediction), sharing_prediction=t2n(sharing_prediction), rec_prediction=t2n(rec_prediction),
             f_pos=t2n(f_pos), sh_pos=t2n(sh_pos), r_pos=t2n(r_pos),
             f_margin=margin(sharing_prediction, rec_prediction, K), sh_margin=margin(social_prediction, rec_prediction, K),
             r_margin=margin(social_prediction, sharing_prediction, K),
             rec_loss=t2n(rec_loss), nd_f=t2n(nd_f), nd_s=t2n(nd_s), nd_r=t2n(nd_r), total=t2n(total), g_user=g_u, g_item=g_i)
    print("users", m.data.user_num, "items", m.data.item_num, "bi nnz", bi.nnz, "social view nnz", soc.nnz, "sharing nnz", sha.nnz,
          "losses", float(rec_loss), float(nd_f), float(nd_s), float(nd_r))


if __name__ == "__main__":
    main()
