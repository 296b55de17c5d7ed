"""Round-2 golden fixtures, again produced by RUNNING THE REFERENCE ITSELF (never copied; /root/reference is only read here).

    python tests/golden/make_golden_r02.py        # needs /root/reference (build container only)

* dnn_encoder.npz  -- ssl4rec.py:162-196: the reference's own `DNNEncoder` (2 layers), its parameters, `forward(u, i)`,
  `cal_cl_loss(i)` in eval mode (nn.Dropout is the identity there: CPU and CUDA generators produce different masks, SURVEY
  8c), the batch_softmax rec loss of the training loop (ssl4rec.py:221-224) and every parameter gradient of the total loss.
* eval_cold.npz    -- ncl.py:253-277 (`test`) + ncl.py:133-178 (`ranking_evaluation`) on the reference's own `Interaction`
  built from a train / test split whose test set contains items that never occur in training: the reference keeps them
  in `test_set`, so they count in the Hit-Ratio / Recall denominators and in the ideal DCG although they cannot be hit.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import OUT, grads_of, load_ref, t2n, tiny_dataset  # noqa: E402


def dnn_encoder():
    torch.manual_seed(3)
    ssl4rec = load_ref("ssl4rec", "ssl4rec.py")
    train, test = tiny_dataset(seed=11)
    data = ssl4rec.Interaction([list(r) for r in train], [list(r) for r in test])
    model = ssl4rec.DNNEncoder(data, 16, 0.1, 0.2, 2)
    model.eval()
    rng = np.random.default_rng(4)
    B = 24
    u = rng.integers(0, data.user_num, B)
    i = rng.integers(0, data.item_num, B)
    q, k = model(torch.as_tensor(u), torch.as_tensor(i))            # forward (ssl4rec.py:189-190)
    cl = model.cal_cl_loss(i.tolist())                              # ssl4rec.py:192-196, dropout off
    rec = ssl4rec.batch_softmax_loss(q, k, 0.2)                     # ssl4rec.py:221
    total = rec + ssl4rec.l2_reg_loss(1e-4, q, k) + 0.1 * cl        # ssl4rec.py:224 (cl_rate = 0.1)
    params = dict(model.named_parameters())
    grads = grads_of(total, *params.values())
    out = {f"param.{n}": t2n(p) for n, p in params.items()}
    out.update({f"grad.{n}": g for n, g in zip(params, grads)})
    out.update(u=u.astype(np.int64), i=i.astype(np.int64), q=t2n(q), k=t2n(k), cl=float(cl), rec=float(rec), total=float(total),
               n_users=data.user_num, n_items=data.item_num, emb_size=16, n_layers=2, tau=0.2, drop_rate=0.1)
    np.savez(OUT / "dnn_encoder.npz", **out)


def eval_cold():
    ncl = load_ref("ncl", "ncl.py", stubs=("faiss",))
    train, test = tiny_dataset(seed=23, n_users=31, n_items=47, n_inter=300, n_dups=5)
    train_items = {r[1] for r in train}
    train_users = sorted({r[0] for r in train})
    test = [r for r in test if r[0] in train_users]                  # the reference's predict() needs a trained user
    test += [[train_users[0], "i900", 1.0], [train_users[3], "i901", 1.0], [train_users[3], "i902", 1.0]]   # cold items
    assert all(r[1] not in train_items for r in test[-3:])
    data = ncl.Interaction({}, [tuple(r) for r in train], [tuple(r) for r in test])
    rng = np.random.default_rng(9)
    d = 8
    user_emb = rng.standard_normal((data.user_num, d)).astype(np.float32)
    item_emb = rng.standard_normal((data.item_num, d)).astype(np.float32)
    top_ns, max_n = [5, 10], 10
    rec_list = {}
    for user in data.test_set:                                       # the body of test(), ncl.py:255-263
        candidates = (torch.from_numpy(user_emb[data.get_user_id(user)]) @ torch.from_numpy(item_emb).T).numpy().copy()
        rated_list, _ = data.user_rated(user)
        for item in rated_list:
            candidates[data.item[item]] = -1e8
        top = torch.topk(torch.tensor(candidates), max_n)
        rec_list[user] = list(zip([data.id2item[j] for j in top.indices.tolist()], top.values.tolist()))
    strings = ncl.ranking_evaluation(data.test_set, rec_list, top_ns)
    np.savez(OUT / "eval_cold.npz", train_users=np.array([r[0] for r in train]), train_items=np.array([r[1] for r in train]),
             test_users=np.array([r[0] for r in test]), test_items=np.array([r[1] for r in test]),
             user_emb=user_emb, item_emb=item_emb, top_ns=np.array(top_ns), strings=np.array(strings))


if __name__ == "__main__":
    dnn_encoder()
    eval_cold()
    print("r02 golden fixtures written to", OUT)
