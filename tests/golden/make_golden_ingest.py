"""Golden fixture for the text ingest path: a small `user item rating` train / test file pair is written, read back with THE
REFERENCE's own `load_data` and `Interaction` classes (ncl.py:46-90,542-543: ids numbered by sorted() of the id strings;
selfcf.py:258-310: by first appearance) and the file bytes plus the reference's dictionaries, index arrays and adjacency
matrices are stored as tests/golden/ingest.npz.  lightgcn.py:29-33 (integer ids, pandas) is replayed with the same pandas call.

    python tests/golden/make_golden_ingest.py     # needs /root/reference (build container only)
"""
from __future__ import annotations

import sys
import tempfile
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import OUT, load_ref  # noqa: E402


def main():
    rng = np.random.default_rng(77)
    # ids of different lengths so that the string order differs from the numeric one ("10" < "2", "u9" > "u10")
    users = [str(k) for k in rng.choice(3000, 60, replace=False)] + [f"u{k}" for k in range(12)]
    items = [str(k) for k in rng.choice(500, 45, replace=False)] + [f"it{k:02d}" for k in range(8)]
    lines = []
    for _ in range(700):
        u, i = users[rng.integers(len(users))], items[int((rng.zipf(1.5) - 1) % len(items))]
        sep = [" ", "\t", "  "][rng.integers(3)]
        lines.append(f"{u}{sep}{i}{sep}{rng.integers(1, 6)}")
    lines += [lines[k] for k in rng.integers(0, len(lines), 30)]            # duplicate interactions
    for pos in rng.integers(0, len(lines), 6):                              # blank / whitespace-only lines are skipped
        lines.insert(int(pos), ["", "   ", "\t"][rng.integers(3)])
    train_txt = "\n".join(lines) + "\n\n"
    test_lines = [f"{users[rng.integers(len(users))]} {items[rng.integers(len(items))]} 1" for _ in range(60)]
    test_lines += ["zz9 7 1", f"{users[0]} nope 1"]                          # ids the training file does not contain
    test_txt = "\n".join(test_lines)                                        # no trailing newline
    ncl = load_ref("ncl", "ncl.py", stubs=("faiss",))
    selfcf = load_ref("selfcf", "selfcf.py")
    with tempfile.TemporaryDirectory() as td:
        tr, te = Path(td) / "train.txt", Path(td) / "test.txt"
        tr.write_text(train_txt); te.write_text(test_txt)
        train, test = ncl.load_data(str(tr)), ncl.load_data(str(te))
        d1 = ncl.Interaction({}, train, test)
        d2 = selfcf.Interaction({}, [list(r) for r in train], [list(r) for r in test])
    out = dict(train_bytes=np.frombuffer(train_txt.encode(), dtype=np.uint8), test_bytes=np.frombuffer(test_txt.encode(), dtype=np.uint8),
               n_records=len(train))
    for tag, d in (("sorted", d1), ("appearance", d2)):
        adj = d.norm_adj.tocoo() if tag == "sorted" else d.norm_adj.tocsr()
        out.update({f"{tag}_user_ids": np.array([d.id2user[k] for k in range(d.user_num)]),
                    f"{tag}_item_ids": np.array([d.id2item[k] for k in range(d.item_num)]),
                    f"{tag}_users": np.array([d.user[r[0]] for r in train], dtype=np.int64),
                    f"{tag}_items": np.array([d.item[r[1]] for r in train], dtype=np.int64),
                    f"{tag}_test_users": np.array([d.user.get(r[0], -1) for r in test], dtype=np.int64),
                    f"{tag}_test_items": np.array([d.item.get(r[1], -1) for r in test], dtype=np.int64)})
        if tag == "sorted":      # raw COO, duplicates kept (ncl.py:76-85)
            out.update(sorted_adj_row=adj.row.astype(np.int64), sorted_adj_col=adj.col.astype(np.int64), sorted_adj_data=adj.data.astype(np.float32))
        else:                    # normalised CSR (selfcf.py:240-255,297-306)
            adj.sort_indices()
            out.update(appearance_adj_indptr=adj.indptr.astype(np.int64), appearance_adj_indices=adj.indices.astype(np.int64),
                       appearance_adj_data=adj.data.astype(np.float32))
    # integer ids (lightgcn.py:29-33): same pandas call on an all-integer file
    import pandas as pd
    num_lines = [f"{rng.integers(0, 90)} {rng.integers(0, 70)} 1" for _ in range(300)]
    num_test = [f"{rng.integers(0, 95)} {rng.integers(0, 60)} 1" for _ in range(40)]
    with tempfile.TemporaryDirectory() as td:
        tr, te = Path(td) / "train.txt", Path(td) / "test.txt"
        tr.write_text("\n".join(num_lines) + "\n"); te.write_text("\n".join(num_test) + "\n")
        a = pd.read_csv(tr, sep=" ", names=["user", "item", "rating"]); b = pd.read_csv(te, sep=" ", names=["user", "item", "rating"])
        out.update(num_train_bytes=np.frombuffer(tr.read_bytes(), dtype=np.uint8), num_test_bytes=np.frombuffer(te.read_bytes(), dtype=np.uint8),
                   num_users=a["user"].values.astype(np.int64), num_items=a["item"].values.astype(np.int64),
                   num_user_count=int(max(a["user"].max(), b["user"].max()) + 1), num_item_count=int(max(a["item"].max(), b["item"].max()) + 1))
    np.savez(OUT / "ingest.npz", **out)
    print("records", len(train), "users", d1.user_num, d2.user_num, "items", d1.item_num, "first sorted ids", list(out["sorted_user_ids"][:6]),
          "first appearance ids", list(out["appearance_user_ids"][:4]))


if __name__ == "__main__":
    main()
