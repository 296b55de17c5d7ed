"""Multi-rank path.  CPU: world_size-2 gloo run of the sharding choreography (ShardPlan index arithmetic + real
collectives, with the oracle's numpy SpMM standing in for the CUDA kernel).  GPU: 2-rank NCCL run of
ShardedLightGCNTrainer against the single-GPU trainer (needs >= 2 devices; skipped otherwise)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from recommendation_b200 import synth
from recommendation_b200.dist import ShardPlan, shard_table, unshard_table


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_plan_arithmetic():
    plan = ShardPlan(10, 7, 4)
    v = torch.arange(plan.n_nodes)
    pos = plan.gathered_pos(v)
    assert plan.n_loc == 5 and plan.n_padded == 20
    assert len(set(pos.tolist())) == plan.n_nodes and pos.max() < plan.n_padded
    assert torch.equal(plan.node_of_pos(pos), v)
    for r in range(4):
        nodes = plan.local_nodes(r)
        assert torch.equal(nodes % 4, torch.full_like(nodes, r))
        assert torch.equal(pos[nodes], r * plan.n_loc + torch.arange(nodes.numel()))
    spans = [plan.triple_range(23, r) for r in range(4)]
    assert spans[0][0] == 0 and spans[-1][1] == 23 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    full = torch.randn(plan.n_nodes, 3)
    gathered = torch.cat([shard_table(full, plan, r) for r in range(4)])
    assert torch.equal(unshard_table(gathered, plan), full)


def _gloo_worker(rank, world, port, n_users, n_items, users, items, x0, k, out_q):
    from oracle import graph_ref

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = ShardPlan(n_users, n_items, world)
        users_t, items_t = torch.from_numpy(users), torch.from_numpy(items)
        rows, cols = plan.local_block_coo(users_t, items_t, rank)
        deg = np.bincount(np.concatenate([users, items + n_users]), minlength=plan.n_nodes).astype(np.float32)
        with np.errstate(divide="ignore"):
            dinv = 1.0 / np.sqrt(deg)
        dinv[np.isinf(dinv)] = 0
        dinv_pos = np.zeros(plan.n_padded, np.float32)
        dinv_pos[plan.gathered_pos(torch.arange(plan.n_nodes)).numpy()] = dinv
        rp, ci, mult = graph_ref.coo_to_canonical_csr(rows.numpy(), cols.numpy(), None, plan.n_loc, plan.n_padded)
        row_of = np.repeat(np.arange(plan.n_loc), np.diff(rp))
        vals = (dinv_pos[rank * plan.n_loc + row_of] * mult) * dinv_pos[ci]
        local = shard_table(torch.from_numpy(x0), plan, rank)
        layers = [local]
        for _ in range(k):
            full = torch.zeros(plan.n_padded, x0.shape[1], dtype=torch.float64)
            dist.all_gather_into_tensor(full, layers[-1].double().contiguous())   # one all-gather per layer
            layers.append(torch.from_numpy(graph_ref.spmm_csr(rp, ci, vals, full.numpy())))
        final = torch.stack([l.double() for l in layers]).sum(0)
        gathered = torch.zeros(plan.n_padded, x0.shape[1], dtype=torch.float64)
        dist.all_gather_into_tensor(gathered, final.contiguous())
        if rank == 0:
            out_q.put(unshard_table(gathered, plan).numpy())
    finally:
        dist.destroy_process_group()


def test_sharded_propagation_gloo_world2():
    from oracle import graph_ref

    inter = synth.power_law_bipartite(60, 90, 700, seed=4)
    U, I, k, d = 60, 90, 3, 8
    x0 = np.random.default_rng(0).standard_normal((U + I, d)).astype(np.float32)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, U, I, inter.users, inter.items, x0, k, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ei = graph_ref.bipartite_edge_index(inter.users, inter.items, U)
    rp, ci, mult = graph_ref.coo_to_canonical_csr(ei[0], ei[1], None, U + I, U + I)
    vals, _, _ = graph_ref.normalize_csr(rp, ci, mult, U + I, U + I, "sym")
    _, want = graph_ref.propagate(rp, ci, vals, x0, k, "sum")
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------- GPU, 2 ranks
def _tables_close(got, want, lr=0.01):
    """Adam divides by sqrt(v): a gradient element at fp32-noise level can flip sign between two summation orders and
    move that one parameter by up to ~2*lr*steps.  Require the tight bound on (almost) every element and the Adam
    step bound on all of them."""
    err = np.abs(got - want)
    tight = err <= 2e-5 + 1e-3 * np.abs(want)
    assert tight.mean() > 0.9999, f"{(~tight).sum()} of {tight.size} table entries outside rtol 1e-3"
    assert err.max() <= 6.5 * lr, f"max table deviation {err.max()}"


def _nccl_worker(rank, world, port, n_users, n_items, users, items, table0, negs, k, out_q, feature_shards=1):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from recommendation_b200.dist import ShardedLightGCNTrainer

        users_t, items_t = torch.from_numpy(users).to(dev), torch.from_numpy(items).to(dev)
        tr = ShardedLightGCNTrainer(users_t, items_t, n_users, n_items, d=table0.shape[1], n_layers=k, lr=0.01,
                                    reg_weight=1e-4, init_table=torch.from_numpy(table0), feature_shards=feature_shards)
        losses = []
        for s in range(negs.shape[0]):
            losses.append(float(tr.step(neg_items=torch.from_numpy(negs[s]).to(dev)).item()))
        table = tr.gathered_table().cpu().numpy()
        if rank == 0:
            out_q.put((losses, table))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_trainer_matches_single_gpu():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    from recommendation_b200.graph import CSRGraph
    from recommendation_b200.lightgcn import FusedLightGCNTrainer

    inter = synth.power_law_bipartite(3000, 4000, 100000, seed=6)
    U, I, d, k, steps = 3000, 4000, 64, 3, 3
    rng = np.random.default_rng(1)
    table0 = (rng.standard_normal((U + I, d)) * 0.05).astype(np.float32)
    negs = rng.integers(0, I, (steps, inter.n_edges))
    dev = torch.device("cuda", 0)
    users_t, items_t = torch.from_numpy(inter.users).to(dev), torch.from_numpy(inter.items).to(dev)
    g = CSRGraph.from_pairs(users_t, items_t, U, I, norm="sym")
    ref = FusedLightGCNTrainer(g, U, I, torch.from_numpy(table0).to(dev), users_t, items_t, n_layers=k, lr=0.01, reg_weight=1e-4)
    want_losses = [float(ref.step(neg_i=torch.from_numpy(negs[s]).to(dev)).item()) for s in range(steps)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, U, I, inter.users, inter.items, table0, negs, k, q)) for r in range(2)]
    for p in procs:
        p.start()
    got_losses, got_table = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    np.testing.assert_allclose(got_losses, want_losses, rtol=1e-4)
    _tables_close(got_table, ref.table.cpu().numpy())


@pytest.mark.gpu
def test_2d_sharded_trainer_matches_single_gpu():
    """rows x features layout: 2 row groups x (2 | 4) feature groups."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 4:
        pytest.skip("needs >= 4 CUDA devices")
    from recommendation_b200.graph import CSRGraph
    from recommendation_b200.lightgcn import FusedLightGCNTrainer

    world = 8 if torch.cuda.device_count() >= 8 else 4
    inter = synth.power_law_bipartite(3000, 4000, 100000, seed=7)
    U, I, d, k, steps = 3000, 4000, 64, 3, 3
    rng = np.random.default_rng(2)
    table0 = (rng.standard_normal((U + I, d)) * 0.05).astype(np.float32)
    negs = rng.integers(0, I, (steps, inter.n_edges))
    dev = torch.device("cuda", 0)
    users_t, items_t = torch.from_numpy(inter.users).to(dev), torch.from_numpy(inter.items).to(dev)
    g = CSRGraph.from_pairs(users_t, items_t, U, I, norm="sym")
    ref = FusedLightGCNTrainer(g, U, I, torch.from_numpy(table0).to(dev), users_t, items_t, n_layers=k, lr=0.01, reg_weight=1e-4)
    want_losses = [float(ref.step(neg_i=torch.from_numpy(negs[s]).to(dev)).item()) for s in range(steps)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, U, I, inter.users, inter.items, table0, negs, k, q, world // 2))
             for r in range(world)]
    for p in procs:
        p.start()
    got_losses, got_table = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    np.testing.assert_allclose(got_losses, want_losses, rtol=1e-4)
    _tables_close(got_table, ref.table.cpu().numpy())


# ---------------------------------------------------------------------------------------------- feature sharding
def _feature_gloo_worker(rank, world, port, n_users, users, items, negs, x, out_q):
    """Host-side identity of the feature-sharded layout, on gloo: partial scores over column slices, one all-reduce
    of E floats, then the pointwise loss -- compared with the unsharded oracle by the parent."""
    from recommendation_b200.dist import feature_slice

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = feature_slice(x.shape[1], world, rank)
        xs = torch.from_numpy(x[:, lo:hi]).double()
        u, p, n = xs[torch.from_numpy(users)], xs[n_users + torch.from_numpy(items)], xs[n_users + torch.from_numpy(negs)]
        part = (u * (p - n)).sum(1)
        reg = (u.pow(2).sum() + p.pow(2).sum()).reshape(1)
        dist.all_reduce(part); dist.all_reduce(reg)
        loss = torch.nn.functional.softplus(-part).mean() + 1e-3 * reg[0]
        if rank == 0:
            out_q.put(float(loss))
    finally:
        dist.destroy_process_group()


def test_feature_slices_and_score_allreduce_gloo_world2():
    from oracle import losses_ref
    from recommendation_b200.dist import feature_slice

    assert [feature_slice(64, 8, r) for r in (0, 7)] == [(0, 8), (56, 64)]
    assert feature_slice(64, 1, 0) == (0, 64)
    with pytest.raises(ValueError):
        feature_slice(64, 3, 0)
    with pytest.raises(ValueError):
        feature_slice(16, 8, 0)          # 2 floats per rank: below the 128-bit access width
    rng = np.random.default_rng(3)
    U, I, E, d = 50, 70, 900, 16
    users, items, negs = rng.integers(0, U, E), rng.integers(0, I, E), rng.integers(0, I, E)
    x = (rng.standard_normal((U + I, d)) * 0.3).astype(np.float32)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_feature_gloo_worker, args=(r, 2, port, U, users, items, negs, x, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    xt = torch.from_numpy(x).double()
    want = losses_ref.bpr_lightgcn(xt[:U], xt[U:], torch.from_numpy(users), torch.from_numpy(items), torch.from_numpy(negs), 1e-3)
    np.testing.assert_allclose(got, float(want), rtol=1e-10)


def _nccl_feature_worker(rank, world, port, n_users, n_items, users, items, table0, negs, k, out_q, loss_layout="rows"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from recommendation_b200.dist import FeatureShardedLightGCNTrainer

        users_t, items_t = torch.from_numpy(users).to(dev), torch.from_numpy(items).to(dev)
        if loss_layout == "rows-peer-unavailable":
            os.environ["GCF_PEER_DISABLE"] = "1"       # what a box without peer access looks like to peer.available()
        tr = FeatureShardedLightGCNTrainer(users_t, items_t, n_users, n_items, d=table0.shape[1], n_layers=k, lr=0.01,
                                           reg_weight=1e-4, init_table=torch.from_numpy(table0),
                                           loss_layout=loss_layout.split("-")[0], overlap=loss_layout.endswith("-overlap"),
                                           exchange="nccl" if loss_layout in ("rows-nccl", "rows-overlap") else "peer",
                                           user_rows="natural" if loss_layout == "rows-natural" else "owner")
        if loss_layout in ("rows", "rows-pad", "rows-peer-overlap"):
            assert tr.owner_major and tr.n_users % world == 0 and tr.n_users >= n_users
        if loss_layout == "rows-peer-overlap":
            assert tr.overlap and tr.exchange == "peer"
        if loss_layout == "rows-peer-unavailable":
            assert tr.exchange == "nccl"               # every rank fell back to the NCCL exchange
        losses = [float(tr.step(neg_items=torch.from_numpy(negs[s]).to(dev)).item()) for s in range(negs.shape[0])]
        table = tr.gathered_table().cpu().numpy()
        if rank == 0:
            out_q.put((losses, table))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("loss_layout", ["rows", "rows-pad", "rows-natural", "rows-peer-overlap", "rows-peer-unavailable", "rows-nccl",
                                         "rows-overlap", "scores"])
def test_feature_sharded_trainer_matches_single_gpu(loss_layout):
    """"rows" = the default: slices pulled out of the peers' memory over NVLink (csrc/peer.cu); "rows-nccl" / "rows-overlap" =
    the NCCL exchanges with layout passes; "scores" = one all-reduce of partial scores."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    from recommendation_b200.graph import CSRGraph
    from recommendation_b200.lightgcn import FusedLightGCNTrainer

    inter = synth.power_law_bipartite(3000, 4000, 100000, seed=6)
    # "rows-pad": three more users than the interactions name (isolated nodes), so U is not a multiple of the world size and
    # the owner-major user blocks carry padding rows
    U, I, d, k, steps = (3003 if loss_layout == "rows-pad" else 3000), 4000, 64, 3, 3
    rng = np.random.default_rng(1)
    table0 = (rng.standard_normal((U + I, d)) * 0.05).astype(np.float32)
    negs = rng.integers(0, I, (steps, inter.n_edges))
    dev = torch.device("cuda", 0)
    users_t, items_t = torch.from_numpy(inter.users).to(dev), torch.from_numpy(inter.items).to(dev)
    g = CSRGraph.from_pairs(users_t, items_t, U, I, norm="sym")
    ref = FusedLightGCNTrainer(g, U, I, torch.from_numpy(table0).to(dev), users_t, items_t, n_layers=k, lr=0.01, reg_weight=1e-4)
    want_losses = [float(ref.step(neg_i=torch.from_numpy(negs[s]).to(dev)).item()) for s in range(steps)]
    world = min(torch.cuda.device_count(), 8)
    world = 8 if world >= 8 else (4 if world >= 4 else 2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_feature_worker, args=(r, world, port, U, I, inter.users, inter.items, table0, negs, k, q,
                                                            loss_layout)) for r in range(world)]
    for p in procs:
        p.start()
    got_losses, got_table = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    np.testing.assert_allclose(got_losses, want_losses, rtol=1e-4)
    assert got_table.shape == (U + I, d)
    _tables_close(got_table, ref.table.cpu().numpy())


# ---------------------------------------------------------------------------------------------- rows-layout loss
def test_user_block_plan():
    from recommendation_b200.dist import UserBlockPlan

    users = torch.tensor([0, 0, 0, 1, 2, 2, 5, 5, 5, 5, 7, 9], dtype=torch.int64)
    plan = UserBlockPlan.build(users, 10, 3)
    assert plan.cuts[0] == 0 and plan.cuts[-1] == 10 and plan.triple_cuts[0] == 0 and plan.triple_cuts[-1] == 12
    assert list(plan.cuts) == sorted(plan.cuts) and list(plan.triple_cuts) == sorted(plan.triple_cuts)
    for r in range(3):                      # every triple of a block's users, and only those, is in the block's range
        (u0, u1), (t0, t1) = plan.block(r), plan.triple_range(r)
        inside = (users >= u0) & (users < u1)
        assert torch.equal(torch.nonzero(inside).flatten(), torch.arange(t0, t1))
    assert sum(plan.block_sizes()) == 10
    # one hub owning most triples: blocks may be empty but still partition users and triples
    hub = torch.tensor([3] * 50 + [4, 6], dtype=torch.int64)
    p2 = UserBlockPlan.build(hub, 8, 4)
    assert p2.cuts[0] == 0 and p2.cuts[-1] == 8 and sum(p2.block_sizes()) == 8
    assert sum(t1 - t0 for t0, t1 in (p2.triple_range(r) for r in range(4))) == 52
    empty = UserBlockPlan.build(torch.empty(0, dtype=torch.int64), 5, 2)
    assert empty.block_sizes() == [0, 5] and empty.triple_range(0) == (0, 0)


def _rows_loss_gloo_worker(rank, world, port, n_users, n_items, users, items, negs, x, reg, out_q):
    """The collective choreography of FeatureShardedLightGCNTrainer._loss_on_rows on gloo; torch permutes stand in for
    gcf_slices_to_rows / gcf_rows_to_slices and fp64 autograd for the fused BPR kernel."""
    from recommendation_b200.dist import UserBlockPlan, feature_slice

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        G, d = world, x.shape[1]
        lo, hi = feature_slice(d, G, rank)
        dg = hi - lo
        final = torch.from_numpy(x[:, lo:hi]).double().contiguous()           # this rank's column slice of all rows
        order = np.argsort(users * n_items + items, kind="stable")
        su, si, sn = (torch.from_numpy(a[order]) for a in (users, items, negs))
        plan = UserBlockPlan.build(su, n_users, G)
        rows = plan.block_sizes()
        (u0, u1), (t0, t1) = plan.block(rank), plan.triple_range(rank)
        ub = u1 - u0
        item_blk = torch.empty(G * n_items * dg, dtype=torch.float64)
        dist.all_gather_into_tensor(item_blk, final[n_users:].reshape(-1))
        user_blk = torch.empty(G * ub, dg, dtype=torch.float64)
        dist.all_to_all_single(user_blk, final[:n_users].contiguous(), output_split_sizes=[ub] * G, input_split_sizes=rows)
        to_rows = lambda blk, n: blk.view(G, n, dg).permute(1, 0, 2).reshape(n, d)
        to_slices = lambda r, n: r.view(n, G, dg).permute(1, 0, 2).contiguous()
        item_full = to_rows(item_blk, n_items).clone().requires_grad_(True)
        user_full = to_rows(user_blk, ub).clone().requires_grad_(True)
        e = users.shape[0]
        u, p, n = user_full[su[t0:t1] - u0], item_full[si[t0:t1]], item_full[sn[t0:t1]]
        loss = torch.nn.functional.softplus(-(u * (p - n)).sum(1)).sum() / e + reg * (u.pow(2).sum() + p.pow(2).sum())
        loss.backward()
        g_final = torch.empty_like(final)
        dist.reduce_scatter_tensor(g_final[n_users:].reshape(-1), to_slices(item_full.grad, n_items).reshape(-1))
        dist.all_to_all_single(g_final[:n_users], to_slices(user_full.grad, ub).reshape(G * ub, dg), output_split_sizes=rows,
                               input_split_sizes=[ub] * G)
        total = loss.detach().clone()
        dist.all_reduce(total)
        parts = [torch.empty_like(g_final) for _ in range(G)]
        dist.all_gather(parts, g_final)
        if rank == 0:
            out_q.put((float(total), torch.cat(parts, dim=1).numpy()))
    finally:
        dist.destroy_process_group()


def test_rows_layout_loss_exchange_gloo_world2():
    from oracle import losses_ref

    rng = np.random.default_rng(5)
    U, I, E, d, reg = 40, 60, 700, 16, 1e-3
    users, items, negs = rng.integers(0, U, E), rng.integers(0, I, E), rng.integers(0, I, E)
    x = (rng.standard_normal((U + I, d)) * 0.3).astype(np.float32)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rows_loss_gloo_worker, args=(r, 2, port, U, I, users, items, negs, x, reg, q)) for r in range(2)]
    for p in procs:
        p.start()
    got_loss, got_grad = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    xt = torch.from_numpy(x).double().requires_grad_(True)
    want = losses_ref.bpr_lightgcn(xt[:U], xt[U:], torch.from_numpy(users), torch.from_numpy(items), torch.from_numpy(negs), reg)
    want.backward()
    np.testing.assert_allclose(got_loss, float(want), rtol=1e-10)
    np.testing.assert_allclose(got_grad, xt.grad.numpy(), rtol=1e-9, atol=1e-12)


def test_owner_major_rows_are_a_blockwise_bijection():
    """dist.owner_major_row: user u -> row (u % G) * ceil(U / G) + u // G is injective, keeps a rank's users contiguous and in
    ascending order, and leaves at most one padding row at the end of a block; consistent with cyclic_user_shard's local rows."""
    from recommendation_b200.dist import cyclic_user_shard, owner_major_row

    for n_users, world in ((10, 2), (11, 4), (3003, 8), (5, 8), (64, 8)):
        per = -(-n_users // world)
        u = torch.arange(n_users)
        rows = owner_major_row(u, world, per)
        assert rows.unique().numel() == n_users and int(rows.max()) < per * world
        for g in range(world):
            mine = rows[u % world == g]
            n_g = len(range(g, n_users, world))
            assert mine.tolist() == list(range(g * per, g * per + n_g)) and per - n_g in (0, 1) or n_g == 0
        # the row inside the owner's block is what cyclic_user_shard hands the loss kernel as the local user row
        ut = torch.sort(torch.randint(0, n_users, (500,))).values
        for g in range(world):
            pos, loc_u, _, _ = cyclic_user_shard(ut, torch.zeros_like(ut), n_users, world, g)
            assert torch.equal(owner_major_row(ut[pos], world, per), g * per + loc_u)


def test_cyclic_user_shard_partitions_the_triples():
    """dist.cyclic_user_shard (host logic of the peer-memory exchange): every triple belongs to exactly one rank, a rank's list
    stays user-major, local rows map back to the users, row counts add up -- for worlds that do and do not divide U."""
    from recommendation_b200.dist import cyclic_user_shard

    rng = np.random.default_rng(3)
    for n_users, world in ((10, 2), (11, 4), (1000, 8), (5, 8)):
        u = np.sort(rng.integers(0, n_users, 4000))
        i = rng.integers(0, 77, 4000)
        ut, it = torch.from_numpy(u), torch.from_numpy(i)
        seen = torch.zeros(4000, dtype=torch.int64)
        rows_total = 0
        for r in range(world):
            pos, loc_u, loc_i, rows = cyclic_user_shard(ut, it, n_users, world, r)
            seen[pos] += 1
            assert bool((pos[1:] > pos[:-1]).all()) if pos.numel() > 1 else True
            assert torch.equal(loc_u * world + r, ut[pos]) and torch.equal(loc_i, it[pos])
            assert bool((loc_u[1:] >= loc_u[:-1]).all()) if loc_u.numel() > 1 else True      # still user-major
            assert loc_u.numel() == 0 or int(loc_u.max()) < rows[r]
            assert rows == [len(range(g, n_users, world)) for g in range(world)]
            rows_total = sum(rows)
        assert bool((seen == 1).all()) and rows_total == n_users


# ---------------------------------------------------------------------------------------------- world-size independence
def _nccl_default_init_worker(rank, world, port, n_users, n_items, users, items, k, steps, kind, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from recommendation_b200.dist import FeatureShardedLightGCNTrainer, ShardedLightGCNTrainer

        users_t, items_t = torch.from_numpy(users).to(dev), torch.from_numpy(items).to(dev)
        kw = dict(d=64, n_layers=k, lr=0.01, reg_weight=1e-4, seed=1234)      # init_table=None, Philox negatives
        if kind == "row-sharded":
            tr = ShardedLightGCNTrainer(users_t, items_t, n_users, n_items, **kw)
        else:
            tr = FeatureShardedLightGCNTrainer(users_t, items_t, n_users, n_items, loss_layout=kind, **kw)
        losses = [float(tr.step().item()) for _ in range(steps)]
        table = tr.gathered_table().cpu().numpy()
        if rank == 0:
            out_q.put((losses, table))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["rows", "scores", "row-sharded"])
def test_seeded_runs_do_not_depend_on_the_number_of_ranks(kind):
    """Initial table (tables.xavier_uniform_table) and Philox negatives (gcf_sample_negatives_at) are functions of the seed and
    of the canonical user-major triple position only: an N-GPU run reproduces the single-GPU run (bench.py's `check` block)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    from recommendation_b200.graph import CSRGraph
    from recommendation_b200.lightgcn import FusedLightGCNTrainer
    from recommendation_b200.tables import xavier_uniform_table

    inter = synth.power_law_bipartite(3000, 4000, 100000, seed=9)
    U, I, d, k, steps = 3000, 4000, 64, 3, 3
    dev = torch.device("cuda", 0)
    users_t, items_t = torch.from_numpy(inter.users).to(dev), torch.from_numpy(inter.items).to(dev)
    g = CSRGraph.from_pairs(users_t, items_t, U, I, norm="sym")
    table0 = xavier_uniform_table(U, I, d, seed=1234, device=dev)
    assert torch.unique(table0, dim=0).shape[0] == U + I
    ref = FusedLightGCNTrainer(g, U, I, table0.clone(), users_t, items_t, n_layers=k, lr=0.01, reg_weight=1e-4, seed=1234)
    want_losses = [float(ref.step().item()) for _ in range(steps)]
    world = min(torch.cuda.device_count(), 8)
    world = 8 if world >= 8 else (4 if world >= 4 else 2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_default_init_worker, args=(r, world, port, U, I, inter.users, inter.items, k, steps, kind, q))
             for r in range(world)]
    for p in procs:
        p.start()
    got_losses, got_table = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    np.testing.assert_allclose(got_losses, want_losses, rtol=1e-4)
    _tables_close(got_table, ref.table.cpu().numpy())
