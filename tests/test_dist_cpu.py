"""CPU tests (gloo, world_size 2) of the multi-rank host logic: sharding of the draws,
distributed Sinkhorn (only the K-vector of column marginals is exchanged) and the
gradient all-reduce scaling.  The CUDA kernels are replaced by torch stand-ins with the
same contract; the product's own loop / all-reduce code is what runs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ganecdotes_oracle as O


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def torch_pass(s, inv_eps, first, u_in, r, c, n_total, ws):
    """contract of gx_sinkhorn_pass + gx_sinkhorn_reduce: local column sums of E*b"""
    k = s.shape[1]
    e = torch.exp(s.double() * inv_eps)
    if first:
        return e.sum(0)
    rr = torch.full((k,), 1.0 / k, dtype=torch.float64) if r is None else r.double()
    a = rr / u_in
    t = (e * a).sum(1)
    cn = (1.0 / n_total) if c is None else c.double()
    b = cn / t
    return (e * b[:, None]).sum(0)


def torch_log_a(u, r):
    k = u.numel()
    rr = torch.full((k,), 1.0 / k, dtype=torch.float64) if r is None else r.double()
    return torch.log(rr / u)


def worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ganecdotes_b200.hfc_with_swav import engine as E
        torch.manual_seed(0)
        n, k = 64, 24
        s_full = 0.08 * torch.randn(n, k)
        per = n // world
        s_loc = s_full[rank * per:(rank + 1) * per]
        group = E.DistGroup(dist.group.WORLD, rank, world)
        la = E.sinkhorn_log_a(s_loc, 10, 0.005, None, n, group, pass_fn=torch_pass, log_a_fn=torch_log_a)
        q_loc = torch.softmax(s_loc.double() / 0.005 + la, dim=1)
        q_ref = O.sinkhorn_knopp(s_full.double(), 10, 0.005)[rank * per:(rank + 1) * per]
        ok1 = torch.allclose(q_loc, q_ref, rtol=1e-8, atol=1e-12)
        # single-rank run on the concatenation gives the same log a
        la1 = E.sinkhorn_log_a(s_full, 10, 0.005, None, n, None, pass_fn=torch_pass, log_a_fn=torch_log_a)
        ok2 = torch.allclose(la, la1, rtol=1e-10, atol=1e-12)
        # gradient convention: every rank scales by 1/N_global, the all-reduce SUMs
        g = torch.full((3,), float(rank + 1))
        dist.all_reduce(g, group=group.pg)
        ok3 = bool((g == sum(range(1, world + 1))).all())
        # transport fallback: without a CUDA device the NVLink exchange cannot be created on any rank; the ranks
        # agree (one all-reduce) to use the all-reduce transport instead of failing or disagreeing
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ok4 = group.ensure_ll(k, "cpu") is None and group.ll_failed and group.ll is None
        ret[rank] = (ok1, ok2, ok3 and ok4)
    finally:
        dist.destroy_process_group()


def test_distributed_sinkhorn_world2():
    world = 2
    port = free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r] == (True, True, True), (r, ret[r])


def test_shard_draws_partitions_the_global_batch():
    from ganecdotes_b200.hfc_with_swav import engine as E
    from ganecdotes_b200.hfc_with_swav.swav_clustering import SwAVClustering
    g = torch.Generator().manual_seed(0)
    b, d, nl, hw, npatch = 4, 8, 3, 16, 2
    view = lambda: E.ViewDraws(layer_no=list(range(b)), pert_z=torch.randn(b, 2 * nl, d, generator=g),
                               angle=[float(i) for i in range(b)], flip=[bool(i % 2) for i in range(b)])
    draws = E.StepDraws(z=torch.randn(b, d, generator=g), view_s=view(), view_t=view(),
                        perms=[[torch.randperm(hw, generator=g) for _ in range(b)] for _ in range(npatch)])
    parts = [SwAVClustering.shard_draws(draws, r, 2) for r in range(2)]
    assert torch.equal(torch.cat([p.z for p in parts]), draws.z)
    assert parts[0].view_s.layer_no + parts[1].view_s.layer_no == draws.view_s.layer_no
    assert torch.equal(torch.cat([p.view_t.pert_z for p in parts]), draws.view_t.pert_z)
    assert parts[1].view_s.angle == draws.view_s.angle[2:]
    for p in range(npatch):
        assert all(torch.equal(a, b_) for a, b_ in zip(parts[0].perms[p] + parts[1].perms[p], draws.perms[p]))
    with pytest.raises(AssertionError):
        SwAVClustering.shard_draws(draws, 0, 3)


def test_row_indices_and_split_k_host_logic():
    from ganecdotes_b200.hfc_with_swav import engine as E
    h = w = 8
    view = E.ViewDraws(layer_no=[0, 1], pert_z=torch.zeros(2, 2, 4), angle=[5.0, -7.0], flip=[True, False])
    perms = [[torch.randperm(h * w) for _ in range(2)] for _ in range(3)]
    row_src, row_img = E.build_row_indices(h, w, view, perms, 10, "cpu")
    assert row_src.shape == (3, 20) and row_src.dtype == torch.int32
    assert row_img.tolist() == [0] * 10 + [1] * 10
    for i in range(2):
        mp_ = O.rotate_flip_index_map(h, w, view.angle[i], view.flip[i])
        assert torch.equal(E.rotate_flip_index_map(h, w, view.angle[i], view.flip[i]), mp_)
        for p in range(3):
            assert torch.equal(row_src[p, i * 10:(i + 1) * 10].long(), mp_[perms[p][i][:10]])
    # split-K keeps every SM busy and never creates empty splits
    for tiles, kit in [(80, 2500), (84, 2500), (3, 40), (1, 5)]:
        s = E.pick_split_k(tiles, kit)
        assert 1 <= s <= max(1, kit // 8)
        work = tiles * s
        assert work / (148 * -(-work // 148)) >= min(1.0, tiles * max(1, kit // 8) / 148) * 0.85
