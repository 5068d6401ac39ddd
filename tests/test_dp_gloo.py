"""world_size-2 gloo tests (CPU) of the data-parallel host logic: the per-rank pieces must reassemble into the
single-process quantities on the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from freud_b200.parallel import DataParallel

        dp = DataParallel()
        g = torch.Generator().manual_seed(0)
        B, T, d = 6, 5, 4
        x = torch.randn(B, T, d, generator=g) + torch.randn(d, generator=g)
        xs = x[rank * (B // world):(rank + 1) * (B // world)]
        # per-rank pieces exactly as freud_topk_prep_x produces them
        colmean = xs.mean(0)
        tv_local = (xs - colmean).pow(2).sum().double().reshape(1)
        tv = dp.global_total_variance(tv_local, colmean, xs.shape[0])
        ref_tv = (x - x.mean(0)).pow(2).sum().double()
        grads = [torch.full((3, 2), float(rank + 1)), torch.full((5,), 10.0 * (rank + 1))]
        dp.all_reduce_grads(grads)
        counts = torch.tensor([rank, 0, 1 - rank], dtype=torch.int32)
        dp.all_reduce_sum(counts)
        ok = (abs(float(tv) - float(ref_tv)) < 1e-5 * float(ref_tv)
              and torch.equal(grads[0], torch.full((3, 2), 3.0)) and torch.equal(grads[1], torch.full((5,), 30.0))
              and counts.tolist() == [1, 0, 1] and dp.world_size == world)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_data_parallel_pieces_reassemble_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out.get(0) is True and out.get(1) is True


def _search_gather_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from types import SimpleNamespace

        from freud_b200.utils.activations import _gather_files

        # 11 files over 2 ranks: blocks of ceil(11/2) = 6 and 5 files (the store's own arithmetic)
        n_total, per = 11, 6
        lo, hi = min(n_total, rank * per), min(n_total, (rank + 1) * per)
        store = SimpleNamespace(shard=(rank, world), per=per, n_total=n_total)
        full_vals = torch.arange(n_total, dtype=torch.float32) * 1.5 - 4.0
        full_arg = torch.arange(n_total, dtype=torch.int32) * 7
        got_v = _gather_files(store, full_vals[lo:hi].clone(), 0.0)
        got_a = _gather_files(store, full_arg[lo:hi].clone(), 0)
        # flat gradient buffer: the no-copy path of all_reduce_grads reduces the views in place
        from freud_b200.parallel import DataParallel

        flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
        views = [flat[:6].view(2, 3), flat[6:]]
        DataParallel().all_reduce_grads(views, flat=flat)
        out[rank] = bool(torch.equal(got_v, full_vals) and torch.equal(got_a, full_arg)
                         and torch.equal(views[0], (torch.arange(6, dtype=torch.float32) * 3).view(2, 3))
                         and torch.equal(views[1], torch.arange(6, 10, dtype=torch.float32) * 3))
    finally:
        dist.destroy_process_group()


def test_sharded_search_gather_and_flat_gradient_bucket_world2():
    """Uneven file shards reassemble in file order on every rank; gradient views of a flat bucket are reduced in place."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_search_gather_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out.get(0) is True and out.get(1) is True
