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
