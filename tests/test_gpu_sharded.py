"""2-GPU feature-sharded parity: G-rank sharded step == single-GPU step on the same batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from freud_b200.models.config import TopKAutoEncoderConfig
        from freud_b200.models.topkautoencoder import TopKAutoEncoder
        from freud_b200.sharded import FeatureShardedTopKTrainer
        from freud_b200.trainer import SAETrainer

        torch.manual_seed(0)
        cfg = TopKAutoEncoderConfig.from_dict({"n_dict_components": 1024, "k": 32})
        model = TopKAutoEncoder(64, cfg)
        g = torch.Generator().manual_seed(3)
        model.b_dec.data = 0.1 * torch.randn(64, generator=g)
        model.encoder.bias.data = 0.05 * torch.randn(1024, generator=g)
        state = {k: v.clone() for k, v in model.state_dict().items()}
        xs = [torch.randn(4, 100, 64, generator=g) * (0.5 + torch.rand(100, 1, generator=g)) for _ in range(2)]
        kw = dict(lr=1e-3, steps=100, clip_thresh=1.0, scheduler="linear", scheduler_params={"num_warmup_steps": 2},
                  precision="fp32")
        sh = FeatureShardedTopKTrainer(state, 32, device=dev, **kw)
        first = None
        for x in xs:
            o = sh.step(x.to(dev))
            if first is None:
                lo, hi = sh.lo, sh.lo + sh.n_local
                first = {"fvu": float(o["fvu"]), "idx": o["top_idx"].cpu(), "sae_out": o["sae_out"].cpu(),
                         "gW_dec": sh.plist[2].grad.cpu().clone(), "gW_enc": sh.plist[0].grad.cpu().clone()}
        full = sh.gathered_state()
        torch.cuda.synchronize()
        # every rank: its slice of the step-0 gradients and the global selection against the ORACLE
        from oracle import sae as osae

        ro = osae.topk_forward(xs[0], state["encoder.weight"], state["encoder.bias"], state["W_dec"], state["b_dec"], 32)
        rg = osae.topk_backward(xs[0], state["encoder.weight"], state["encoder.bias"], state["W_dec"], state["b_dec"],
                                ro, 32)
        oe = {"oracle.fvu": abs(first["fvu"] - float(ro.fvu)) / float(ro.fvu),
              "oracle.sae_out": float((first["sae_out"] - ro.sae_out.reshape(-1, 64)).abs().max() / ro.sae_out.abs().max()),
              "oracle.set_mismatch_frac": 1.0 - float((torch.sort(first["idx"].long(), -1).values ==
                                                       torch.sort(ro.top_indices.reshape(-1, 32), -1).values).all(-1).float().mean()),
              "oracle.gW_dec": float((first["gW_dec"] - rg["W_dec"][lo:hi]).abs().max() / rg["W_dec"].abs().max()),
              "oracle.gW_enc": float((first["gW_enc"] - rg["encoder.weight"][lo:hi]).abs().max() / rg["encoder.weight"].abs().max())}
        for k_, v_ in oe.items():
            out[f"r{rank}.{k_}"] = v_
        if rank == 0:
            ref = SAETrainer(model.to(dev), optimizer="adam", dead_feature_threshold=1e9, **kw)
            for x in xs:
                r = ref.step(x.to(dev))
            torch.cuda.synchronize()
            errs = {k: float((full[k] - ref.params[k].data).abs().max() / ref.params[k].data.abs().max()) for k in full}
            errs["fvu"] = abs(float(o["fvu"]) - float(r["fvu"])) / float(r["fvu"])
            same = (torch.sort(o["top_idx"].long(), -1).values == torch.sort(r["top_idx"].long(), -1).values).all(-1)
            errs["set_mismatch_frac"] = 1.0 - float(same.float().mean())
            out.update(errs)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_feature_sharded_step_equals_single_gpu():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out, "rank 0 reported nothing"
    for k, v in out.items():
        assert v < (0.01 if k.endswith("set_mismatch_frac") else 3e-5), (k, v)


def _auxk_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from freud_b200.models.config import TopKAutoEncoderConfig
        from freud_b200.models.topkautoencoder import TopKAutoEncoder
        from freud_b200.sharded import FeatureShardedTopKTrainer
        from freud_b200.trainer import SAETrainer

        torch.manual_seed(0)
        n, d = 1024, 64
        cfg = TopKAutoEncoderConfig.from_dict({"n_dict_components": n, "k": 32, "auxk_alpha": 1 / 32})
        model = TopKAutoEncoder(d, cfg)
        g = torch.Generator().manual_seed(5)
        model.b_dec.data = 0.1 * torch.randn(d, generator=g)
        model.encoder.bias.data = 0.05 * torch.randn(n, generator=g)
        state = {k: v.clone() for k, v in model.state_dict().items()}
        xs = [torch.randn(4, 100, d, generator=g) * (0.5 + torch.rand(100, 1, generator=g)) for _ in range(3)]
        # 50 dead latents (> k_aux = 32), unevenly spread over the two shards; and later only 7 (< k_aux)
        dead_sets = [torch.randperm(n, generator=g)[:50], torch.cat((torch.arange(3), torch.arange(600, 604)))]
        kw = dict(lr=1e-3, steps=100, clip_thresh=1.0, scheduler="linear", scheduler_params={"num_warmup_steps": 2},
                  precision="fp32", dead_feature_threshold=100)
        sh = FeatureShardedTopKTrainer(state, 32, device=dev, auxk_alpha=1 / 32, **kw)
        lo, hi = rank * (n // world), (rank + 1) * (n // world)
        aux = []
        for i, x in enumerate(xs):
            if i > 0:
                ds = dead_sets[i - 1]
                mine = ds[(ds >= lo) & (ds < hi)] - lo
                sh.num_frames_since_fired[mine.to(dev)] = 10 ** 9
            o = sh.step(x.to(dev))
            aux.append(float(o["auxk_loss"]))
        full = sh.gathered_state()
        torch.cuda.synchronize()
        if rank == 0:
            ref = SAETrainer(model.to(dev), optimizer="adam", **kw)
            raux = []
            for i, x in enumerate(xs):
                if i > 0:
                    ref.num_frames_since_fired[dead_sets[i - 1].to(dev)] = 10 ** 9
                r = ref.step(x.to(dev))
                raux.append(float(r["auxk_loss"]))
            torch.cuda.synchronize()
            errs = {k: float((full[k] - ref.params[k].data).abs().max() / ref.params[k].data.abs().max()) for k in full}
            errs["fvu"] = abs(float(o["fvu"]) - float(r["fvu"])) / float(r["fvu"])
            for i in (1, 2):
                errs[f"auxk{i}"] = abs(aux[i] - raux[i]) / abs(raux[i])
            errs["auxk_live"] = 0.0 if (raux[1] > 0 and raux[2] > 0 and raux[0] == 0) else 1.0
            out.update(errs)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_feature_sharded_auxk_equals_single_gpu():
    """AuxK over dead latents spread across shards (more and fewer than k_aux of them) == single-GPU step."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_auxk_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out, "rank 0 reported nothing"
    for k, v in out.items():
        assert v < 3e-5, (k, v)


def _multi_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from freud_b200.models.config import TopKAutoEncoderConfig
        from freud_b200.models.topkautoencoder import TopKAutoEncoder
        from freud_b200.sharded import FeatureShardedTopKTrainer
        from freud_b200.trainer import SAETrainer

        torch.manual_seed(0)
        n, d = 1024, 64
        cfg = TopKAutoEncoderConfig.from_dict({"n_dict_components": n, "k": 32, "multi_topk": True})
        model = TopKAutoEncoder(d, cfg)
        g = torch.Generator().manual_seed(7)
        model.b_dec.data = 0.1 * torch.randn(d, generator=g)
        model.encoder.bias.data = 0.05 * torch.randn(n, generator=g)
        state = {k: v.clone() for k, v in model.state_dict().items()}
        xs = [torch.randn(4, 100, d, generator=g) * (0.5 + torch.rand(100, 1, generator=g)) for _ in range(2)]
        kw = dict(lr=1e-3, steps=100, clip_thresh=1.0, scheduler="linear", scheduler_params={"num_warmup_steps": 2},
                  precision="fp32")
        sh = FeatureShardedTopKTrainer(state, 32, device=dev, multi_topk=True, **kw)
        for x in xs:
            o = sh.step(x.to(dev))
        full = sh.gathered_state()
        torch.cuda.synchronize()
        if rank == 0:
            ref = SAETrainer(model.to(dev), optimizer="adam", dead_feature_threshold=1e9, **kw)
            for x in xs:
                r = ref.step(x.to(dev))
            torch.cuda.synchronize()
            errs = {k: float((full[k] - ref.params[k].data).abs().max() / ref.params[k].data.abs().max()) for k in full}
            errs["fvu"] = abs(float(o["fvu"]) - float(r["fvu"])) / float(r["fvu"])
            errs["multi_topk_fvu"] = abs(float(o["multi_topk_fvu"]) - float(r["multi_topk_fvu"])) / float(r["multi_topk_fvu"])
            errs["sae_out"] = float((o["sae_out"] - r["sae_out"]).abs().max() / r["sae_out"].abs().max())
            same = (torch.sort(o["top_idx"].long(), -1).values == torch.sort(r["top_idx"].long(), -1).values).all(-1)
            errs["set_mismatch_frac"] = 1.0 - float(same.float().mean())
            errs["width"] = 0.0 if o["top_idx"].shape[-1] == 128 == r["top_idx"].shape[-1] else 1.0
            errs["frames"] = float((torch.cat([sh.num_frames_since_fired.cpu()]) !=
                                    ref.num_frames_since_fired.cpu()[: sh.n_local]).float().mean())
            out.update(errs)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_feature_sharded_multi_topk_equals_single_gpu():
    """multi-TopK over shards (top-4k of all latents: per-shard candidates -> all-gather -> global top-4k, partial
    decodes all-reduced) == the single-GPU step: parameters after two steps, both FVUs, the returned 4k encoding and
    the dead-latent counters taken from it (topkautoencoder.py:129-138, train_sae.py:441-446)."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_multi_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out, "rank 0 reported nothing"
    for k, v in out.items():
        assert v < (0.01 if k == "set_mismatch_frac" else 3e-5), (k, v)
