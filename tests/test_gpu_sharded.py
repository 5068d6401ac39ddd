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
        assert v < (0.01 if k == "set_mismatch_frac" else 3e-5), (k, v)
