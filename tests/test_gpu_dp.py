"""2-GPU data-parallel parity: N-rank step == 1-rank step on the concatenated batch."""
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


def _build(device, precision):
    from freud_b200.models.config import TopKAutoEncoderConfig
    from freud_b200.models.topkautoencoder import TopKAutoEncoder
    from freud_b200.trainer import SAETrainer

    torch.manual_seed(0)
    cfg = TopKAutoEncoderConfig.from_dict({"n_dict_components": 1024, "k": 32, "auxk_alpha": 1 / 32})
    model = TopKAutoEncoder(64, cfg).to(device)
    return model


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from freud_b200.parallel import DataParallel
        from freud_b200.trainer import SAETrainer

        g = torch.Generator().manual_seed(7)
        B, T, d = 8, 96, 64
        xs = [torch.randn(B, T, d, generator=g) * (0.5 + torch.rand(T, 1, generator=g)) + 0.2 * torch.randn(d, generator=g)
              for _ in range(2)]
        kw = dict(lr=1e-3, steps=100, clip_thresh=1.0, optimizer="adam", scheduler="linear",
                  scheduler_params={"num_warmup_steps": 2}, dead_feature_threshold=1e9, precision="fp32")
        model = _build(dev, "fp32")
        init = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        tr = SAETrainer(model, dp=DataParallel(), **kw)
        per = B // world
        g0 = None
        for x in xs:
            o = tr.step(x[rank * per:(rank + 1) * per].to(dev))
            if g0 is None:
                g0 = {k: p.grad.detach().cpu().clone() for k, p in tr.params.items()}
                fvu0 = float(o["fvu"])
        torch.cuda.synchronize()
        if rank == 0:
            # the N-rank step against the ORACLE on the concatenated batch (not only against this build's 1-rank step)
            from oracle import sae as osae

            ro = osae.topk_forward(xs[0], init["encoder.weight"], init["encoder.bias"], init["W_dec"], init["b_dec"], 32)
            rgr = osae.topk_backward(xs[0], init["encoder.weight"], init["encoder.bias"], init["W_dec"], init["b_dec"],
                                     ro, 32)
            out["oracle.fvu"] = abs(fvu0 - float(ro.fvu)) / float(ro.fvu)
            for k in g0:
                out["oracle.grad." + k] = float((g0[k] - rgr[k]).abs().max() / rgr[k].abs().max())
            ref = SAETrainer(_build(dev, "fp32"), dp=None, **kw)
            for x in xs:
                r = ref.step(x.to(dev))
            torch.cuda.synchronize()
            errs = {k: float((tr.params[k].data - ref.params[k].data).abs().max() / ref.params[k].data.abs().max())
                    for k in tr.params}
            errs["fvu"] = abs(float(o["fvu"]) - float(r["fvu"])) / float(r["fvu"])
            errs["frames"] = float((tr.num_frames_since_fired != ref.num_frames_since_fired).sum())
            out.update(errs)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_step_equals_single_rank_on_concatenated_batch():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out, "rank 0 reported nothing"
    for k, v in out.items():
        assert v < (1e-12 if k == "frames" else 2e-5), (k, v)


def _l1_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from freud_b200.models.config import L1AutoEncoderConfig
        from freud_b200.models.l1autoencoder import L1AutoEncoder
        from freud_b200.parallel import DataParallel
        from freud_b200.trainer import SAETrainer

        def build():
            torch.manual_seed(0)
            return L1AutoEncoder(64, L1AutoEncoderConfig.from_dict({"n_dict_components": 96, "recon_alpha": 1e2})).to(dev)

        g = torch.Generator().manual_seed(11)
        B, T, d = 8, 80, 64
        xs = [torch.randn(B, T, d, generator=g) for _ in range(3)]
        xs[1][0, :5] = -1.0  # ignored_index entries: the masked count is a global quantity too
        kw = dict(lr=4e-4, steps=100, clip_thresh=1.0, optimizer="radam", scheduler="cosine", precision="fp32")
        tr = SAETrainer(build(), dp=DataParallel(), **kw)
        per = B // world
        for x in xs:
            o = tr.step(x[rank * per:(rank + 1) * per].to(dev))
        torch.cuda.synchronize()
        if rank == 0:
            ref = SAETrainer(build(), dp=None, **kw)
            for x in xs:
                r = ref.step(x.to(dev))
            torch.cuda.synchronize()
            pa, pb = dict(tr.model.named_parameters()), dict(ref.model.named_parameters())
            errs = {k: float((pa[k].data - pb[k].data).abs().max() / pb[k].data.abs().max()) for k in pa}
            for k in ("loss_recon", "loss_l1"):
                errs[k] = abs(float(o[k]) - float(r[k])) / abs(float(r[k]))
            out.update(errs)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_l1_step_equals_single_rank_on_concatenated_batch():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_l1_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out, "rank 0 reported nothing"
    for k, v in out.items():
        assert v < 2e-5, (k, v)


def _fused_worker(rank, world, port, precision, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from freud_b200.parallel import DataParallel
        from freud_b200.trainer import SAETrainer

        g = torch.Generator().manual_seed(7)
        B, T, d = 8, 96, 64
        xs = [torch.randn(B, T, d, generator=g) * (0.5 + torch.rand(T, 1, generator=g)) + 0.2 * torch.randn(d, generator=g)
              for _ in range(3)]
        kw = dict(lr=1e-3, steps=100, clip_thresh=1.0, optimizer="adam", scheduler="linear",
                  scheduler_params={"num_warmup_steps": 2}, dead_feature_threshold=1e9, precision=precision)
        model = _build(dev, precision)
        init = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        tr = SAETrainer(model, dp=DataParallel(fused=True), **kw)
        assert tr.fused_dp
        per = B // world
        fvu0 = norm0 = None
        for x in xs:
            o = tr.step(x[rank * per:(rank + 1) * per].to(dev))
            if fvu0 is None:
                fvu0, norm0 = float(o["fvu"]), float(o["grad_sumsq"].sqrt())
        # the module refuses to run on the sharded (stale) fp32 weights ...
        if precision == "bf16":
            try:
                model(xs[0][:1].to(dev))
                out[f"r{rank}.stale_guard"] = 1.0
            except RuntimeError:
                out[f"r{rank}.stale_guard"] = 0.0
        tr.consolidate()  # ... until they are gathered
        sd = tr.optimizer.state_dict()
        torch.cuda.synchronize()
        if rank == 0:
            from oracle import optim as ooptim
            from oracle import sae as osae

            ro = osae.topk_forward(xs[0], init["encoder.weight"], init["encoder.bias"], init["W_dec"], init["b_dec"], 32,
                                   mode=precision)
            rgr = osae.topk_backward(xs[0], init["encoder.weight"], init["encoder.bias"], init["W_dec"], init["b_dec"],
                                     ro, 32, mode=precision)
            keys = ["encoder.weight", "encoder.bias", "W_dec", "b_dec"]
            _, total = ooptim.clip_grad_norm([rgr[k] for k in keys], 1.0)
            tol = 1e-5 if precision == "fp32" else 4e-3
            out["oracle.fvu"] = abs(fvu0 - float(ro.fvu)) / float(ro.fvu) / tol
            out["oracle.grad_norm"] = abs(norm0 - float(total)) / float(total) / tol
            ref = SAETrainer(_build(dev, precision), dp=None, **kw)
            for x in xs:
                r = ref.step(x.to(dev))
            torch.cuda.synchronize()
            ptol = 2e-5 if precision == "fp32" else 2e-4
            for k in tr.params:
                out["param." + k] = float((tr.params[k].data - ref.params[k].data).abs().max() /
                                          ref.params[k].data.abs().max()) / ptol
            out["fvu"] = abs(float(o["fvu"]) - float(r["fvu"])) / float(r["fvu"]) / ptol
            rsd = ref.optimizer.state_dict()
            for i in range(4):
                for key in ("exp_avg", "exp_avg_sq"):
                    a, b = sd["state"][i][key].cpu(), rsd["state"][i][key].cpu()
                    out[f"state.{i}.{key}"] = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)) / (10 * ptol)
            if precision == "bf16":  # every rank's bf16 copies are the rounded masters
                for k in ("encoder.weight", "W_dec"):
                    out["shadow." + k] = float((tr.optimizer.shadows[k].float() -
                                                tr.params[k].data.to(torch.bfloat16).float()).abs().max())
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_rank_fused_optimizer_step_equals_single_rank(precision):
    """Fused peer-memory gradient exchange + sharded Adam (csrc/collective.cu) == single-rank step on the concatenated
    batch (parameters, optimiser state, loss) and == the oracle (loss, clipped-norm) at step 0; values are normalised by
    their tolerance, so every entry must be < 1."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_fused_worker, args=(2, _free_port(), precision, out), nprocs=2, join=True)
    assert out and "fvu" in out, "rank 0 reported nothing"
    for k, v in out.items():
        assert v < 1.0, (k, v)
