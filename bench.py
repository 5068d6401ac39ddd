#!/usr/bin/env python
"""Benchmark of the hot path BASELINE.json names: SAE train activation-tokens/s.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--workload c3|c2] [--precision bf16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the train-step body (src/scripts/train_sae.py:421-453: dead mask, TopK SAE forward, loss,
backward, global-norm clip, Adam, LR schedule, dead-latent bookkeeping) over one synthetic batch.

Workload (config.workload):
  c3 (default): Whisper-small residual, d=768, n=24576 (32x), k=32, B=32 files x 1500 frames per GPU, data parallel
                (weak scaling; the configuration BASELINE.json quotes "at 1/2/4/8 B200")
  c2          : Whisper-tiny block 2, d=384, n=6144 (16x), k=32, B=50 (configs/train/tiny_topk.json)
Prints ONE JSON line (rank 0).  `value` = device-timed throughput with inputs resident in HBM; `e2e` = the same
metric through the public API with host (pinned) inputs copied in and the loss read back every step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c3": dict(d=768, n=24576, k=32, B=32, T=1500, lr=1e-4, warmup_steps=1000, auxk_alpha=1 / 32,
               dead_feature_threshold=1e6, desc="TopK SAE d=768 n=24576 (32x) k=32, B=32x1500 tokens per GPU"),
    "c2": dict(d=384, n=6144, k=32, B=50, T=1500, lr=1e-4, warmup_steps=1000, auxk_alpha=1 / 32,
               dead_feature_threshold=1e6, desc="TopK SAE d=384 n=6144 (16x) k=32, B=50x1500 tokens"),
    # feature-sharded (strong scaling): the same 24 000-token batch on every rank, dictionary rows split over ranks
    "c4": dict(d=1280, n=81920, k=32, B=16, T=1500, lr=1e-4, warmup_steps=1000, auxk_alpha=0.0,
               dead_feature_threshold=None, sharded=True,
               desc="TopK SAE d=1280 n=81920 (64x) k=32, B=16x1500 tokens, dictionary sharded over the ranks"),
}
METRIC = "sae_train_activation_tokens_per_sec"


def synth_batch(B, T, d, seed, device="cpu", pin=False):
    """x = randn([B,1500,d]) with a per-frame scale sigma_t so total_variance is non-trivial.
    (Zero mean: a large shared mean makes a handful of latents win every token and most others die, which turns
    the step into the AuxK-live variant; that variant is measured separately, see `auxk_live`.)"""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, d, generator=g) * (0.5 + torch.rand(T, 1, generator=g))
    if pin:
        x = x.pin_memory()
    return x.to(device) if device != "cpu" else x


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region through NVML (the fields of the
    B200_PROFILING.md nvidia-smi line: clocks.sm, clocks.max.sm, power.draw, clocks_event_reasons.*).  NVML is read
    in-process every 20 ms: spawning `nvidia-smi -lms` next to a 100 ms timed region stalls the GPU for tens of
    milliseconds per poll and distorts the measurement it is supposed to qualify."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.handle = None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() \
                else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception as ex:  # noqa: BLE001
            self.handle = None
            self.error = repr(ex)
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:  # noqa: BLE001
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                pw = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                self.samples.append((time.perf_counter(), mhz, rs, pw))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def wait_first_sample(self, timeout=2.0):
        t0 = time.perf_counter()
        while self.thread is not None and not self.samples and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        return time.perf_counter()

    def stop(self, t_begin=None, t_end=None):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "error", "")]}
        self.stop_flag.set()
        self.thread.join(timeout=1.0)
        inside = [s for s in self.samples if t_begin is not None and t_begin <= s[0] <= t_end]
        use = inside if inside else self.samples
        sm = sorted(s[1] for s in use)
        reasons = set()
        for s in use:
            for name, bit in self.REASONS.items():
                if s[2] & bit:
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
                "samples": len(use), "power_w_max": max((s[3] for s in use), default=None), "source": "nvml"}


# ---------------------------------------------------------------------------------------------- reference arm / CPU
def reference_available():
    from oracle import ref_shims

    return ref_shims.reference_available()


def reference_step_factory(w, device, use_autocast=True, seed=0):
    """The UNMODIFIED reference (oracle/_ref, populated by oracle/make_ref.sh; /root/reference in the build
    container): src.models.topkautoencoder.TopKAutoEncoder under torch autograd + clip_grad_norm_ + torch.optim.Adam +
    transformers' linear warm-up, stepped by a verbatim restatement of the loop body train_sae.py:421-453 (that body is
    not a function upstream, it sits inside train()).  Returns (model, step_fn)."""
    from oracle import ref_shims

    ref_shims.install()
    from src.models.config import TopKAutoEncoderConfig
    from src.models.topkautoencoder import TopKAutoEncoder
    from torch.amp import autocast
    from torch.optim import Adam
    from transformers import get_linear_schedule_with_warmup
    import contextlib

    torch.manual_seed(seed)
    autoencoder_config = {"n_dict_components": w["n"], "k": w["k"], "auxk_alpha": w["auxk_alpha"], "multi_topk": False,
                          "normalize_decoder": True, "dead_feature_threshold": w["dead_feature_threshold"] or 1e30}
    cfg = TopKAutoEncoderConfig.from_dict(autoencoder_config)
    model = TopKAutoEncoder(activation_size=w["d"], cfg=cfg)
    dist_model = model.to(device)
    optimizer = Adam(dist_model.parameters(), lr=w["lr"])
    scheduler = get_linear_schedule_with_warmup(optimizer, num_warmup_steps=w["warmup_steps"],
                                                num_training_steps=100000)
    num_frames_since_fired = torch.zeros(model.n_dict_components, device=device, dtype=torch.long)
    dev_type = torch.device(device).type

    def step(activations):
        did_fire = torch.zeros(model.n_dict_components, device=device, dtype=torch.bool)
        optimizer.zero_grad()
        with (autocast(dev_type) if use_autocast else contextlib.nullcontext()):
            dead_mask = num_frames_since_fired > autoencoder_config["dead_feature_threshold"]
            out = dist_model(activations, dead_mask=dead_mask)
            loss = out.fvu + out.auxk_loss + out.multi_topk_fvu / 8
            did_fire[out.encoded.top_indices.flatten()] = True
            num_frames_since_fired.add_(activations.shape[1] * activations.shape[0])
            num_frames_since_fired[did_fire] = 0
        loss.backward()
        torch.nn.utils.clip_grad_norm_(dist_model.parameters(), 1.0)
        optimizer.step()
        scheduler.step()
        return loss, out

    return model, step


def oracle_cpu_step_factory(w, B):
    """Fallback when oracle/_ref is absent: one train step of the oracle port (oracle/sae.py + oracle/optim.py)."""
    from oracle import optim as ooptim
    from oracle import sae as osae

    osae.FAST_TOPK = True  # time torch.topk, as the reference does
    torch.manual_seed(0)
    d, n, k = w["d"], w["n"], w["k"]
    enc = torch.nn.Linear(d, n)
    W_enc = enc.weight.data.clone()
    b_enc = torch.zeros(n)
    W_dec = osae.set_decoder_norm_to_unit_norm(W_enc.clone())
    b_dec = torch.zeros(d)
    state = dict(p=[W_enc, b_enc, W_dec, b_dec], m=[torch.zeros_like(t) for t in (W_enc, b_enc, W_dec, b_dec)],
                 v=[torch.zeros_like(t) for t in (W_enc, b_enc, W_dec, b_dec)], step=0)
    keys = ["encoder.weight", "encoder.bias", "W_dec", "b_dec"]
    x = synth_batch(B, w["T"], d, 100)

    def step():
        W_enc, b_enc, W_dec, b_dec = state["p"]
        out = osae.topk_forward(x, W_enc, b_enc, W_dec, b_dec, k, auxk_alpha=w["auxk_alpha"], mode="bf16")
        grads = osae.topk_backward(x, W_enc, b_enc, W_dec, b_dec, out, k, auxk_alpha=w["auxk_alpha"], mode="bf16")
        clipped, _ = ooptim.clip_grad_norm([grads[k_] for k_ in keys], 1.0)
        state["step"] += 1
        lr = ooptim.linear_warmup_lr(w["lr"], state["step"] - 1, w["warmup_steps"], 100000)
        for i, g in enumerate(clipped):
            state["p"][i], state["m"][i], state["v"][i] = ooptim.adam_step(state["p"][i], g, state["m"][i],
                                                                           state["v"][i], state["step"], lr)
        return float(out.fvu)

    return step, B * w["T"]


def run_cpu_reference(w, steps, warmup, B):
    """Times the reference's own CPU path on all host cores.  Returns (tokens/s, ms/step, threads, kind)."""
    torch.set_num_threads(os.cpu_count() or 1)
    if reference_available():
        _, step = reference_step_factory(w, "cpu", use_autocast=True)  # autocast('cpu') = bf16, as train_sae.py:431
        xs = [synth_batch(B, w["T"], w["d"], 100 + i) for i in range(2)]
        for i in range(warmup):
            float(step(xs[i % 2])[0])
        t0 = time.perf_counter()
        for i in range(steps):
            float(step(xs[i % 2])[0])  # loss.item() every step (train_sae.py:455)
        dt = time.perf_counter() - t0
        return B * w["T"] * steps / dt, dt / steps * 1e3, torch.get_num_threads(), "reference"
    step, tokens = oracle_cpu_step_factory(w, B)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return tokens * steps / dt, dt / steps * 1e3, torch.get_num_threads(), "port"


def run_torch_eager_cuda(w, steps, warmup, device):
    """The same unmodified reference modules on the B200 (stock PyTorch eager: cuBLAS GEMMs, torch.topk, dense
    scatter decode, autograd, foreach Adam) under autocast('cuda') as train_sae.py:431 runs them -- the competitor
    SURVEY.md 2.1 names.  Inputs resident in HBM, CUDA-event timed.  Returns dict or None if the reference is absent."""
    if not reference_available():
        return None
    B, T, d = w["B"], w["T"], w["d"]
    try:
        _, step = reference_step_factory(w, device, use_autocast=True)
        xs = [synth_batch(B, T, d, 1000 + i, device=device) for i in range(3)]
        for i in range(warmup):
            step(xs[i % 3])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            loss, _ = step(xs[i % 3])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res = {"value": B * T / (ms / 1e3), "unit": "tokens/s", "ms_per_step": ms, "steps": steps, "warmup": warmup,
               "loss": float(loss), "what": "unmodified reference TopKAutoEncoder + autograd + clip_grad_norm_ + "
               "torch.optim.Adam, stock PyTorch eager on this GPU under autocast('cuda') (fp16), inputs in HBM"}
    except torch.cuda.OutOfMemoryError as ex:  # the dense [N,n] buffers of the eager path
        res = {"value": None, "error": f"out of memory: {str(ex)[:80]}"}
    finally:
        step = xs = None
        torch.cuda.empty_cache()
    return res


def reference_arm(args, w, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the step on this box's host cores, at the SAME
    config as our arm (same B x T tokens per step, same shape, Adam + clip + warm-up), bounded to a few steps."""
    if rank != 0:
        return
    B = w["B"]
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    val, ms, cores, kind = run_cpu_reference(w, steps, warmup, B)
    sample = (f"{B}x{w['T']} tokens per step (the bench batch; d={w['d']}, n={w['n']}, k={w['k']}), {steps} timed steps "
              f"after {warmup} warm-up, {ms:.0f} ms/step")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if w.get("sharded") else "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['desc']}", "global_batch_tokens": B * w["T"],
                       "note": ("unmodified reference modules (oracle/_ref) under autocast('cpu') = bf16, "
                                "torch.set_num_threads(all cores)") if kind == "reference" else
                       "oracle/_ref absent: CPU port of the reference step (oracle/), bf16-operand mode"},
            "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if torch.cuda.is_available() and not w.get("sharded"):
        torch.cuda.set_device(0)
        line["torch_eager_b200"] = run_torch_eager_cuda(w, 5, 2, torch.device("cuda", 0))
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- our arm
def build_trainer(w, precision, dp, device):
    from freud_b200.models.config import TopKAutoEncoderConfig
    from freud_b200.models.topkautoencoder import TopKAutoEncoder
    from freud_b200.trainer import SAETrainer

    torch.manual_seed(0)
    if w.get("sharded"):
        from freud_b200.sharded import FeatureShardedTopKTrainer

        # same init recipe as TopKAutoEncoder.__init__ (kaiming-uniform encoder, unit-norm decoder copy), built
        # row-block by row-block so no rank ever holds more than it needs
        enc = torch.nn.Linear(w["d"], w["n"])
        W = enc.weight.data
        state = {"encoder.weight": W, "encoder.bias": torch.zeros(w["n"]),
                 "W_dec": W / (W.norm(dim=1, keepdim=True) + torch.finfo(torch.float32).eps),
                 "b_dec": torch.zeros(w["d"])}
        return FeatureShardedTopKTrainer(state, w["k"], lr=w["lr"], steps=100000, clip_thresh=1.0, scheduler="linear",
                                         scheduler_params={"num_warmup_steps": w["warmup_steps"]},
                                         precision=precision, device=device)
    cfg = TopKAutoEncoderConfig.from_dict({"n_dict_components": w["n"], "k": w["k"], "auxk_alpha": w["auxk_alpha"],
                                           "multi_topk": False, "normalize_decoder": True})
    model = TopKAutoEncoder(w["d"], cfg).to(device)
    return SAETrainer(model, lr=w["lr"], steps=100000, clip_thresh=1.0, optimizer="adam", scheduler="linear",
                      scheduler_params={"num_warmup_steps": w["warmup_steps"]},
                      dead_feature_threshold=w["dead_feature_threshold"], precision=precision, dp=dp)


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


_KEYS = ("encoder.weight", "encoder.bias", "W_dec", "b_dec")


def parity_check(w, precision, dp, device, x_local, rank, world, sharded):
    """Run BEFORE the timed region, on the bench workload's own shape and batch; the result rides in the JSON line.

    world == 1 : one step of this build in fp32 mode and in `precision` mode against ONE step of the unmodified
                 reference modules (oracle/_ref) in eager fp32 on this GPU, same init, same batch: fvu, selection
                 flip rate, reconstruction and gradient error.  (Rigorous per-row gradient parity, with tie handling,
                 lives in tests/test_gpu_bench_shapes.py; across 48 000 rows a handful of exact k-th/k+1-th ties flip
                 and each moves one token's contribution between two dictionary rows, so the bench-level gradient
                 check is an aggregate one.)
    world > 1  : one N-rank step (data-parallel, or feature-sharded for c4) against the single-rank step of the same
                 build on the concatenated (dp) / same (sharded) batch, on rank 0: loss, gradients, updated parameters.
    """
    import torch.distributed as dist

    res = {}
    if world == 1 and not sharded:
        if not reference_available():
            return {"ok": None, "skipped": "oracle/_ref absent (run oracle/make_ref.sh where /root/reference exists)"}
        ref_model, ref_step = reference_step_factory(w, device, use_autocast=False)
        ref_model.zero_grad()
        out = ref_model(x_local)  # forward + backward of train_sae.py:433-448 without the optimiser step
        (out.fvu + out.auxk_loss + out.multi_topk_fvu / 8).backward()
        ref_g = {k: p.grad.detach().clone() for k, p in ref_model.named_parameters()}
        ref_fvu, ref_out = out.fvu.detach().clone(), out.sae_out.detach().reshape(-1, w["d"]).clone()
        ref_idx = torch.sort(out.encoded.top_indices.reshape(-1, w["k"]), -1).values
        del out, ref_model, ref_step
        torch.cuda.empty_cache()
        ok = True
        for mode, tol_fvu, tol_out in (("fp32", 1e-5, 1e-5), (precision, 2e-2, 2e-2)) if precision != "fp32" \
                else (("fp32", 1e-5, 1e-5),):
            tr = build_trainer(w, mode, None, device)
            o = tr.step(x_local)
            torch.cuda.synchronize()
            flips = float((torch.sort(o["top_idx"].long(), -1).values != ref_idx).any(-1).float().mean())
            r = {"fvu_rel": abs(float(o["fvu"]) - float(ref_fvu)) / float(ref_fvu), "selection_flip_rate": flips,
                 "sae_out_rel_l2": _rel_l2(o["sae_out"], ref_out),
                 "grad_rel_l2": max(_rel_l2(tr.params[k].grad, ref_g[k]) for k in _KEYS)}
            good = r["fvu_rel"] < tol_fvu and (mode != "fp32" or (flips < 1e-3 and r["grad_rel_l2"] < 2e-2))
            r["ok"] = bool(good)
            ok = ok and good
            res[mode + "_vs_reference_fp32"] = r
            del tr, o
            torch.cuda.empty_cache()
        res["against"] = "unmodified reference modules (oracle/_ref), eager fp32 on this GPU, same init and batch"
        res["ok"] = bool(ok)
        return res
    # ---- multi-rank: N-rank step vs the single-rank step on rank 0
    tr = build_trainer(w, precision, dp, device)
    o = tr.step(x_local)
    torch.cuda.synchronize()
    if sharded:
        state = tr.gathered_state()
        full_x = x_local
        grads = None
    elif getattr(tr, "fused_dp", False):
        tr.consolidate()  # fp32 masters are sharded by the fused optimiser: gather them for the comparison
        parts = [torch.empty_like(x_local) for _ in range(world)]
        dist.all_gather(parts, x_local)
        full_x = torch.cat(parts, 0) if rank == 0 else None
        del parts
        state = {k: tr.params[k].data for k in _KEYS}
        grads = None  # each rank holds the summed gradient of its own slice only
    else:
        parts = [torch.empty_like(x_local) for _ in range(world)]
        dist.all_gather(parts, x_local)
        full_x = torch.cat(parts, 0) if rank == 0 else None
        del parts
        state = {k: tr.params[k].data for k in _KEYS}
        grads = {k: tr.params[k].grad for k in _KEYS}
    if rank == 0:
        if sharded:
            single = build_trainer({**w, "sharded": False}, precision, None, device)
        else:
            single = build_trainer(w, precision, None, device)
        so = single.step(full_x)
        torch.cuda.synchronize()
        errs = {"loss_rel": abs(float(o["loss"]) - float(so["loss"])) / abs(float(so["loss"]))}
        errs["param_rel"] = max(_rel(state[k], single.params[k].data) for k in _KEYS)
        if grads is not None:
            errs["grad_rel"] = max(_rel(grads[k], single.params[k].grad) for k in _KEYS)
        errs["max_rel"] = max(errs.values())
        errs["ok"] = bool(errs["max_rel"] < 1e-4)
        errs["against"] = "single-rank step of this build on the " + ("same" if sharded else "concatenated") + " batch"
        res = errs
        del single, so
    del tr, o, full_x
    torch.cuda.empty_cache()
    dist.barrier()
    return res


def symmetric_memory_ok(device):
    """All ranks agree on whether torch's symmetric memory (NVLink peer mappings) works on this box."""
    import torch.distributed as dist

    ok = 1
    try:
        import torch.distributed._symmetric_memory as symm_mem

        t = symm_mem.empty(1024, dtype=torch.float32, device=device)
        symm_mem.rendezvous(t, dist.group.WORLD)
    except Exception as ex:  # noqa: BLE001
        print(f"[bench] symmetric memory unavailable ({ex!r}): falling back to the NCCL all-reduce path", file=sys.stderr)
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item())


# ---------------------------------------------------------------------------------------------- secondary workloads
def _timed(fn, iters, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(iters):
        fn(i)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def extras_single_gpu(device, peaks):
    """The other BASELINE.json configurations, measured briefly on this GPU so that they ride in the driver's BENCH
    line (the headline workload above is C3): C2 (bf16), C3 with AuxK live, C1 (L1 SAE), C5 (feature search)."""
    from freud_b200 import ops
    from freud_b200.models.config import L1AutoEncoderConfig
    from freud_b200.models.l1autoencoder import L1AutoEncoder
    from freud_b200.trainer import SAETrainer

    out = {}

    def guarded(name, fn):
        try:
            out[name] = fn()
        except Exception as ex:  # noqa: BLE001 -- a secondary measurement must never take the headline line down
            out[name] = {"error": repr(ex)[:200]}
        torch.cuda.empty_cache()

    def topk_case(wl, dead_frac):
        w = WORKLOADS[wl]
        tr = build_trainer(w, "bf16", None, device)
        xs = [synth_batch(w["B"], w["T"], w["d"], 50 + i, device=device) for i in range(3)]
        if dead_frac:
            dead = torch.randperm(w["n"], device=device)[: int(w["n"] * dead_frac)]
            tr.tokens_seen = 10 ** 12

            def step(i):
                tr.num_frames_since_fired[dead] = 10 ** 9  # a fixed share of the latents stays dead: AuxK every step
                tr.step(xs[i % 3])
        else:
            def step(i):
                tr.step(xs[i % 3])
        ms = _timed(step, 10)
        return {"ms_per_step": ms, "tokens_per_s": w["B"] * w["T"] / ms * 1e3, "workload": w["desc"]}

    guarded("c2_bf16", lambda: topk_case("c2", 0.0))
    guarded("c3_auxk_live_10pct_dead", lambda: topk_case("c3", 0.1))

    def l1_case():
        torch.manual_seed(0)
        m = L1AutoEncoder(384, L1AutoEncoderConfig.from_dict({"n_dict_components": 200, "recon_alpha": 1e4})).to(device)
        tr = SAETrainer(m, lr=4e-4, steps=100000, clip_thresh=1.0, optimizer="radam", scheduler="cosine",
                        precision="bf16", materialize_outputs=False)
        xs = [synth_batch(100, 1500, 384, 70 + i, device=device) for i in range(3)]
        ms = _timed(lambda i: tr.step(xs[i % 3]), 10)
        # 7.7 kB/token of minimum traffic with the latent kept
        return {"ms_per_step": ms, "tokens_per_s": 150000 / ms * 1e3, "hbm_floor_ms": 150000 * 7.7e3 / (peaks["hbm"] * 1e9) * 1e3,
                "workload": "C1: L1 SAE d=384 n=200 B=100x1500 (configs/train/tiny_l1.json), RAdam + cosine"}

    guarded("c1_l1_bf16", l1_case)

    def search_case():
        n_files, T, F = 10000, 1500, 384
        g = torch.Generator(device=device).manual_seed(0)
        n_frames = torch.randint(50, 1501, (n_files,), generator=g, device=device, dtype=torch.int32)
        dense = torch.empty((n_files, T, F), dtype=torch.float32, device=device)
        for s0 in range(0, n_files, 500):
            dense[s0:s0 + 500].normal_(generator=g)
        feats = [int(v) for v in torch.randint(0, F, (16,), generator=g, device=device)]
        tab_ms = _timed(lambda i: ops.search_table_dense(dense, n_frames), 3, warm=1)
        tv, ta, tb = ops.search_table_dense(dense, n_frames)
        scanned = float(n_frames.sum()) * F * 4

        def q(i):
            f = feats[i % 16]
            ops.search_topn(tv[:, f].contiguous(), tb[:, f].contiguous(), bool(i & 1), None, None, 20)

        q_ms = _timed(q, 16)
        exact = True
        for f in feats[:4]:  # rankings of the table path == rankings of the per-feature scan of the store
            vmax, amax, vabs, _ = ops.search_dense(dense, n_frames, f, False)
            for ab in (False, True):
                a, ca = ops.search_topn(vmax, vabs, ab, None, None, 20)
                b, cb = ops.search_topn(tv[:, f].contiguous(), tb[:, f].contiguous(), ab, None, None, 20)
                exact = exact and torch.equal(a, b) and int(ca) == int(cb) and torch.equal(ta[:, f], amax)
        return {"files": n_files, "table_build_ms": tab_ms, "table_build_GBps": scanned / tab_ms / 1e6,
                "frac_of_hbm": scanned / tab_ms / 1e6 / peaks["hbm"], "query_ms": q_ms,
                "files_per_s_per_query": n_files / q_ms * 1e3, "rankings_exact": bool(exact),
                "workload": "C5: 10 000 files x 1500 frames x 384 features fp32 (23 GB), trimmed lengths U{50..1500}; one "
                            "all-feature table pass, then each query is a column gather + ranking kernel"}

    guarded("c5_search_dense", search_case)
    return out


def extra_c4_sharded(device, rank, world):
    """C4 (d=1280, n=81920, dictionary rows sharded over the ranks) on the ranks of this run: ms/step and a parity
    check of one sharded step against the unsharded step on rank 0."""
    import torch.distributed as dist

    w = WORKLOADS["c4"]
    tr = build_trainer(w, "bf16", None, device)
    x = synth_batch(w["B"], w["T"], w["d"], 1000, device=device)
    o = tr.step(x)
    state = tr.gathered_state()
    res = {}
    if rank == 0:
        single = build_trainer({**w, "sharded": False}, "bf16", None, device)
        so = single.step(x)
        torch.cuda.synchronize()
        res["parity"] = {"loss_rel": abs(float(o["loss"]) - float(so["loss"])) / abs(float(so["loss"])),
                         "param_rel": max(_rel(state[k], single.params[k].data) for k in _KEYS)}
        res["parity"]["ok"] = bool(max(res["parity"].values()) < 1e-4)
        del single, so
        torch.cuda.empty_cache()
    dist.barrier()
    xs = [synth_batch(w["B"], w["T"], w["d"], 1001 + i, device=device) for i in range(2)]
    ms = _timed(lambda i: tr.step(xs[i % 2]), 10)
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res.update({"ms_per_step": float(t.item()), "tokens_per_s": w["B"] * w["T"] / float(t.item()) * 1e3,
                "n_gpus": world, "scaling": "strong", "workload": w["desc"]})
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the pre-timing parity check")
    ap.add_argument("--e2e-dtype", default="fp16", choices=["fp16", "bf16", "fp32"],
                    help="dtype of the pinned host batches of the end-to-end loop (fp16: CUDA-collected Whisper stores)")
    ap.add_argument("--dp-mode", default="fused", choices=["fused", "nccl"],
                    help="data parallel gradient exchange: fused peer-memory reduce-scatter + sharded Adam, or NCCL")
    ap.add_argument("--no-eager", action="store_true", help="skip the stock-PyTorch-eager-on-this-GPU arm")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads (C1, C2, C4, C5, AuxK live)")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel share table (json) here")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, w, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    import torch.distributed as dist

    from freud_b200 import _lib

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dp = None
    sharded = bool(w.get("sharded"))
    if world > 1 or sharded:
        if world == 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=device)
        else:
            dist.init_process_group("nccl", device_id=device)
        if not sharded:
            from freud_b200.parallel import DataParallel

            fused = world > 1 and args.dp_mode == "fused" and symmetric_memory_ok(device)
            dp = DataParallel(fused=fused)
    # all work runs on an explicit non-blocking stream: the legacy default stream serialises with copy streams
    main_stream = torch.cuda.Stream(device)
    torch.cuda.set_stream(main_stream)
    tr = build_trainer(w, args.precision, dp, device)
    B, T, d = w["B"], w["T"], w["d"]
    tokens_per_step = B * T * (1 if sharded else world)  # sharded: every rank sees the same batch (strong scaling)
    n_bufs = 3  # 3 x 147 MB (c3) of distinct inputs, each larger than the 126 MB L2
    host = [synth_batch(B, T, d, 1000 + (0 if sharded else 17 * rank) + i, pin=True) for i in range(n_bufs)]
    dev_x = [h.to(device) for h in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    parity = None
    if not args.no_parity:
        parity = parity_check(w, args.precision, dp, device, dev_x[0], rank, world, sharded)

    # ---------------- device-resident throughput (`value`) + live per-kernel timing
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("FREUD_BENCH_NO_CLOCKS"):  # (diagnostic switch: is the sampler perturbing?)
        sampler.start()
        sampler.wait_first_sample()
    # clocks / power state settle first: 40 untimed steps (0.1-0.3 s; the same count on every rank, the steps hold
    # collectives) on a THROW-AWAY trainer, so the measured trainer still starts from step 0: a loop that starts on
    # a GPU still ramping up from the idle set-up phase was measured up to 20 % slow, and 40 extra steps on the
    # measured trainer would move the timed window into a different phase of training (past the dead-latent
    # threshold, where some latents of the randomly initialised c2 dictionary are dead and AuxK is live)
    ramp = build_trainer(w, args.precision, dp, device)
    for i in range(40):
        ramp.step(dev_x[i % n_bufs])
    torch.cuda.synchronize()
    del ramp
    torch.cuda.empty_cache()
    for i in range(args.warmup):
        tr.step(dev_x[i % n_bufs])
    barrier()
    t_begin = sampler.mark()
    k0 = _lib.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        out = tr.step(dev_x[i % n_bufs])
    ev1.record()
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    gpu_launches = _lib.kernel_launches - k0
    clocks = sampler.stop(t_begin, sampler.mark()) if rank == 0 else None
    # second pass of the same K steps with every C-ABI call bracketed by CUDA events (the roofline's live kernel
    # durations); kept out of the timed region above because ~40 event records per step cost a few percent
    _lib.profile = {}
    pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pv0.record()
    for i in range(args.steps):
        tr.step(dev_x[i % n_bufs])
    pv1.record()
    barrier()
    prof = _lib.profile_summary()
    prof_ms_per_step = pv0.elapsed_time(pv1) / args.steps
    _lib.profile = None
    ms_per_step = ms_total / args.steps
    value = tokens_per_step * args.steps / (ms_total / 1e3)
    last_loss = float(out["loss"].item())

    # ---------------- end to end: pinned host batch -> H2D -> step -> loss read back, every step
    # Host batches in float16 by default: the dtype of activation stores collected with Whisper on CUDA
    # (hooked_model.py:106-108 decodes with fp16=True) -- the step widens them on the device.  The same
    # loop with float32 host batches (CPU-collected stores; twice the PCIe bytes) is reported as `e2e_fp32_input`.
    copy_stream = torch.cuda.Stream(device)
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    # feature-sharded ranks all need the SAME batch: each copies B/G files over its own PCIe link and the parts are
    # all-gathered over NVLink (FeatureShardedTopKTrainer.gather_batch) instead of G full copies from the host
    split_feed = sharded and world > 1 and B % world == 0
    per = B // world if split_feed else B

    def run_e2e(dtype):
        hosts = host if dtype == torch.float32 else [h.to(dtype).pin_memory() for h in host]
        stage = [torch.empty((B, T, d), dtype=dtype, device=device) for _ in range(2)]
        part = [torch.empty((per, T, d), dtype=dtype, device=device) for _ in range(2)] if split_feed else None

        def prefetch(i):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[s])
                if split_feed:
                    part[s].copy_(hosts[i % n_bufs][rank * per:(rank + 1) * per], non_blocking=True)
                else:
                    stage[s].copy_(hosts[i % n_bufs], non_blocking=True)
                ready[s].record(copy_stream)

        loss_pin = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_ev = [torch.cuda.Event() for _ in range(2)]

        def e2e_loop(n):
            for s in range(2):
                freed[s].record()
            prefetch(0)
            loss_sum, prev = 0.0, None
            for i in range(n):
                if i + 1 < n:
                    prefetch(i + 1)  # overlaps the next batch's H2D with this step's compute
                torch.cuda.current_stream().wait_event(ready[i % 2])
                if split_feed:
                    tr.gather_batch(part[i % 2], out=stage[i % 2])
                o = tr.step(stage[i % 2])
                freed[i % 2].record()
                # every step's loss is read back (4-byte D2H, train_sae.py:455): the copy into pinned memory is enqueued
                # right behind the step's own kernels and the host picks the value up one step later, after it has
                # enqueued the next step -- a blocking `.item()` here would drain the stream (it waits for everything
                # enqueued so far, the step just launched included) and leave the GPU idle during the next enqueue
                loss_pin[i % 2].copy_(o["loss"].detach().reshape(1), non_blocking=True)
                loss_ev[i % 2].record()
                if prev is not None:
                    loss_ev[prev].synchronize()
                    loss_sum += float(loss_pin[prev][0])
                prev = i % 2
            loss_ev[prev].synchronize()
            loss_sum += float(loss_pin[prev][0])
            return loss_sum

        e2e_loop(2)
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        e2e_loop(args.steps)
        ev1.record()
        barrier()
        ms = max_over_ranks(max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3))
        return ms, (per if split_feed else B) * T * d * hosts[0].element_size() * world

    e2e_dtype = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[args.e2e_dtype]
    e2e_ms, e2e_h2d = run_e2e(e2e_dtype)
    e2e_value = tokens_per_step * args.steps / (e2e_ms / 1e3)
    e2e32 = None
    if e2e_dtype != torch.float32:
        ms32, h2d32 = run_e2e(torch.float32)
        e2e32 = {"value": tokens_per_step * args.steps / (ms32 / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": h2d32,
                 "d2h_bytes_per_step": 4 * world, "ms_per_step": ms32 / args.steps}
    stage = None

    # ---------------- roofline of the dominant kernel (fused encoder GEMM + top-k), from the live events
    peaks = measured_peaks()
    enc_calls, enc_ms = prof.get("freud_topk_encode", (0, 0.0))
    step_kernel_ms = sum(v[1] for v in prof.values())
    roofline = None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_encoder_traffic.json")
    if os.path.exists(tpath) and world == 1:
        with open(tpath) as f:
            traffic = json.load(f).get(f"{args.workload}.{args.precision}")  # bytes/launch from the ncu capture
    if enc_calls:
        # 2*d*n per token (algorithmic, one bf16 pass); a shard multiplies n/world features
        flops = 2.0 * B * T * d * (w["n"] // world if sharded else w["n"])
        achieved = flops / (enc_ms / enc_calls * 1e-3) / 1e12
        roofline = {"kernel": "sm100_topk_kernel (freud_topk_encode)", "bound": "tensor",
                    "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["tf_sustained"], "traffic": traffic,
                    "algorithmic_bytes": (B * T * d + (w["n"] // world if sharded else w["n"]) * d) * 2 + B * T * 256,
                    "peak_source": peaks["src"] +
                    " bf16 sustained (kernel timed inside a long step)",
                    "share_of_step": enc_ms / step_kernel_ms if step_kernel_ms else None}
    shares = {k_: {"calls": c, "ms_per_step": ms / args.steps, "share": ms / step_kernel_ms}
              for k_, (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    if args.profile_out and rank == 0:
        with open(args.profile_out, "w") as f:
            json.dump({"workload": args.workload, "precision": args.precision, "n_gpus": world,
                       "ms_per_step": ms_per_step, "kernels": shares}, f, indent=1)

    extras = None
    want_extras = not args.no_extras and args.workload == "c3" and args.precision == "bf16"
    dp_exchange = None
    if dp is not None:
        dp_exchange = ("fused peer-memory reduce-scatter + sharded Adam + bf16 all-gather" +
                       (" (multimem)" if getattr(tr.optimizer, "multicast", False) else "")) \
            if getattr(tr, "fused_dp", False) else "NCCL all-reduce"
    if want_extras and world > 1:
        # the feature-sharded configuration on the same ranks, so that the driver's scaling run carries it
        del tr, dev_x
        torch.cuda.empty_cache()
        try:
            extras = {"c4_feature_sharded": extra_c4_sharded(device, rank, world)}
        except Exception as ex:  # noqa: BLE001
            extras = {"c4_feature_sharded": {"error": repr(ex)[:200]}}
        tr = dev_x = None
    if rank != 0:
        if dist.is_initialized():
            dist.destroy_process_group()
        return
    cpu_baseline = None
    eager = None
    if world == 1 and not sharded:
        # free this arm's trainers first: the eager reference materialises the dense [N,n] buffers
        del tr, dev_x
        torch.cuda.empty_cache()
        if want_extras:
            extras = extras_single_gpu(device, peaks)
        if not args.no_eager:
            eager = run_torch_eager_cuda(w, 5, 2, device)
        if not args.no_cpu_baseline:
            Bc = 4 if args.workload == "c3" else 8
            v, ms, cores, kind = run_cpu_reference(w, 2, 1, Bc)
            what = ("unmodified reference modules (oracle/_ref) under autocast('cpu')" if kind == "reference"
                    else "oracle/ CPU port of the reference step")
            cpu_baseline = {"value": v, "unit": "tokens/s", "cores": cores, "kind": kind,
                            "sample": f"{Bc}x{T} tokens/step of the same shape, 2 timed steps after 1 warm-up, "
                                      f"{ms:.0f} ms/step ({what}); `--impl reference` times the full {B}x{T} batch"}
    line = {
        "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if sharded else "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['desc']}", "global_batch_tokens": tokens_per_step,
                   "parallelism": (f"feature-sharded x{world}" if sharded else f"dp{world}"),
                   "dp_exchange": dp_exchange,
                   "optimizer": "adam+clip(1.0)+linear-warmup",
                   "l2": f"{n_bufs} rotating input batches of {B * T * d * 4 / 1e6:.0f} MB (> 126 MB L2)"
                   if B * T * d * 4 > 126e6 else f"{n_bufs} rotating input batches ({B * T * d * 4 / 1e6:.0f} MB each, "
                   f"{n_bufs * B * T * d * 4 / 1e6:.0f} MB total > 126 MB L2)"},
        "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": e2e_h2d,
                "d2h_bytes_per_step": 4 * world, "ms_per_step": e2e_ms / args.steps,
                "input_dtype": args.e2e_dtype + " pinned host batches, fed to the step as stored (widened inside its kernels)",
                "loss_readback": "every step: 4-byte D2H into pinned memory enqueued behind the step, read by the host "
                                 "one step later (a blocking .item() would drain the stream)"},
        "e2e_fp32_input": e2e32,

        "gpu_launches": gpu_launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "loss": last_loss, "parity_check": parity, "torch_eager_b200": eager, "extra": extras,
        "kernel_shares": {k_: round(v["share"], 4) for k_, v in shares.items()},
    }
    print(json.dumps(line), flush=True)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
