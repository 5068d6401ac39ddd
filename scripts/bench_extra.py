"""Secondary measurements (not the headline bench line): AuxK-live step, fp32 mode, L1 step (C1), search (C5).
Writes one JSON document; run on the GPU box:  python scripts/bench_extra.py > gpurun_out/bench_extra.json"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from freud_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
out = {}
peaks = bench.measured_peaks()


def timed(fn, iters, warm=4):
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


# ---- TopK step variants
for wl, prec, dead_frac in (("c3", "bf16", 0.0), ("c3", "bf16", 0.1), ("c2", "bf16", 0.0), ("c2", "bf16", 0.1),
                            ("c2", "fp32", 0.0), ("c3", "fp32", 0.0)):
    w = bench.WORKLOADS[wl]
    tr = bench.build_trainer(w, prec, None, dev)
    xs = [bench.synth_batch(w["B"], w["T"], w["d"], 50 + i).to(dev) for i in range(3)]
    if dead_frac:
        n = w["n"]
        dead = torch.randperm(n, device=dev)[: int(n * dead_frac)]
        tr.tokens_seen = 10 ** 12

        def step(i, tr=tr, xs=xs, dead=dead):
            tr.num_frames_since_fired[dead] = 10 ** 9  # keep a fixed 10 % of the latents dead -> AuxK live every step
            tr.step(xs[i % 3])
    else:
        def step(i, tr=tr, xs=xs):
            tr.step(xs[i % 3])
    ms = timed(step, 8)
    out[f"topk_step.{wl}.{prec}.dead{int(dead_frac * 100)}"] = {
        "ms_per_step": ms, "tokens_per_s": w["B"] * w["T"] / ms * 1e3}
    del tr, xs, step  # (the closure's default arguments hold the trainer and its batches)
    import gc
    gc.collect()
    torch.cuda.empty_cache()

# ---- L1 step, C1 (configs/train/tiny_l1.json: d=384, n=200, B=100, RAdam 4e-4, cosine, recon_alpha 1e4)
from freud_b200.models.config import L1AutoEncoderConfig  # noqa: E402
from freud_b200.models.l1autoencoder import L1AutoEncoder  # noqa: E402
from freud_b200.trainer import SAETrainer  # noqa: E402

for prec in ("bf16", "fp32"):
    torch.manual_seed(0)
    m = L1AutoEncoder(384, L1AutoEncoderConfig.from_dict({"n_dict_components": 200, "recon_alpha": 1e4})).to(dev)
    tr = SAETrainer(m, lr=4e-4, steps=100000, clip_thresh=1.0, optimizer="radam", scheduler="cosine", precision=prec,
                    materialize_outputs=False)
    xs = [bench.synth_batch(100, 1500, 384, 70 + i).to(dev) for i in range(3)]
    ms = timed(lambda i: tr.step(xs[i % 3]), 8)
    out[f"l1_step.c1.{prec}"] = {"ms_per_step": ms, "tokens_per_s": 150000 / ms * 1e3}
    del tr, xs, m
    torch.cuda.empty_cache()

# ---- search, C5: 10 000 files x 1500 frames; dense F=384 fp32 (23 GB) and indexed k=32 (vals fp32 + idx int64)
n_files, T, F, k, n = 10000, 1500, 384, 32, 6144
g = torch.Generator(device=dev).manual_seed(0)
n_frames = torch.randint(50, 1501, (n_files,), generator=g, device=dev, dtype=torch.int32)
dense = torch.empty((n_files, T, F), dtype=torch.float32, device=dev)
for s in range(0, n_files, 500):
    dense[s:s + 500].normal_(generator=g)
feats = [int(v) for v in torch.randint(0, F, (16,), generator=g, device=dev)]


def q_dense(i):
    vmax, amax, vabs, _ = ops.search_dense(dense, n_frames, feats[i % 16], False)
    ops.search_topn(vmax, vabs, bool(i & 1), None, None, 20)


ms = timed(q_dense, 16)
sector_bytes = float(n_frames.sum()) * 32  # one 32-byte sector per (file, frame) as stored row-major
out["search.c5.dense"] = {"ms_per_query": ms, "files_per_s": n_files / ms * 1e3,
                          "algorithmic_GBps": float(n_frames.sum()) * 4 / ms / 1e6,
                          "sector_GBps": sector_bytes / ms / 1e6, "hbm_peak_GBps": peaks["hbm"],
                          "frac_of_hbm_sector_traffic": sector_bytes / ms / 1e6 / peaks["hbm"]}
# all-feature table: ONE pass over the 23 GB store, then a query is a column gather + the ranking kernel
torch.cuda.synchronize()
tab_ms = timed(lambda i: ops.search_table_dense(dense, n_frames), 3, warm=1)
tv, ta, tb = ops.search_table_dense(dense, n_frames)
table_bytes = float(n_frames.sum()) * F * 4  # trimmed frames only are read
out["search.c5.dense_table_build"] = {"ms": tab_ms, "GBps": table_bytes / tab_ms / 1e6, "hbm_peak_GBps": peaks["hbm"],
                                      "frac_of_hbm": table_bytes / tab_ms / 1e6 / peaks["hbm"],
                                      "note": "one scan serves every later query (all 384 features)"}


def q_table(i):
    f = feats[i % 16]
    ops.search_topn(tv[:, f].contiguous(), tb[:, f].contiguous(), bool(i & 1), None, None, 20)


ms = timed(q_table, 16)
out["search.c5.dense_query_from_table"] = {"ms_per_query": ms, "files_per_s": n_files / ms * 1e3,
                                           "note": "column gather of the table + ranking kernel, no scan"}
# exact-ranking check of the table path against the per-feature scan
ok = True
for f in feats[:4]:
    vmax, amax, vabs, _ = ops.search_dense(dense, n_frames, f, False)
    a, ca = ops.search_topn(vmax, vabs, False, None, None, 20)
    b, cb = ops.search_topn(tv[:, f].contiguous(), tb[:, f].contiguous(), False, None, None, 20)
    ok = ok and torch.equal(a, b) and int(ca) == int(cb)
out["search.c5.dense_query_from_table"]["rankings_equal_per_feature_scan"] = bool(ok)
del tv, ta, tb
del dense
torch.cuda.empty_cache()
vals = torch.rand((n_files, T, k), generator=g, device=dev)
idx = torch.randint(0, n, (n_files, T, k), generator=g, device=dev, dtype=torch.int32)  # narrowed at load time
feats = [int(v) for v in torch.randint(0, n, (16,), generator=g, device=dev)]


def q_idx(i):
    vmax, amax, vabs, _ = ops.search_indexed(vals, idx, n_frames, feats[i % 16], False)
    ops.search_topn(vmax, vabs, bool(i & 1), None, None, 20)


ms = timed(q_idx, 16)
idx_bytes = float(n_frames.sum()) * k * 4
out["search.c5.indexed"] = {"ms_per_query": ms, "files_per_s": n_files / ms * 1e3,
                            "index_scan_GBps": idx_bytes / ms / 1e6, "hbm_peak_GBps": peaks["hbm"],
                            "frac_of_hbm": idx_bytes / ms / 1e6 / peaks["hbm"]}
print(json.dumps(out, indent=1))
