"""GPU parity of the feature search against the reference's top_activations goldens (exact rankings)."""
import json
import os

import numpy as np
import pytest
import torch

from tests.util import load_golden

pytestmark = pytest.mark.gpu


def _write_sets(tmp, z, filenames):
    dense, idx, vals = z["dense"], z["idx"], z["vals"]
    n_files, T, F = dense.shape
    k = idx.shape[-1]
    os.makedirs(f"{tmp}/dense")
    np.save(f"{tmp}/dense/layer_tensors.npy", dense.reshape(n_files, -1))
    json.dump({"tensor_shape": [T, F], "activation_shape": [T, F], "filenames": filenames},
              open(f"{tmp}/dense/layer_metadata.json", "w"))
    os.makedirs(f"{tmp}/indexed")
    np.save(f"{tmp}/indexed/layer_activation_values.npy", vals.reshape(n_files, -1))
    np.save(f"{tmp}/indexed/layer_feature_indices.npy", idx.reshape(n_files, -1))
    json.dump({"tensor_shape": [T, k], "activation_shape": [T, 64], "filenames": filenames},
              open(f"{tmp}/indexed/layer_metadata.json", "w"))


def test_top_activations_matches_reference(tmp_path):
    from freud_b200.dataset.activations import DeviceActivationStore, MemoryMappedActivationDataLoader
    from freud_b200.utils.activations import attach_store, top_activations

    z, meta = load_golden("search")
    filenames = meta["filenames"]
    num_samples = {f: int(s) for f, s in zip(filenames, z["num_samples"])}
    _write_sets(str(tmp_path), z, filenames)
    loaders = {}
    for kind in ("dense", "indexed"):
        dl = MemoryMappedActivationDataLoader(f"{tmp_path}/{kind}", "layer", batch_size=8, dl_max_workers=0)
        attach_store(dl, DeviceActivationStore(dl.dataset, num_samples=num_samples))
        loaders[kind] = dl
    for q, query in enumerate(meta["queries"]):
        pq, mpf = top_activations(loaders[query["kind"]], query["feature"], query["n_files"], query["max_val"],
                                  query["min_val"], query["abs"], True)
        assert [filenames.index(p[0]) for p in pq] == z[f"q{q}.files"].tolist(), query
        assert np.array_equal(np.array([p[2] for p in pq], dtype=np.float64), z[f"q{q}.values"]), query
        assert np.array_equal(np.array([p[3] for p in pq], dtype=np.float64), z[f"q{q}.times"]), query
        assert [len(p[1]) for p in pq] == z[f"q{q}.lens"].tolist()
        assert np.array_equal(np.array(mpf, dtype=np.float64), z[f"q{q}.max_per_file"]), query
        if pq:
            assert np.array_equal(pq[0][1].numpy(), z[f"q{q}.trace0"])


def test_search_large_against_oracle():
    """2000 files x 1500 frames, dense and indexed, 8 seeded queries x {plain, abs, filtered}: exact rankings."""
    from freud_b200 import ops
    from oracle import search as osearch

    rng = np.random.default_rng(1)
    n_files, T, F, k, n = 2000, 1500, 48, 32, 6144
    dense = rng.standard_normal((n_files, T, F)).astype(np.float32)
    n_frames = np.array([osearch.n_frames_from_samples(int(s)) for s in rng.integers(16000, 480001, n_files)])
    idx = rng.integers(0, n, (n_files, T, k)).astype(np.int64)
    vals = np.abs(rng.standard_normal((n_files, T, k))).astype(np.float32)
    names = [str(i) for i in range(n_files)]
    nf_dev = torch.tensor(n_frames, dtype=torch.int32, device="cuda")
    d_dev, v_dev, i_dev = torch.from_numpy(dense).cuda(), torch.from_numpy(vals).cuda(), torch.from_numpy(idx).cuda()
    for feature in rng.integers(0, F, 4):
        for kind in ("dense", "indexed"):
            if kind == "dense":
                acts = dense[:, :, feature]
                vmax, amax, vabs, _ = ops.search_dense(d_dev, nf_dev, int(feature), False)
            else:
                f2 = int(idx[0, 0, feature % k])
                acts = osearch.dense_from_indexed(vals, idx, f2)
                vmax, amax, vabs, _ = ops.search_indexed(v_dev, i_dev, nf_dev, f2, False)
            for (mx, mn, ab) in ((None, None, False), (None, None, True), (3.2, 0.5, False)):
                pq, mpf = osearch.top_activations(acts, names, n_frames, 20, mx, mn, ab, True)
                files, cnt = ops.search_topn(vmax, vabs, ab, mn, mx, 20)
                assert files[: int(cnt)].tolist() == [int(p[0]) for p in pq]
                stat = (vabs if ab else vmax).cpu().numpy().astype(np.float64)
                assert np.array_equal(stat, np.array(mpf))
                sel = files[: int(cnt)].long()
                assert np.array_equal(amax[sel].cpu().numpy() * osearch.TIMESTEP_S, np.array([p[3] for p in pq]))


def _sharded_search_worker(rank, world, port, tmp, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from freud_b200.dataset.activations import DeviceActivationStore, MemoryMappedActivationDataLoader
        from freud_b200.utils.activations import attach_store, top_activations

        z, meta = load_golden("search")
        filenames = meta["filenames"]
        num_samples = {f: int(s) for f, s in zip(filenames, z["num_samples"])}
        bad = []
        for kind in ("dense", "indexed"):
            dl = MemoryMappedActivationDataLoader(f"{tmp}/{kind}", "layer", batch_size=8, dl_max_workers=0)
            attach_store(dl, DeviceActivationStore(dl.dataset, num_samples=num_samples, shard=(rank, world)))
            for q, query in enumerate(meta["queries"]):
                if query["kind"] != kind:
                    continue
                pq, mpf = top_activations(dl, query["feature"], query["n_files"], query["max_val"], query["min_val"],
                                          query["abs"], True)
                ok = ([filenames.index(p[0]) for p in pq] == z[f"q{q}.files"].tolist()
                      and np.array_equal(np.array([p[2] for p in pq], dtype=np.float64), z[f"q{q}.values"])
                      and np.array_equal(np.array([p[3] for p in pq], dtype=np.float64), z[f"q{q}.times"])
                      and np.array_equal(np.array(mpf, dtype=np.float64), z[f"q{q}.max_per_file"])
                      and (not pq or np.array_equal(pq[0][1].numpy(), z[f"q{q}.trace0"])))
                if not ok:
                    bad.append(q)
        out[rank] = bad
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_file_sharded_search_matches_reference(tmp_path):
    """Files split over two ranks: every rank returns the reference's rankings, values and traces."""
    import socket

    import torch.multiprocessing as mp

    z, meta = load_golden("search")
    _write_sets(str(tmp_path), z, meta["filenames"])
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sharded_search_worker, args=(2, port, str(tmp_path), out), nprocs=2, join=True)
    assert dict(out) == {0: [], 1: []}


def test_all_feature_table_equals_per_feature_scans():
    """freud_search_table_dense: column f of the one-pass tables == freud_search_dense(feature=f), for fp32 and fp16
    stores, F that needs several frame groups (F=48), exactly one (F=384) and more quads than threads (F=1280)."""
    from freud_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(5)
    for n_files, T, F, dt in ((300, 333, 48, torch.float32), (200, 1500, 384, torch.float32),
                              (64, 200, 1280, torch.float16), (50, 97, 200, torch.float32)):
        acts = torch.randn((n_files, T, F), device="cuda", generator=g).to(dt)
        acts[:, ::7, :] = acts[:, 1::7, :][:, : acts[:, ::7, :].shape[1]]  # repeated values: arg-max ties
        nf = torch.randint(0, T + 1, (n_files,), device="cuda", generator=g, dtype=torch.int32)
        tv, ta, tb = ops.search_table_dense(acts, nf)
        for f in (0, 1, F // 2, F - 1):
            vmax, amax, vabs, _ = ops.search_dense(acts, nf, f, False)
            assert torch.equal(tv[:, f], vmax) and torch.equal(ta[:, f], amax)
            assert torch.equal(tb[:, f].nan_to_num(nan=-7.0), vabs.nan_to_num(nan=-7.0))


def test_indexed_scan_int32_equals_int64():
    """The narrowed-index scan (4 frames per load instruction, no shuffles on the hit path) == the int64 kernel."""
    from freud_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(6)
    n_files, T, k, n = 500, 700, 32, 300
    r = torch.rand((n_files, T, n), device="cuda", generator=g)
    idx = torch.topk(r, k, dim=-1).indices  # distinct indices per frame
    vals = torch.randn((n_files, T, k), device="cuda", generator=g)  # signed values: abs statistics differ from max
    nf = torch.randint(0, T + 1, (n_files,), device="cuda", generator=g, dtype=torch.int32)
    for f in (0, 5, 150, 299):
        a = ops.search_indexed(vals, idx, nf, f, False)
        b = ops.search_indexed(vals, idx.to(torch.int32), nf, f, False)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
        assert torch.equal(a[2].nan_to_num(nan=-7.0), b[2].nan_to_num(nan=-7.0))
