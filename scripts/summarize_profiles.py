"""Turn gpurun_out/cap_<tag>/ (scripts/capture_profiles.sh) into the tracked artefacts under profiles/.

    python scripts/summarize_profiles.py r01

Needs `ncu` on PATH to read the .ncu-rep files (no GPU needed)."""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
cap = os.path.join(ROOT, "gpurun_out", f"cap_{tag}")
out = os.path.join(ROOT, "profiles")

KEEP = [
    "gpu__time_duration.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second",
]


def bench_lines():
    for name in ("bench_c3", "bench_c2", "bench_c2_fp32", "bench_c4_1gpu", "bench_reference_arm", "bench_extra"):
        src = os.path.join(cap, name + ".json")
        if not os.path.exists(src):
            continue
        text = open(src).read().strip()
        if name != "bench_extra":
            text = json.dumps(json.loads(text.splitlines()[-1]), indent=1)
        with open(os.path.join(out, f"{tag}_{name}.json"), "w") as f:
            f.write(text + "\n")


def launch_list():
    src = os.path.join(cap, "launches_c3.csv")
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    prof = json.load(open(os.path.join(cap, "prof_c3.json")))
    with open(os.path.join(out, f"{tag}_launches_c3.md"), "w") as f:
        f.write(f"# {tag} -- ncu launch list of the bench step (workload c3, bf16, 1 GPU)\n\n")
        f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "
                "launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (scripts/capture_profiles.sh)\n")
        f.write("Per-launch times are cold-cache and serialised under ncu: compare SHARES with the live CUDA-event "
                "shares below.\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:100]}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% |\n")
        f.write("\n## Live CUDA-event shares from `python bench.py --profile-out` (same build; profiled pass of "
                f"{prof.get('steps', 20)} steps)\n\n")
        f.write(f"ms_per_step of the clean timed loop: {prof['ms_per_step']:.3f}\n\n")
        f.write("| C-ABI call | calls | ms/step | share |\n|---|---:|---:|---:|\n")
        for k, v in prof["kernels"].items():
            f.write(f"| `{k}` | {v['calls']} | {v['ms_per_step']:.3f} | {100 * v['share']:.1f}% |\n")


def full_captures():
    traffic = {}
    for fn in sorted(os.listdir(cap)):
        if fn.endswith(".raw.csv"):  # exported on the GPU box (capture_profiles.sh), the report itself is not kept
            name = fn[:-len(".raw.csv")]
            raw = open(os.path.join(cap, fn)).read()
        elif fn.endswith(".ncu-rep"):
            name = fn[:-len(".ncu-rep")]
            raw = subprocess.run(["ncu", "-i", os.path.join(cap, fn), "--page", "raw", "--csv"], capture_output=True,
                                 text=True, check=True).stdout
        else:
            continue
        rows = list(csv.reader(io.StringIO(raw)))
        h, units = rows[0], rows[1]
        r = rows[2]
        with open(os.path.join(out, f"{tag}_ncu_{name}.csv"), "w") as f:
            for r in rows[2:]:  # one block per captured launch
                f.write(f"Kernel Name,,{r[h.index('Kernel Name')]}\n")
                for m in KEEP:
                    if m in h:
                        f.write(f"{m},{units[h.index(m)]},{r[h.index(m)]}\n")
                stalls = []
                for i, c in enumerate(h):
                    if c.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in c:
                        try:
                            stalls.append((float(r[i]), c))
                        except ValueError:
                            pass
                for v, c in sorted(stalls, reverse=True)[:6]:
                    f.write(f"{c},samples,{v:.0f}\n")
            r = rows[2]
        if name.startswith("full_enc"):
            def val(m):
                v, u = float(r[h.index(m)]), units[h.index(m)]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            traffic["c2.bf16" if name.endswith("c2") else "c3.bf16"] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    if traffic:
        traffic["note"] = (f"dram__bytes_read.sum + dram__bytes_write.sum of the fused encoder kernel per launch, "
                           f"ncu --set full, profiles/{tag}_ncu_full_enc_c3.csv")
        with open(os.path.join(out, f"{tag}_encoder_traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)


bench_lines()
launch_list()
if '--no-ncu' not in sys.argv:
    full_captures()
print(sorted(os.listdir(out)))
