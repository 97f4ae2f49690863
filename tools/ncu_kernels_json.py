#!/usr/bin/env python
"""profiles/r02_kernels.json from the round's `ncu --set full` captures (tools/gpu_r02_profile.sh):
python tools/ncu_kernels_json.py church gpurun_out/r02_ncu_church_  -> per bench bucket (select / form_factor / update) the
kernel's duration, issue-slot and FP32-pipe utilisation, active lanes per instruction, DRAM bytes per launch."""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wl, prefix = sys.argv[1], sys.argv[2]
CAPS = {"form_factor": [("ff_batch_kernel", "ff_batch_kernel_s1")], "select": [("predict_kernel", "predict_kernel_s1")],
        "update": [("chain_kernel", "chain_kernel_s1"), ("apply_kernel<false>", "apply_kernel_s2"), ("apply_kernel<true>", "apply_kernel_s3")]}


def metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u, r = rows[0], rows[1], rows[2]
    def g(name, scale=1.0):
        v = float(r[h.index(name)].replace(",", ""))
        unit = u[h.index(name)]
        mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
        return v * mult * scale
    st = sorted(((float(r[i]), n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for i, n in enumerate(h)
                 if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio") and r[i]), reverse=True)
    return {"duration_us": round(g("gpu__time_duration.sum"), 2), "grid": int(g("launch__grid_size")), "block": int(g("launch__block_size")),
            "registers": int(g("launch__registers_per_thread")),
            "issue_frac": round(g("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100, 4),
            "fp32_frac": round(g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active") / 100, 4),
            "active_lanes": round(g("smsp__thread_inst_executed_per_inst_executed.ratio"), 2),
            "warps_active_frac": round(g("sm__warps_active.avg.pct_of_peak_sustained_active") / 100, 4),
            "l1_hit": round(g("l1tex__t_sector_hit_rate.pct") / 100, 4), "l2_hit": round(g("lts__t_sector_hit_rate.pct") / 100, 4),
            "dram_bytes_per_launch": int(g("dram__bytes_read.sum") + g("dram__bytes_write.sum")),
            "dram_frac": round(g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") / 100, 5),
            "top_stalls": {n: round(v, 2) for v, n in st[:4]}}


res = {}
for bucket, caps in CAPS.items():
    ks = {name: metrics(prefix + tag + ".ncu-rep") for name, tag in caps}
    first = ks[caps[0][0]]
    res[bucket] = dict(first, kernel=caps[0][0])
    if len(ks) > 1:
        res[bucket]["kernels"] = ks
    res[bucket]["bound"] = ("issue (SIMT divergence)" if first["issue_frac"] > 0.6 else "latency") if first["dram_frac"] < 0.3 else "hbm"
path = os.path.join(ROOT, "profiles", "r02_kernels.json")
try:
    allk = json.load(open(path))
except (OSError, ValueError):
    allk = {}
allk[wl] = res
allk["_source"] = "ncu --set full --clock-control none, one launch per kernel of the second round of a converged run (tools/gpu_r02_profile.sh); text summaries: profiles/r02_ncu_<workload>_<kernel>.txt"
json.dump(allk, open(path, "w"), indent=1)
print(json.dumps(res, indent=1))
