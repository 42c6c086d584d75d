#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` report -> JSON + markdown (for profiles/).

    python scripts/ncu_summary.py <report.ncu-rep> <out_prefix>
"""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1:3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
M = {
    "duration_us": ("gpu__time_duration.sum", 1e-3, "ns"),
    "dram_read_bytes": ("dram__bytes_read.sum", None, None),
    "dram_write_bytes": ("dram__bytes_write.sum", None, None),
    "dram_pct_of_peak": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1, None),
    "sm_throughput_pct": ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1, None),
    "issue_active_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1, None),
    "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1, None),
    "threads_per_inst": ("smsp__thread_inst_executed_per_inst_executed.ratio", 1, None),
    "warp_instructions": ("smsp__inst_executed.sum", 1, None),
    "registers": ("launch__registers_per_thread", 1, None),
    "fp64_pipe_pct": ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 1, None),
    "tensor_pipe_pct": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1, None),
    "grid": ("launch__grid_size", 1, None),
    "block": ("launch__block_size", 1, None),
    "smem_per_block": ("launch__shared_mem_per_block_static", 1, None),
}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def val(r, name):
    if name not in ix:
        return None
    try:
        v = float(r[ix[name]].replace(",", ""))
    except ValueError:
        return None
    u = units[ix[name]]
    if name.startswith("dram__bytes"):
        return v * UNIT.get(u, 1)
    if name == "gpu__time_duration.sum":
        return v * UNIT.get(u, 1)
    return v


agg = {}
for r in rows[2:]:
    if len(r) <= ix["Kernel Name"]:
        continue
    k = r[ix["Kernel Name"]]
    d = agg.setdefault(k, {"launches": 0})
    d["launches"] += 1
    for key, (name, _, _) in M.items():
        v = val(r, name)
        if v is not None:
            d.setdefault(key, []).append(v)
res = {}
for k, d in agg.items():
    e = {"launches_profiled": d["launches"]}
    for key in M:
        if key in d:
            e[key] = sum(d[key]) / len(d[key])
    if "dram_read_bytes" in e:
        e["dram_traffic_bytes"] = e["dram_read_bytes"] + e.get("dram_write_bytes", 0.0)
        e["dram_GBps"] = e["dram_traffic_bytes"] / (e["duration_us"] * 1e-6) / 1e9
    res[k] = e
json.dump(res, open(out + ".json", "w"), indent=1)
with open(out + ".md", "w") as f:
    f.write(f"# ncu --set full summary of `{rep}` (per launch averages)\n\n")
    f.write("| kernel | µs | DRAM read MB | DRAM write MB | DRAM GB/s | issue active % | warps active % | thr/inst | warp instr (M) | regs | fp64 % | tensor % |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for k, e in res.items():
        f.write("| `{}` | {:.1f} | {:.1f} | {:.1f} | {:.0f} | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {:.0f} | {:.1f} | {:.1f} |\n".format(
            k[:70], e.get("duration_us", 0), e.get("dram_read_bytes", 0) / 1e6, e.get("dram_write_bytes", 0) / 1e6,
            e.get("dram_GBps", 0), e.get("issue_active_pct", 0), e.get("warps_active_pct", 0),
            e.get("threads_per_inst", 0), e.get("warp_instructions", 0) / 1e6, e.get("registers", 0),
            e.get("fp64_pipe_pct", 0), e.get("tensor_pipe_pct", 0) or 0))
print(open(out + ".md").read())
