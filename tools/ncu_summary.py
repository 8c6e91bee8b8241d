"""Summarise one kernel of an .ncu-rep (ncu --set full capture) as JSON: the metrics DESIGN.md quotes plus the top
warp-stall sites of the SASS source page.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_kernel_ncu_summary.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
]


def page(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--csv", *args], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = page(rep, "--page", "raw")
    head, units, vals = raw[0], raw[1], raw[2]
    d = dict(zip(head, vals))
    u = dict(zip(head, units))
    out = {"report": rep.split("/")[-1], "Kernel Name": d.get("Kernel Name")}
    for k in KEYS:
        if k in d:
            out[k] = d[k] + (" " + u[k] if u.get(k) else "")
    out["warp_stalls_per_issue"] = {k.split("issue_stalled_")[1].split("_per_issue")[0]: round(float(d[k]), 3) for k in d
                                    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
                                    and float(d[k] or 0) >= 0.05}
    src = page(rep, "--page", "source", "--print-source", "sass")
    H, rows = src[1], src[2:]
    ix = {h: i for i, h in enumerate(H)}
    stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in rows) or 1
    top = sorted(rows, key=lambda r: -int(r[ix["# Samples"]] or 0))[:12]
    out["top_stall_sites"] = [{"sass": r[ix["Source"]].strip()[:80], "share_of_samples": round(int(r[ix["# Samples"]]) / tot, 4),
                               "reason": max(stalls, key=lambda s: int(r[ix[s]] or 0))[6:]} for r in top]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
