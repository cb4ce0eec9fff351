#!/usr/bin/env python
"""Summaries of ncu output for profiles/: `launches <csv>` aggregates a gpu__time_duration launch list per kernel;
`rep <file.ncu-rep> [more...]` prints the handful of raw-page metrics DESIGN.md quotes (duration, DRAM bytes,
throughput, occupancy, top stall reasons)."""
import collections
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "lts__t_bytes.sum"]


def launches(path, skip=0):
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows[skip:]:
        k = r["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e3
    tot = sum(v[1] for v in agg.values())
    print("kernel, launches, total_us, avg_us, share")
    for k, v in agg.items():
        print(f"{k}, {v[0]}, {v[1]:.1f}, {v[1] / v[0]:.1f}, {v[1] / tot:.3f}")
    print(f"total_us, {tot:.1f}")


def step_bytes(path, steps, out_json=None, source=""):
    """Launch list taken with --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over `steps`
    identical steps: per kernel launches / time / DRAM bytes, and the DRAM bytes of ONE step (synth_kernel excluded)."""
    import json
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        a = agg.setdefault(k, {"ids": set(), "ns": 0.0, "rd": 0.0, "wr": 0.0})
        a["ids"].add(r["ID"])
        v = float(r["Metric Value"].replace(",", "") or 0)
        unit = r["Metric Unit"]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}.get(unit, 1.0)
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum": a["ns"] += v * scale
        elif m == "dram__bytes_read.sum": a["rd"] += v * scale
        elif m == "dram__bytes_write.sum": a["wr"] += v * scale
    print("kernel, launches_per_step, us_per_step, dram_read_MB_per_step, dram_write_MB_per_step")
    tot_b = tot_us = 0.0
    per = {}
    for k, a in agg.items():
        if k.startswith("synth"):
            continue
        n = len(a["ids"]) / steps
        us, rd, wr = a["ns"] / 1e3 / steps, a["rd"] / 1e6 / steps, a["wr"] / 1e6 / steps
        print(f"{k}, {n:.2f}, {us:.1f}, {rd:.1f}, {wr:.1f}")
        tot_b += (rd + wr) * 1e6
        tot_us += us
        per[k] = {"launches": n, "us": us, "dram_read_bytes": rd * 1e6, "dram_write_bytes": wr * 1e6}
    print(f"per step: {tot_us:.1f} us (cold, serialised), {tot_b / 1e6:.1f} MB of DRAM traffic")
    if out_json:
        json.dump({"dram_bytes_per_step": tot_b, "ncu_us_per_step": tot_us, "steps_captured": steps, "source": source, "kernels": per},
                  open(out_json, "w"), indent=1)


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"== {path}: {data[0][hdr.index('Kernel Name')][:90]}")
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k} [{units[i]}]: " + ", ".join(r[i] for r in data))
    st = [(h, [float(r[i] or 0) for r in data]) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    st.sort(key=lambda x: -x[1][0])
    print("top stalls (warps per issue-active): " + "; ".join(f"{h.split('issue_stalled_')[1].split('_per_')[0]}={v[0]:.2f}" for h, v in st[:6]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    elif sys.argv[1] == "step_bytes":   # step_bytes <csv> <steps> [out.json] [source text]
        step_bytes(sys.argv[2], int(sys.argv[3]), sys.argv[4] if len(sys.argv) > 4 else None, sys.argv[5] if len(sys.argv) > 5 else "")
    else:
        for p in sys.argv[2:]:
            rep(p)
