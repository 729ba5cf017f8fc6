#!/usr/bin/env python
"""Extract the judged numbers from an `ncu --set full` report (run HERE, no GPU needed):

    python tools/ncu_extract.py gpurun_out/prof.ncu-rep --summary profiles/rN_ncu_full_summary.txt \
        --traffic profiles/rN_ncu_traffic.json --source "<command + commit>"

One block per distinct kernel (first launch after the warm-up launches of that kernel): duration,
DRAM bytes, pipe / issue / occupancy figures; the traffic json maps bench.py's kernel labels to
dram__bytes_read.sum + dram__bytes_write.sum per launch (bench.py reports it as roofline.traffic).
"""
import argparse
import csv
import io
import json
import subprocess

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__block_size", "launch__grid_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]
LABELS = {"dot_fwd": "dot_fwd_kernel(K1+K4)", "dot_bwd": "dot_bwd_kernel(K4 bwd)",
          "seg_apply": "seg_apply(K2: segment reduce + sparse Adam)"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--summary")
    ap.add_argument("--traffic")
    ap.add_argument("--source", default="")
    ap.add_argument("--skip", type=int, default=1, help="launches of each kernel to skip (warm-up)")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(head)}
    seen, blocks, traffic = {}, [], {}
    for r in body:
        name = r[col["Kernel Name"]]
        seen[name] = seen.get(name, 0) + 1
        if seen[name] != a.skip + 1:
            continue
        vals = {}
        for m in METRICS:
            if m in col:
                vals[m] = (r[col[m]], units[col[m]])
        blocks.append((name, vals))
        def num(m):
            v, u = vals[m]
            return float(v.replace(",", "")) * SCALE.get(u, 1.0)
        for key, label in LABELS.items():
            if key in name:
                rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
                traffic[label] = {"traffic_bytes": rd + wr, "read": rd, "write": wr,
                                  "ncu_duration_us": num("gpu__time_duration.sum"), "kernel": name[:80]}
    text = [f"Source: {a.source}", ""]
    for name, vals in blocks:
        text.append(f"== {name[:100]}")
        for m, (v, u) in vals.items():
            text.append(f"   {m:<88}{v} {u}")
        text.append("")
    out = "\n".join(text)
    if a.summary:
        open(a.summary, "w").write(out)
    else:
        print(out)
    if a.traffic:
        json.dump({"source": a.source, "kernels": traffic}, open(a.traffic, "w"), indent=1)


if __name__ == "__main__":
    main()
