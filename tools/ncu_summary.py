"""Summarise ncu output for profiles/: either a launch list CSV (gpu__time_duration.sum per launch)
or the raw page of a `--set full` report.

    python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/launches_rXX.md
    python tools/ncu_summary.py report gpurun_out/conv.ncu-rep  > profiles/conv_rXX.md
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_uniform.sum", "launch__occupancy_limit_shared_mem",
]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(row["Metric Unit"], 1e-6)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"total {tot:.3f} ms over {sum(n for n, _ in agg.values())} launches (serialised, cold-cache: compare shares)\n")
    print("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t:.3f} | {100 * t / tot:.1f} % |")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"### {d.get('Kernel Name', '?')}  grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            for h, u, v in zip(hdr, units, r):
                if h == k:
                    print(f"| {h} | {v} | {u} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
