"""profiles/membound_rXX.md from gpurun_out/membound.csv (tools/ncu_membound.sh): per memory-bound kernel of one PC
step, the measured DRAM traffic, duration and achieved GB/s against MEASURED_PEAKS.json hbm_gbs."""
import collections
import csv
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
path = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "gpurun_out" / "membound.csv")
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs", 6536.4) if (ROOT / "MEASURED_PEAKS.json").exists() else 6536.4
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
iI, iK, iM, iV, iG = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
launch = collections.OrderedDict()
for r in rows[1:]:
    d = launch.setdefault(r[iI], {"k": r[iK], "grid": r[iG]})
    d[r[iM]] = float(r[iV].replace(",", ""))
agg = collections.OrderedDict()
for d in launch.values():
    name = re.sub(r"\(.*", "", d["k"]).replace("void dsep::", "").replace("dsep::", "")
    a = agg.setdefault(name, {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0, "top": None})
    a["n"] += 1
    a["ns"] += d["gpu__time_duration.sum"]
    a["rd"] += d["dram__bytes_read.sum"]
    a["wr"] += d["dram__bytes_write.sum"]
    if a["top"] is None or d["gpu__time_duration.sum"] > a["top"]["gpu__time_duration.sum"]:
        a["top"] = d
print(f"| kernel | launches | total ms | DRAM read MB | DRAM write MB | achieved GB/s (all launches) | largest launch: ms, GB/s, % of {peak:.0f} |")
print("|---|---:|---:|---:|---:|---:|---|")
for name, a in sorted(agg.items(), key=lambda x: -x[1]["ns"]):
    gbs = (a["rd"] + a["wr"]) / a["ns"] if a["ns"] else 0.0
    t = a["top"]
    tg = (t["dram__bytes_read.sum"] + t["dram__bytes_write.sum"]) / t["gpu__time_duration.sum"]
    print(f"| `{name}` | {a['n']} | {a['ns'] / 1e6:.3f} | {a['rd'] / 1e6:.1f} | {a['wr'] / 1e6:.1f} | {gbs:.0f} | "
          f"{t['gpu__time_duration.sum'] / 1e6:.3f} ms, {tg:.0f} GB/s, {100 * tg / peak:.0f} % (grid {t['grid']}) |")
