"""Opcode evidence for the shipped binary: per kernel of libdsep.so, the counts of the SASS mnemonics that show which
hardware path it uses (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG,
tcgen05.commit -> UTCBAR, legacy mma.sync -> HMMA), plus registers / spills from the ELF.  Runs without a GPU.

    python tools/sass_report.py > profiles/sass_r02.md
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from diffsep_b200 import build as b  # noqa: E402

LIB = ROOT / "diffsep_b200" / "libdsep.so"
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMAPF", "SYNCS", "HMMA", "LDG", "STG", "LDS", "STS", "MUFU",
        "F2FP", "REDG", "ATOMG", "LDL", "STL"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", str(LIB)], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
        if m and cur:
            usage[cur] = tuple(int(v) for v in m.groups())
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
            kernels[cur]["_total"] += 1
    print(f"# SASS opcode evidence, libdsep.so (source hash {b.source_hash()}; built hash {b.built_hash()})\n")
    print("`python tools/sass_report.py` — cuobjdump -sass / -res-usage of the shipped library, sm_100a.  `UTCHMMA` / `UTCQMMA` = "
          "tcgen05.mma kind::f16 / kind::f8f6f4, `UTCBAR` = tcgen05.commit, `LDTM` = tcgen05.ld, `UTMALDG` = TMA tensor load, "
          "`SYNCS` = mbarrier ops; `HMMA` (legacy mma.sync) must be absent.\n")
    cols = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UTMALDG", "HMMA", "MUFU", "LDL+STL"]
    print("| kernel | SASS instr. | regs | local B | " + " | ".join(cols) + " |")
    print("|---|---:|---:|---:|" + "---:|" * len(cols))
    for k, c in kernels.items():
        name = demangle(k)
        name = re.sub(r"\(.*", "", name).replace("void dsep::", "").replace("dsep::", "")
        reg, _, local = usage.get(k, (0, 0, 0))
        vals = [c["UTCHMMA"], c["UTCQMMA"], c["UTCBAR"], c["LDTM"], c["UTMALDG"], c["HMMA"], c["MUFU"], c["LDL"] + c["STL"]]
        print(f"| `{name}` | {c['_total']} | {reg} | {local} | " + " | ".join(str(v) for v in vals) + " |")
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print(f"\nWhole library: {tot['UTCHMMA']} UTCHMMA, {tot['UTCQMMA']} UTCQMMA, {tot['LDTM']} LDTM, {tot['UTCBAR']} UTCBAR, "
          f"{tot['UTMALDG']} UTMALDG, {tot['HMMA']} HMMA.")


if __name__ == "__main__":
    main()
