"""Attributes the warp-state samples / executed instructions of an `ncu --set full --import-source on` capture of
conv_fused_kernel<128, FP8, TWO> to the kernel's roles (TMA producer, MMA issuer, patch builders: load / touch /
convert, epilogue) — the conv_fused.cu counterpart of tools/ncu_regions.py.

    ncu -i conv.ncu-rep --page source --csv --print-source sass > sass.csv
    nvcc <build flags> -cubin -o conv_fused.cubin conv_fused.cu     # the SAME source the capture ran
    nvdisasm -g -c conv_fused.cubin > disasm.txt
    python tools/ncu_regions_fused.py sass.csv disasm.txt [conv_fused.cu of that build] [ILi128ELb1ELb0] [lines]
"""
import collections
import csv
import re
import sys

sass_csv, disasm = sys.argv[1], sys.argv[2]
src_path = sys.argv[3] if len(sys.argv) > 3 else "diffsep_b200/csrc/conv_fused.cu"
inst = sys.argv[4] if len(sys.argv) > 4 else "ILi128ELb1ELb0"
mode = sys.argv[5] if len(sys.argv) > 5 else "regions"
src = open(src_path).read().split("\n")
fname = src_path.split("/")[-1]


def find(marker):
    for i, l in enumerate(src):
        if marker in l:
            return i + 1
    raise SystemExit(f"marker not found: {marker}")


WIDE = "conv_wide" in fname
if WIDE:
    marks = sorted([
        ("setup", find("conv_wide_kernel(const __grid_constant__")),
        ("tma", find("TMA producer: weight stages")),
        ("mma", find("--- MMA issuer")),
        ("builder:plan", find("--- patch builders")),
        ("builder:loop", find("float4 vx[6][2], vy[6][2];")),
        ("epi:setup", find("--- epilogue (4 warps)")),
        ("epi:tile-head", find("const int as = it & 1;")),
        ("epi:residual", find("touch the current buffer BEFORE")),
        ("epi:tmem", find("uint32_t v[32];")),
        ("epi:rows(out,stats)", find("float* const orow")),
        ("teardown", find("the peer may still multicast")),
    ], key=lambda x: x[1])
    bmarks = None
else:
  marks = sorted([
    ("builder:load_rows", find("__device__ __forceinline__ void load_rows")),
    ("builder:touch_rows", find("__device__ __forceinline__ void touch_rows")),
    ("builder:convert", find("__device__ __forceinline__ void convert_store")),
    ("setup", find("conv_fused_kernel(const __grid_constant__")),
    ("tma", find("TMA producer: weight stages")),
    ("mma", find("MMA issuer (TWO")),
    ("builder:plan", find("--- patch builders")),
    ("builder:loop", find("float4 vx[6][2], vy[6][2];")),
    ("epi:setup", find("--- epilogue (4 warps)")),
    ("epi:tile-head", find("const int as = it & 1;")),
    ("epi:chunk", find("const bool n_ok = whole || n < p.cout_store;")),
    ("epi:tmem+transpose", find("uint32_t v[32];")),
    ("epi:rows(out,stats)", find("float4 s1 = make_float4")),
    ("teardown", find("the peer may still multicast")),
  ], key=lambda x: x[1])
# helper functions that live in conv_builders.cuh (conv_wide.cu build): attributed by their own file lines
try:
    bsrc = open(src_path.replace(fname, "conv_builders.cuh")).read().split("\n")
    def bfind(marker):
        return next(i + 1 for i, l in enumerate(bsrc) if marker in l)
    bmarks = sorted([("builder:load_rows", bfind("void load_rows")), ("builder:touch_rows", bfind("void touch_rows")),
                     ("builder:convert", bfind("void convert_store"))], key=lambda x: x[1])
except (OSError, StopIteration):
    bmarks = None


def region(line):
    r = "pre"
    for n, l in marks:
        if line is not None and line >= l:
            r = n
    return r


text = open(disasm).read().split("\n")
start = next(i for i, l in enumerate(text) if l.startswith(".text.") and inst in l)
end = next((i for i, l in enumerate(text) if i > start and l.startswith(".text.")), len(text))
cur = ctx = None
offmap = {}
for l in text[start:end]:
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m:
        f = m.group(1).split("/")[-1]
        cur = (f, int(m.group(2)))
        if f == fname:
            ctx = cur[1]
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        offmap[int(m.group(1), 16)] = (cur, ctx, m.group(2))
rows = list(csv.reader(open(sass_csv)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
base = int(data[0][0], 16)
if len(data) != len(offmap):
    print(f"WARNING: capture has {len(data)} instructions, disassembly {len(offmap)} — not the same build", file=sys.stderr)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(collections.Counter)
for r in data:
    c, x, t = offmap[int(r[0], 16) - base]
    if mode == "regions":
        key = region(x)
        if bmarks and c and c[0] == "conv_builders.cuh":
            key = "builder:plan"
            for n_, l_ in bmarks:
                if c[1] >= l_:
                    key = n_
        if c and c[0] == "common.cuh" and 78 <= c[1] <= 113:
            key += " [mbarrier wait]"
    else:
        key = (region(x), x, c if c and c[0] != fname else "")
    a = agg[key]
    a["samples"] += float(r[ix["# Samples"]] or 0)
    a["inst"] += float(r[ix["Instructions Executed"]] or 0)
    for s in stalls:
        a[s] += float(r[ix[s]] or 0)
ts = sum(a["samples"] for a in agg.values())
ti = sum(a["inst"] for a in agg.values())
print(f"total samples {ts:.0f}, warp instructions {ti / 1e6:.1f} M\n")
print("| region | samples | instructions | top stall reasons (share of the region's samples) |")
print("|---|---:|---:|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:60]:
    top = sorted(((s, a[s]) for s in stalls), key=lambda kv: -kv[1])[:4]
    print(f"| `{k}` | {a['samples'] / ts * 100:.1f} % | {a['inst'] / ti * 100:.1f} % | "
          + ", ".join(f"{s[6:]} {v / max(a['samples'], 1) * 100:.0f} %" for s, v in top) + " |")
