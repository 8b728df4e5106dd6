"""Attributes the warp-state samples / executed instructions of an `ncu --set full --import-source on`
capture of conv_tc_kernel<128, halo> to source lines and to the kernel's roles.

    ncu -i conv.ncu-rep --page source --csv --print-source sass > sass.csv
    cuobjdump -xelf conv_tc.sm_100a.cubin diffsep_b200/libdsep.so; nvdisasm -g -c conv_tc.sm_100a.cubin > disasm.txt
    python tools/ncu_regions.py sass.csv disasm.txt [regions|lines] [conv_tc.cu of that build]
"""
import collections
import csv
import re
import sys

sass_csv, disasm = sys.argv[1], sys.argv[2]
mode = sys.argv[3] if len(sys.argv) > 3 else "regions"
src_path = sys.argv[4] if len(sys.argv) > 4 else "diffsep_b200/csrc/conv_tc.cu"   # the source the binary was built from
src = open(src_path).read().split("\n")
# role boundaries from marker comments in the source
def find(marker):
    for i, l in enumerate(src):
        if marker in l:
            return i + 1
    return None
marks = [("patch_load", find("// phase 1: issue the global loads")), ("patch_store", find("// phase 2: y = act")),
         ("kernel-setup", find("conv_tc_kernel(const __grid_constant__")), ("build_tile", find("auto build_tile_patches")),
         ("tma-halo", find("TMA producer (halo mode)")), ("mma-halo", find("MMA issuer (halo mode)")),
         ("tma-tap", find("TMA producer (per-tap mode)")), ("epi-setup", find("--- epilogue")),
         ("epi-fast", find("// ---- fast path")), ("epi-generic", find("constexpr int kChunks = NT / 64;          //")),
         ("worker-loop", find("bool worker_builds = false;")), ("teardown", find("the peer may still multicast"))]
marks = sorted([(n, l) for n, l in marks if l], key=lambda x: x[1])
def region(line):
    r = "pre"
    for n, l in marks:
        if line is not None and line >= l:
            r = n
    return r

text = open(disasm).read().split("\n")
start = next(i for i, l in enumerate(text) if l.startswith(".text.") and "ILi128ELb1ELb0" in l)
end = next(i for i, l in enumerate(text) if i > start and l.startswith(".text."))
cur = ctx = None
offmap = {}
for l in text[start:end]:
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m:
        f = m.group(1).split("/")[-1]; cur = (f, int(m.group(2)))
        if f == "conv_tc.cu":
            ctx = cur[1]
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        offmap[int(m.group(1), 16)] = (cur, ctx, m.group(2))
rows = list(csv.reader(open(sass_csv)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
base = int(data[0][0], 16)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(collections.Counter)
for r in data:
    c, x, t = offmap[int(r[0], 16) - base]
    if mode == "regions":
        key = region(x)
        if c and c[0] == "common.cuh" and 57 <= c[1] <= 72:
            key += ":mbar_wait"
    else:
        key = (region(x), x, c if c and c[0] != "conv_tc.cu" else "")
    a = agg[key]
    a["samples"] += float(r[ix["# Samples"]] or 0); a["inst"] += float(r[ix["Instructions Executed"]] or 0)
    for s in stalls:
        a[s] += float(r[ix[s]] or 0)
ts = sum(a["samples"] for a in agg.values()); ti = sum(a["inst"] for a in agg.values())
print(f"total samples {ts:.0f}, warp instructions {ti / 1e6:.1f} M")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:45]:
    top = sorted(((s, a[s]) for s in stalls), key=lambda kv: -kv[1])[:4]
    print(f"{str(k):50s} samples {a['samples'] / ts * 100:5.1f}% inst {a['inst'] / ti * 100:5.1f}% ",
          " ".join(f"{s[6:]}={v / max(a['samples'], 1) * 100:.0f}%" for s, v in top))
