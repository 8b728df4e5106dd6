"""Builds ``libdsep.so`` (the C-ABI library of hand-written sm_100a kernels) in-tree with nvcc.

nvcc cross-compiles without a GPU, so this runs in the build container as well as on the B200
box.  The library is rebuilt whenever the hash of its sources differs from the one embedded in the binary.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libdsep.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")
    return cand


def sources():
    return sorted(CSRC.glob("*.cu"))


def source_hash() -> str:
    """sha256 over everything the binary depends on: kernel sources, headers, the C-ABI header and the compiler
    flags.  The library carries the hash it was built from (``dsep_source_hash()``), so "is this binary the one these
    sources produce" has an answer that does not depend on file times (a snapshot copy resets those)."""
    import hashlib
    h = hashlib.sha256()
    deps = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "dsep.h"]
    for d in deps:
        h.update(d.name.encode())
        h.update(d.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def built_hash():
    """the source hash embedded in the existing libdsep.so (None: no library / an older build without one)"""
    if not LIB.exists():
        return None
    import re
    m = re.search(rb"dsep-source-hash=([0-9a-f]{16}|unknown)\0", LIB.read_bytes())
    return m.group(1).decode() if m else None


def needs_build() -> bool:
    return built_hash() != source_hash()


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    obj_dir = PKG / "build"
    obj_dir.mkdir(exist_ok=True)
    digest = source_hash()
    nvcc = _nvcc()
    procs = []
    objs = []
    for src in sources():
        obj = obj_dir / (src.stem + ".o")
        objs.append(str(obj))
        cmd = [nvcc, *NVCC_FLAGS, f'-DDSEP_SOURCE_HASH="{digest}"', "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{out}")
    tmp = LIB.with_suffix(".so.tmp")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp), *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
