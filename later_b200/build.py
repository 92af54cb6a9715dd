"""In-tree build of liblater_b200.so (sm_100a only) with plain nvcc.

`python -m later_b200.build` or `__graft_entry__.build()`.  Objects go to build/ (git-ignored),
the shared library to later_b200/liblater_b200.so (git-ignored, but it travels with gpurun).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "later_b200" / "csrc"
BUILD = ROOT / "build"
LIB = ROOT / "later_b200" / "liblater_b200.so"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", *ARCH]
# compat.cu and panel32.cu define __global__ kernels (setEye, clearTri, s2h; mgs_kernel, mgs_kernel2)
# that the reference's test drivers launch from their own translation units -> relocatable device
# code for those files only.
SOURCES = {
    "tc_gemm.cu": [],
    "tc_update.cu": [],
    "tc_gram_cast.cu": [],
    "panel.cu": [],
    "panel_tc.cu": [],
    "panel32.cu": ["-rdc=true"],
    "rgsqrf.cu": [],
    "ormqr.cu": [],
    "qdwh.cu": [],
    "hou.cu": [],
    "mgpu.cu": [],
    "peer_comm.cu": [],
    "compat.cu": ["-rdc=true"],
}


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the CUDA 12.9 toolkit is required to build later_b200")
    return exe


def _digest(extra: list[str]) -> str:
    h = hashlib.sha256()
    for src in sorted(CSRC.iterdir()):          # every source and header, whatever includes what
        if src.suffix in (".cu", ".cuh", ".h"):
            h.update(src.name.encode())
            h.update(src.read_bytes())
    for inc in sorted((ROOT / "include").glob("*.h")):
        h.update(inc.read_bytes())
    h.update(" ".join(COMMON + extra).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, defines: list[str] | None = None) -> Path:
    defines = defines or []
    BUILD.mkdir(exist_ok=True)
    stamp = BUILD / "stamp.txt"
    digest = _digest(defines)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    objs = []
    procs = []
    for src, extra in SOURCES.items():
        obj = BUILD / (src.replace(".cu", ".o"))
        cmd = [nvcc(), *COMMON, *extra, *defines, "-I", str(ROOT / "include"), "-c", str(CSRC / src),
               "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(str(obj))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError(f"nvcc failed on {src}")
    link = [nvcc(), "-shared", *ARCH, *objs, "-o", str(LIB), "-lcurand", "-ldl",
            "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
    if verbose:
        print(" ".join(link), flush=True)
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        sys.stderr.write(r.stdout.decode())
        raise RuntimeError("link of liblater_b200.so failed")
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose=True,
                 defines=[a for a in sys.argv[1:] if a.startswith("-D")])
    print(path)
