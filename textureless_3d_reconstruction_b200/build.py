"""In-tree build of libt3d.so (sm_100a only) with plain nvcc.

The shared library is the C-ABI product (include/t3d.h).  It is built next to
this package so that it travels to the GPU box with the source snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libt3d.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fno-fast-math",
    "--fmad=false",          # bit-exact f32/f64 parity with the gcc oracle
    "-Xptxas", "-v",
]


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "t3d.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libt3d.so cannot be built (no CPU fallback exists)")
    obj_dir = PKG_DIR / "build"
    obj_dir.mkdir(exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = obj_dir / (src.stem + ".o")
        objs.append(obj)
        headers = list(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "t3d.h"]
        if (not force and obj.exists() and obj.stat().st_mtime > src.stat().st_mtime
                and all(obj.stat().st_mtime > hdr.stat().st_mtime for hdr in headers)):
            continue
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log_lines = []
    for src, p in procs:
        out, _ = p.communicate()
        log_lines.append(f"==== {src.name}\n{out}")
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src}")
    (obj_dir / "ptxas.log").write_text("\n".join(log_lines))
    if verbose:
        print("\n".join(log_lines))
    # the link step gets the same -gencode: without it nvcc emits a stub for its default (sm_52) target
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB_PATH), *map(str, objs), "-lcudart"]
    subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
