"""Builds ninterp_b200/libninterp_cuda.so (sm_100a) with nvcc, in-tree.

The extension is plain CUDA C++ behind a C ABI (include/ninterp_cuda.h); it does not link torch.
`-fmad=false` is part of the contract: the reference never contracts a*b+c, and bit-exact parity
for Linear depends on executing its operation sequence as written.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
# NINT_LIB_VARIANT=checked builds (and ninterp_b200._lib loads) libninterp_cuda_checked.so: the same sources with
# -DNINT_BOUNDS_CHECK, every index checked on the device (csrc/nint_device.cuh)
VARIANT = os.environ.get("NINT_LIB_VARIANT", "")
OBJ = PKG / ("_build" + ("_" + VARIANT if VARIANT else ""))
SO = PKG / ("libninterp_cuda" + ("_" + VARIANT if VARIANT else "") + ".so")

SOURCES = ["nint_api.cu", "nint_fixed_f64.cu", "nint_fixed_f32.cu", "nint_nd_f64.cu", "nint_nd_f32.cu", "nint_tile_f64.cu", "nint_tile_f32.cu", "nint_part_f64.cu", "nint_part_f32.cu", "nint_selftest.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def _stamp() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "ninterp_cuda.h",
                                                                      Path(__file__)]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(os.environ.get("NINT_NVCC_EXTRA", "").encode())
    return h.hexdigest()


def _record_commit() -> None:
    """Leaves the commit the library was built at beside the objects: the GPU box gets a snapshot without .git, and
    bench.py stamps its JSON line from this file there."""
    try:
        root = PKG.parent
        head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=root, capture_output=True, text=True).stdout.strip()
        if not head:
            return
        dirty = subprocess.run(["git", "status", "--porcelain", "--untracked-files=no"], cwd=root, capture_output=True, text=True).stdout.strip()
        OBJ.mkdir(exist_ok=True)
        (OBJ / "commit").write_text(head + ("+dirty" if dirty else ""))
    except Exception:
        pass


def build(force: bool = False, verbose: bool = False) -> Path:
    _record_commit()
    stamp_file = OBJ / "stamp"
    stamp = _stamp()
    if not force and SO.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return SO
    nvcc = _nvcc()
    OBJ.mkdir(exist_ok=True)
    extra = os.environ.get("NINT_NVCC_EXTRA", "").split()  # tuning experiments, e.g. -DNINT_MIN_CTAS=3
    if VARIANT == "checked":
        extra.append("-DNINT_BOUNDS_CHECK")

    def compile_one(src: str):
        obj = OBJ / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJ / (src + ".log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stderr[-4000:]}")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(SO), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    stamp_file.write_text(stamp)
    if verbose:
        print(f"built {SO}")
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
