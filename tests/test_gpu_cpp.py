"""Builds tests/cpp/kat.cpp against include/ninterp.hpp + libninterp_cuda.so and runs it on the GPU."""
import os
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _build():
    from ninterp_b200 import build as nb_build

    so = nb_build.build()
    exe = ROOT / "tests" / "cpp" / "_kat"
    src = ROOT / "tests" / "cpp" / "kat.cpp"
    if not exe.exists() or exe.stat().st_mtime < max(src.stat().st_mtime, (ROOT / "include" / "ninterp.hpp").stat().st_mtime):
        subprocess.run(["g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), str(src), "-o", str(exe),
                        f"-L{so.parent}", "-lninterp_cuda", f"-Wl,-rpath,{so.parent}"], check=True)
    return exe


def test_cpp_mirror_compiles_and_links():
    assert _build().exists()


@pytest.mark.gpu
def test_reference_kats_through_cpp_mirror():
    exe = _build()
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
