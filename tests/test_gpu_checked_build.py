"""The bounds-checked build (-DNINT_BOUNDS_CHECK: every device-side index is checked and the kernel traps on a
violation) evaluates the route matrix of tools/sanitize_case.py bit-exactly.  compute-sanitizer is closed on the GPU
pool this project is measured on, so this is the memory-safety evidence for the speculative gathers, the slack-pair
load, the shared-memory lists of the slab passes and the record lists / TMA tiles of the binned route."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.gpu
def test_route_matrix_on_the_bounds_checked_build():
    so = ROOT / "ninterp_b200" / "libninterp_cuda_checked.so"
    if not so.exists():
        subprocess.run([sys.executable, "-m", "ninterp_b200.build"], check=True, cwd=str(ROOT),
                       env=dict(os.environ, NINT_LIB_VARIANT="checked"))
    env = dict(os.environ, NINT_LIB_VARIANT="checked")
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "sanitize_case.py")], cwd=str(ROOT), env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "all cases bit-exact against the oracle" in r.stdout
