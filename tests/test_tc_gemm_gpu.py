"""GPU: the tcgen05 3xTF32 row-GEMM against the FP32 FFMA kernel and a float64 reference, all prologue / epilogue
combinations, partial tiles, device-side row counts (scripts/check_tc.py is the same check as a script)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tcgen05_gemm_matches_fp64(cuda):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "check_tc.py")], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, GADDPG_NO_REBUILD="1"))
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
