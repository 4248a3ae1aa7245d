"""GPU: the K-chunked tcgen05 NT kernel and the tcgen05 TN (weight-gradient, MN-major operands) kernel against the FP32
FFMA kernels and a float64 reference: every prologue / epilogue pair, ragged N / K, device-side row counts, bias
sums and the column rotation of the SA2 / SA3 first layers (scripts/check_tc2.py is the same check as a script)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tcgen05_kc_nt_and_tn_match_fp64(cuda):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "check_tc2.py")], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, GADDPG_NO_REBUILD="1"))
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
