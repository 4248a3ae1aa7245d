"""CPU: the built library really contains the Blackwell instructions DESIGN.md claims — checked on the SASS of
ga-ddpg_b200/lib/libgaddpg_b200.so with cuobjdump (mnemonics per /opt/skills/guides/B200_PROFILING.md): tcgen05 MMAs with TMEM
accumulators in the row-GEMM kernels, TMA loads / stores, cp.async (LDGSTS) operand staging in the weight-gradient kernel, L2
prefetches in the resident-weight kernel, mma.sync TF32 in the small-M kernel.  A regression that silently falls back to FFMA
code (or drops the TMA path) fails here, before any GPU time is spent."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    from gaddpg_b200 import build

    lib = build.build()
    txt = subprocess.run([exe, "-sass", lib], check=True, capture_output=True, text=True).stdout
    per_fn, name = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            per_fn[name] = []
        elif name is not None:
            per_fn[name].append(line)
    return {k: "\n".join(v) for k, v in per_fn.items()}


def _count(sass, fn_substr, mnemonic):
    return sum(v.count(mnemonic) for k, v in sass.items() if fn_substr in k)


def test_sm100a_only(sass):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    from gaddpg_b200 import build

    elf = subprocess.run([exe, "-lelf", build.LIB], check=True, capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", elf))
    assert archs == {"sm_100a"}, archs


def test_tcgen05_tmem_in_the_row_gemms(sass):
    for fn in ("tc_gemm_nt_kernel", "tc_nt_kc_kernel", "tc_tn_kernel", "sa1_fused_kernel"):
        assert _count(sass, fn, "UTCHMMA") > 0, fn        # tcgen05.mma
        assert _count(sass, fn, "LDTM") > 0, fn           # tcgen05.ld (TMEM -> registers)
        assert _count(sass, fn, "UTCBAR") > 0, fn         # tcgen05.commit -> mbarrier
        assert _count(sass, fn, "SYNCS") > 0, fn          # mbarrier waits / arrives


def test_tma_loads_and_stores(sass):
    assert _count(sass, "tc_gemm_nt_kernel", "UTMALDG") > 0       # pre-split weight images
    assert _count(sass, "tc_gemm_nt_kernel", "UTMASTG") > 0       # output blocks of the store epilogue
    assert _count(sass, "sa1_fused_kernel", "UTMALDG") > 0 and _count(sass, "sa1_fused_kernel", "UTMASTG") > 0


def test_cp_async_ring_and_l2_prefetch(sass):
    assert _count(sass, "tc_tn_kernel", "LDGSTS") > 0             # raw operand ring of the weight-gradient kernel
    assert _count(sass, "skinny_nt_kernel", "LDGSTS") > 0         # cp.async ring of the small-M kernel
    assert _count(sass, "tc_gemm_nt_kernel", "CCTL.E.PF2") > 0    # prefetch.global.L2 of the producers / mask source


def test_mma_sync_tf32_in_the_small_m_kernel(sass):
    assert _count(sass, "skinny_nt_kernel", "HMMA.1688.F32.TF32") > 0
