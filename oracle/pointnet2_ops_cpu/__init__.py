"""CPU stand-in for the un-vendored ``pointnet2_ops`` package (test infrastructure only).

Restates the Python surface the reference imports
(/root/reference/core/networks.py:10 ``pointnet2_ops.pointnet2_modules``,
/root/reference/core/utils.py:32 ``pointnet2_ops.pointnet2_utils``) on top of
``oracle/pointnet2_cpu.c``.  PARITY UNPINNED — see oracle/__init__.py.
"""
from . import pointnet2_utils, pointnet2_modules  # noqa: F401
