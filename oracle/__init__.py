"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement of GA-DDPG's offline actor-critic update path, used as the
checker for the CUDA product in ``ga-ddpg_b200/``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import anything from here; the product never does.

Pinning status:
  * ``ddpg_cpu`` / ``nets_cpu`` / ``losses_cpu`` restate in-tree reference Python
    (core/networks.py, agent.py, ddpg.py, bc.py, loss.py, utils.py) and ARE pinned:
    ``oracle/make_golden.py`` runs the unmodified reference modules (imported from a
    scratch copy of /root/reference) on the same seeds and checks equality before it
    writes ``tests/golden/``.
  * ``pointnet2_cpu.c`` + ``pointnet2_ops_cpu`` restate the un-vendored third-party
    ``pointnet2_ops`` extension: PARITY UNPINNED (no reference test or fixture exists
    and the extension cannot be built offline); they follow SURVEY.md §8 Spec S1-S3.
"""
