"""CPU: the C-ABI library loads and exports every symbol include/gaddpg_b200.h declares; the ctypes struct mirrors
match the header; compute calls fail loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    from gaddpg_b200 import capi

    hdr = open(os.path.join(ROOT, "include", "gaddpg_b200.h")).read()
    declared = set(re.findall(r"GADDPG_API[^;(]*?\b(gaddpg_\w+)\s*\(", hdr))
    assert len(declared) >= 45 and declared == set(capi.lib.protos), declared ^ set(capi.lib.protos)
    lib = capi.lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert capi.lib.gaddpg_version() == 4
    assert b"sm_100a" in capi.lib.gaddpg_build_info()


def test_struct_mirrors_match_header():
    from gaddpg_b200 import structs

    structs.check_sizes()
    assert ctypes.sizeof(structs.NTGroup) == 4 * ctypes.sizeof(structs.NTProblem)


def _header_fields(name):
    """Member names of ``typedef struct <name> { ... } <name>;`` in declaration order."""
    hdr = open(os.path.join(ROOT, "include", "gaddpg_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, flags=re.S).group(1)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        first, *rest = decl.split(",")
        out.append(re.search(r"(\w+)(\[\w+\])?$", first.strip()).group(1))
        out += [re.search(r"(\w+)$", r.strip()).group(1) for r in rest]
    return out


def test_struct_mirror_field_order_matches_header():
    """sizeof() equality (above) does not catch two swapped members of the same size: compare the member names, in order, of every
    descriptor struct the Python side fills (Operand, BatchNorm tail, NT / TN problems, the per-level composites)."""
    from gaddpg_b200 import structs

    for cname, cls in (("gaddpg_operand", structs.Operand), ("gaddpg_bn_tail", structs.BNTail), ("gaddpg_nt_problem", structs.NTProblem),
                       ("gaddpg_tn_problem", structs.TNProblem), ("gaddpg_sa_layer", structs.SALayer), ("gaddpg_sa_level", structs.SALevel)):
        assert _header_fields(cname) == [f[0] for f in cls._fields_], cname


def test_opt_n_threads_rule_matches_oracle():
    from gaddpg_b200 import capi
    from oracle.pointnet2_ops_cpu import pointnet2_utils as U

    for n in (1, 2, 31, 32, 33, 500, 512, 513, 4096, 8192, 100000):
        assert capi.lib.gaddpg_opt_n_threads(n) == U.opt_n_threads(n)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from gaddpg_b200 import agent, capi, ops

    with pytest.raises(RuntimeError):
        capi.lib.gaddpg_device_info(None, None, None, None)
    with pytest.raises(RuntimeError):
        ops.furthest_point_sample(torch.zeros(1, 16, 3), 4)
    with pytest.raises(RuntimeError):
        agent.make_agent("DDPG")


def test_config_defaults_match_oracle():
    from gaddpg_b200.config import DEFAULTS, LOSS_KEYS
    from oracle import ddpg_cpu

    assert LOSS_KEYS == ddpg_cpu.LOSS_KEYS
    for k, v in ddpg_cpu.DEFAULTS.items():
        assert DEFAULTS[k] == v, k
