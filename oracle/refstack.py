"""Load the UNMODIFIED reference Python on CPU (build-container only; test infrastructure).

Used by ``oracle/make_golden.py`` and ``tests/test_oracle_vs_reference.py`` to pin the
oracle's restatement of the in-tree reference code.  Nothing here runs on the GPU box:
``/root/reference`` does not exist there, ``available()`` returns False and callers skip.

What it does (SURVEY.md §8(c) "Can the reference's own implementation be imported"):
  * copies ``core/`` and ``experiments/`` to a scratch dir (importing
    ``experiments.config`` from the read-only tree would try to mkdir inside it,
    experiments/config.py:22-29);
  * registers stub modules for absent packages (IPython, matplotlib, GPUtil, easydict,
    transforms3d, tensorboardX) and ``oracle.pointnet2_ops_cpu`` as ``pointnet2_ops``;
  * maps the hard-coded ``.cuda()`` / ``"cuda"`` / ``torch.cuda.FloatTensor`` uses
    (agent.py:26,67-69,100-104,219-222; networks.py:329-337; loss.py:21-22,29-30;
    utils.py:968,980,991,1005) onto the CPU.
The reference source files themselves are imported byte-for-byte.
"""
import os
import shutil
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("GADDPG_REFERENCE_ROOT", "/root/reference")

_state = {"loaded": False, "dir": None}


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "core"))


class _EasyDict(dict):
    """Minimal attribute-dict (stands in for easydict.EasyDict)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(_EasyDict(x) if isinstance(x, dict) and not isinstance(x, _EasyDict) else x for x in v)
        super().__setattr__(k, v)
        super().__setitem__(k, v)

    __setitem__ = __setattr__


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    import numpy as np

    if "IPython" not in sys.modules:
        _stub("IPython", embed=lambda *a, **k: None)
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib", use=lambda *a, **k: None)
        mpl.pyplot = _stub("matplotlib.pyplot")
    if "GPUtil" not in sys.modules:
        _stub("GPUtil", getGPUs=lambda: [])
    if "easydict" not in sys.modules:
        _stub("easydict", EasyDict=_EasyDict)
    if "tensorboardX" not in sys.modules:
        _stub("tensorboardX", SummaryWriter=object)
    if "transforms3d" not in sys.modules:
        t3 = _stub("transforms3d")

        def _nope(*a, **k):
            raise NotImplementedError("transforms3d stub: env/pose helper outside the update path")

        names = {
            "quaternions": ["quat2mat", "mat2quat", "qmult", "qinverse", "quat2axangle", "axangle2quat"],
            "euler": ["euler2mat", "mat2euler", "euler2quat", "quat2euler"],
            "axangles": ["axangle2mat", "mat2axangle"],
        }
        for sub, fns in names.items():
            sm = _stub("transforms3d." + sub, **{f: _nope for f in fns})
            sm.__all__ = fns
            setattr(t3, sub, sm)
    # numpy>=1.24 dropped np.int / np.float; the reference's replay code uses np.int
    for alias, typ in (("int", int), ("float", float), ("bool", bool)):
        if not hasattr(np, alias):
            setattr(np, alias, typ)


def _patch_cuda_to_cpu():
    import torch

    if torch.cuda.is_available():
        return
    ident = lambda self, *a, **k: self  # noqa: E731
    torch.Tensor.cuda = ident
    torch.nn.Module.cuda = ident
    torch.cuda.FloatTensor = torch.FloatTensor

    def _is_cuda(d):
        return (isinstance(d, str) and d.startswith("cuda")) or (isinstance(d, torch.device) and d.type == "cuda")

    _t_to = torch.Tensor.to

    def tensor_to(self, *a, **k):
        a = tuple("cpu" if _is_cuda(x) else x for x in a)
        if _is_cuda(k.get("device")):
            k["device"] = "cpu"
        return _t_to(self, *a, **k)

    torch.Tensor.to = tensor_to
    _m_to = torch.nn.Module.to

    def module_to(self, *a, **k):
        a = tuple("cpu" if _is_cuda(x) else x for x in a)
        if _is_cuda(k.get("device")):
            k["device"] = "cpu"
        return _m_to(self, *a, **k)

    torch.nn.Module.to = module_to
    for fname in ("zeros", "ones", "tensor", "empty", "full", "eye", "arange"):
        orig = getattr(torch, fname)

        def wrapped(*a, __orig=orig, **k):
            if _is_cuda(k.get("device")):
                k["device"] = "cpu"
            return __orig(*a, **k)

        setattr(torch, fname, wrapped)


def load():
    """Import the reference; returns a namespace with the modules used by the update path."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if not _state["loaded"]:
        _install_stubs()
        _patch_cuda_to_cpu()
        from . import pointnet2_ops_cpu

        sys.modules["pointnet2_ops"] = pointnet2_ops_cpu
        sys.modules["pointnet2_ops.pointnet2_utils"] = pointnet2_ops_cpu.pointnet2_utils
        sys.modules["pointnet2_ops.pointnet2_modules"] = pointnet2_ops_cpu.pointnet2_modules
        d = tempfile.mkdtemp(prefix="gaddpg_refstack_")
        for sub in ("core", "experiments"):
            shutil.copytree(os.path.join(REFERENCE_ROOT, sub), os.path.join(d, sub))
        sys.path.insert(0, d)
        _state["dir"] = d
        _state["loaded"] = True
    import importlib

    ns = types.SimpleNamespace()
    ns.root = _state["dir"]
    ns.config = importlib.import_module("experiments.config")
    ns.utils = importlib.import_module("core.utils")
    ns.networks = importlib.import_module("core.networks")
    ns.loss = importlib.import_module("core.loss")
    ns.agent = importlib.import_module("core.agent")
    ns.ddpg = importlib.import_module("core.ddpg")
    ns.bc = importlib.import_module("core.bc")
    return ns


def make_reference_agent(policy="DDPG", extra_latent=1, seed=123456, overrides=None):
    """Build the reference agent exactly the way core/train_test_offline.py:305-349 does
    (seed, make_nets_opts_schedulers from the model-spec YAML, agent ctor, setup_feature_extractor)."""
    import copy

    import torch
    import yaml

    ns = load()
    cfg = ns.config.cfg
    ns.config.process_cfg()
    CONFIG = copy.deepcopy(cfg.RL_TRAIN)
    CONFIG.RL = policy == "DDPG"
    for k, v in (overrides or {}).items():
        setattr(CONFIG, k, v)
    if CONFIG.sa_channel_concat:
        CONFIG.value_model = True
    spec_path = os.path.join(ns.root, "experiments/model_spec/rl_pointnet_model_spec.yaml")
    spec = yaml.safe_load(open(spec_path))
    spec["state_feature_extractor"]["net_kwargs"]["extra_latent"] = extra_latent
    tmp_spec = os.path.join(ns.root, "model_spec_extra%d.yaml" % extra_latent)
    with open(tmp_spec, "w") as f:
        yaml.safe_dump(spec, f)
    torch.manual_seed(seed)
    net_dict = ns.utils.make_nets_opts_schedulers(tmp_spec, CONFIG)
    cls = ns.ddpg.DDPG if policy == "DDPG" else ns.bc.BC
    agent = cls(CONFIG.feature_input_dim, ns.utils.PandaTaskSpace6D(), CONFIG)
    agent.setup_feature_extractor(net_dict)
    return ns, agent, CONFIG
