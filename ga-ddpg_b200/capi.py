"""ctypes binding of libgaddpg_b200.so, generated from include/gaddpg_b200.h at import time.

The prototypes are parsed from the header so the Python side can never drift from the C ABI; every
declared symbol must resolve (tests/test_capi_symbols.py).  There is NO fallback: if the library is
missing it is built with nvcc; if that fails, or a call returns an error code, a RuntimeError is raised.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "gaddpg_b200.h")

_CT = {
    "int": ctypes.c_int,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "long long": ctypes.c_longlong,
    "void": None,
    "const char*": ctypes.c_char_p,
}


def _ctype(t: str):
    t = " ".join(t.replace("*", " * ").split()).replace(" *", "*")
    if t in _CT:
        return _CT[t]
    if t.endswith("*"):
        return ctypes.c_void_p  # every pointer is an opaque device (or host out-) pointer
    raise ValueError("unsupported C type in header: %r" % t)


def parse_header(path: str = HEADER):
    """-> {name: (restype, [(argtype, argname), ...])} for every GADDPG_API prototype."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"GADDPG_API\s+([\w\s\*]+?)\s*\b(gaddpg_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        arglist = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.+?)(\w+)$", a)
                arglist.append((mm.group(1).strip(), mm.group(2)))
        protos[name] = (ret, arglist)
    return protos


class _Lib:
    def __init__(self):
        self._lib = None
        self.protos = parse_header()

    def load(self):
        if self._lib is None:
            from . import build as _build

            if os.environ.get("GADDPG_LIB"):       # kernel experiments: load an explicitly named build
                path = os.environ["GADDPG_LIB"]
            else:
                path = _build.LIB if os.path.exists(_build.LIB) and os.environ.get("GADDPG_NO_REBUILD") else _build.build()
            lib = ctypes.CDLL(path)
            for name, (ret, args) in self.protos.items():
                fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
                fn.restype = _ctype(ret)
                fn.argtypes = [_ctype(t) for t, _ in args]
            self._lib = lib
        return self._lib

    def __getattr__(self, name):
        lib = self.load()
        fn = getattr(lib, name)
        if self.protos[name][0] != "int" or name in ("gaddpg_version", "gaddpg_opt_n_threads", "gaddpg_launch_count", "gaddpg_get_tensor_core",
                                                   "gaddpg_gemm_nt_path", "gaddpg_sa1_fused_grid"):
            return fn

        def checked(*a):
            prof = PROFILE
            if prof is not None:
                return prof.call(name, fn, a, lib)
            rc = fn(*a)
            if rc != 0:
                raise RuntimeError("%s failed (%d): %s" % (name, rc, lib.gaddpg_last_error().decode()))
            return rc

        checked.__name__ = name
        setattr(self, name, checked)
        return checked


lib = _Lib()
PROFILE = None  # set to a profiler.KernelProfile to bracket every C-ABI call with CUDA events (bench.py roofline leg)


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream():
    import torch

    return torch.cuda.current_stream().cuda_stream
