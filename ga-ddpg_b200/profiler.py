"""Per-entry-point device timing for the roofline leg of bench.py: every C-ABI call made while a ``KernelProfile``
is installed is bracketed by CUDA events on the launching stream (eager mode only — graphs are bypassed), and the
caller may tag the call with its algorithmic FLOPs / bytes.  Nothing here runs in the timed benchmark region."""
from collections import OrderedDict

import torch

from . import capi


class KernelProfile:
    def __init__(self):
        self.records = []   # (label, ev0, ev1, meta)
        self.tag = None     # set by engine helpers right before the call they describe

    def call(self, name, fn, args, lib):
        stream = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        rc = fn(*args)
        e1.record(stream)
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (name, rc, lib.gaddpg_last_error().decode()))
        tag, self.tag = self.tag, None
        self.records.append((name, e0, e1, tag))
        return rc

    def __enter__(self):
        capi.PROFILE = self
        return self

    def __exit__(self, *exc):
        capi.PROFILE = None
        torch.cuda.synchronize()

    def table(self, resolve_m=None):
        """-> OrderedDict label -> dict(ms, calls, flops, bytes); resolve_m maps an M_dev pointer to its live value."""
        out = OrderedDict()
        for name, e0, e1, tag in self.records:
            label = name.replace("gaddpg_", "")
            flops = nbytes = 0.0
            if tag:
                label += ":" + tag.get("kind", "")
                M = tag.get("M_max", 0)
                if tag.get("M_dev") and resolve_m:
                    M = resolve_m.get(tag["M_dev"], M)
                flops = tag.get("flops_per_row", 0.0) * M
                nbytes = tag.get("bytes_per_row", 0.0) * M + tag.get("bytes_fixed", 0.0)
            r = out.setdefault(label, dict(ms=0.0, calls=0, flops=0.0, bytes=0.0))
            r["ms"] += e0.elapsed_time(e1)
            r["calls"] += 1
            r["flops"] += flops
            r["bytes"] += nbytes
        return out
