"""Sample-sharded data parallelism: one process per GPU, full replica per GPU, the replay minibatch split by sample,
one NCCL all-reduce per optimiser phase over the flat gradient arenas (NVLink 5 / NVSwitch).

Replaces the reference's single-process ``nn.DataParallel`` (/root/reference/core/utils.py:202), which replicates the
encoders every forward and reduces gradients to GPU0.  Semantics kept from DataParallel: BatchNorm statistics are
per replica (no SyncBN).  Two all-reduces per step are required to keep the reference's ordering (the critic is
stepped before the actor forward, ddpg.py:160-174); each is ONE call on a contiguous arena.
"""
import os

import torch
import torch.distributed as dist


class World:
    def __init__(self, backend=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.size = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        self.backend = backend
        if self.size > 1 and not dist.is_initialized():
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            kw = {}
            if backend == "nccl":
                kw["device_id"] = torch.device("cuda", self.local_rank)
            dist.init_process_group(backend=backend, rank=self.rank, world_size=self.size, **kw)

    def all_reduce_mean(self, t, async_op=False):
        """Mean over ranks, in place.  ``async_op=True`` returns a handle for ``wait``: the collective runs on the process
        group's own stream (it starts once the work already enqueued on the current stream is done) and the current
        stream only blocks where ``wait`` is called — what the agents use to hide the reduce behind the SA1 backward."""
        if self.size == 1:
            return None if async_op else t
        if self.backend == "nccl":
            w = dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op)
            return (w, None) if async_op else t
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=async_op)   # gloo has no AVG
        if async_op:
            return (w, t)
        t.mul_(1.0 / self.size)
        return t

    def wait(self, handle):
        if handle is None:
            return
        w, scale_me = handle
        w.wait()
        if scale_me is not None:
            scale_me.mul_(1.0 / self.size)

    def all_reduce_max(self, t):
        if self.size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t

    def barrier(self):
        if self.size > 1:
            dist.barrier()

    def shard(self, n):
        """[lo, hi) sample range of this rank for a global minibatch of n (n % size == 0)."""
        assert n % self.size == 0
        per = n // self.size
        return self.rank * per, (self.rank + 1) * per

    def close(self):
        if self.size > 1 and dist.is_initialized():
            dist.destroy_process_group()
