"""Double-buffered feed loop: sampler -> ``update_parameters`` with the GPU never waiting for the host.

Plays the role of the inner loop of the reference's trainers: ``train_off_policy`` (/root/reference/core/
train_test_offline.py:117-127: ``memory.sample`` -> ``agent.update_parameters`` -> ``step_scheduler``, strictly one after
the other) and ``Trainer.train_iter`` (/root/reference/core/trainer.py:202-293), which already overlaps the two by asking
ray for the NEXT minibatch in the same ``ray.get`` as the current update (:215-228).  Here the overlap is done in-process:

* device-resident replay (``ReplayMemoryB200``): ``update_parameters(..., defer=True)`` only enqueues a step (index
  upload, gather, the captured graphs, one 64-byte read-back) and returns a ``PendingResult``; the loop enqueues step
  i+1 — including its host-side index draw — while the GPU still runs step i, and reads step i's scalars afterwards.
  The agent's host staging is a ring (agent.RING slots), so nothing a running step reads is overwritten.
* host replay (the reference's ``BaseMemory``: float64 numpy arrays): a worker thread draws the next minibatch and
  converts it to pinned float32 while the current step runs (numpy / torch release the GIL in those copies); the
  minibatch sequence is the sampler's own sequence (one thread samples, in order), so results equal the synchronous
  loop's bit for bit.

``gaps_us()`` reports, from CUDA events, how long the GPU sat idle between the end of one step and the first operation
of the next — the quantity the loop exists to drive to ~0 (tests/test_feed_gpu.py).
"""
import queue
import threading

import numpy as np
import torch

CLOUD_KEYS = ("point_state_batch", "next_point_state_batch")


class _HostPrefetcher(threading.Thread):
    """Worker: ``memory.sample(B)`` -> pinned float32 clouds, ``depth`` minibatches ahead."""

    def __init__(self, memory, batch_size, n, depth):
        super().__init__(daemon=True)
        self.memory, self.B, self.n, self.q = memory, batch_size, n, queue.Queue(maxsize=depth)
        self.pins = [dict() for _ in range(depth + 2)]   # one more than can be queued + the one being consumed
        self.err = None

    def run(self):
        try:
            for i in range(self.n):
                b = dict(self.memory.sample(self.B))
                pin = self.pins[i % len(self.pins)]
                for k in CLOUD_KEYS:
                    if k in b and b[k] is not None and not torch.is_tensor(b[k]):
                        a = np.asarray(b[k])
                        if k not in pin or pin[k].shape != a.shape:
                            pin[k] = torch.zeros(a.shape, dtype=torch.float32).pin_memory()
                        pin[k].copy_(torch.from_numpy(a))        # float64 -> float32 here, off the critical path
                        b[k] = pin[k]
                self.q.put(b)
        except BaseException as e:  # surfaced by the consumer
            self.err = e
            self.q.put(None)

    def next(self):
        b = self.q.get()
        if b is None:
            raise self.err
        return b


class FeedLoop:
    def __init__(self, agent, memory, batch_size, prefetch_depth=2, measure_gaps=False):
        from .replay_memory import ReplayMemoryB200

        self.agent, self.memory, self.B, self.depth = agent, memory, batch_size, prefetch_depth
        self.device_replay = isinstance(memory, ReplayMemoryB200)
        self.measure = measure_gaps
        self._starts, self._ends = [], []

    def train_iter(self, updates, on_loss=None, pipelined=True):
        """``updates`` update steps; returns the list of their loss dicts (the 11 keys of utils.py:1008-1020), in order.
        ``pipelined=False`` is the reference's synchronous order (sample, update, read the scalars, repeat)."""
        ag = self.agent
        if self.measure:
            ag.step_start_events, ag.step_end_events = self._starts, self._ends
        src = None
        if not self.device_replay and pipelined:
            src = _HostPrefetcher(self.memory, self.B, updates, self.depth)
            src.start()
        losses, pending = [], None

        def settle(p):
            r = p.result()
            losses.append(r)
            if on_loss is not None:
                on_loss(r)

        try:
            for i in range(updates):
                batch = src.next() if src is not None else self.memory.sample(self.B)
                h = ag.update_parameters(batch, ag.update_step, i, defer=pipelined)
                ag.step_scheduler(ag.update_step)
                if not pipelined:
                    losses.append(h)
                    if on_loss is not None:
                        on_loss(h)
                    continue
                if pending is not None:
                    settle(pending)          # step i-1: finished (or finishing) while step i is already queued
                pending = h
            if pending is not None:
                settle(pending)
        finally:
            if self.measure:
                ag.step_start_events = ag.step_end_events = None
        return losses

    def gaps_us(self):
        """GPU idle time between consecutive steps (end of step i-1 -> first operation of step i), microseconds."""
        torch.cuda.synchronize()
        n = min(len(self._starts), len(self._ends))
        return [1e3 * self._ends[i - 1].elapsed_time(self._starts[i]) for i in range(1, n)]
