"""Device-resident replay buffer: ``ReplayMemoryB200`` — drop-in for the reference's ``BaseMemory`` on the training
side.

Same surface as /root/reference/core/replay_memory.py (``push`` :178-207, ``add_episode`` :209-232, ``sample`` :166-176,
``post_process_batch`` :251-272, ``upper_idx`` :129-130, ``reset``, ``__len__``, ``recompute_return_with_gamma``
:152-164, ``save`` / ``load`` :274-357) and the same dict out of ``sample`` (the keys ``Agent.prepare_data`` reads,
agent.py:211-240) — but the clouds are stored ONCE, as float32, in HBM, and a minibatch is assembled by one CUDA
gather (`gaddpg_replay_gather`, csrc/replay.cu) instead of numpy fancy indexing of float64 arrays + conversion +
pageable H2D.  The values are the ones the reference feeds its networks: ``torch.cuda.FloatTensor(float64 ndarray)``
rounds to float32 at sample time, this buffer rounds the same numbers at push time.

What stays on the host (numpy, as in the reference): the small per-transition arrays (they are also mirrored in the
device record table), the ring bookkeeping (``cur_idx``, ``is_full``, ``episode_map``) and the index draw — it uses
numpy's global random state exactly like ``BaseMemory.sample`` so ``np.random.seed`` reproduces the reference's
minibatches.  A 180 GB B200 holds 10.9 M transitions of the default 1030-column clouds (16.5 KB each).

Default configuration only (``use_image = False``, ``self_supervision = False``; experiments/config.py:105,113).
"""
import os
import weakref

import numpy as np
import torch

from .capi import current_stream, lib

# columns of the float32 record table (one 128-byte row per transition)
REC_W = 32
C_ACTION, C_EXPERT_ACTION, C_GOAL = 0, 6, 12
C_REWARD, C_RETURN, C_TERMINAL, C_TIMESTEP, C_EXPERT, C_PERTURB, C_COLLIDE, C_GRASP, C_TARGET = 19, 20, 21, 22, 23, 24, 25, 26, 27
_SCALARS = (("reward", C_REWARD), ("returns", C_RETURN), ("terminal", C_TERMINAL), ("timestep", C_TIMESTEP),
            ("expert_flags", C_EXPERT), ("perturb_flags", C_PERTURB), ("collide", C_COLLIDE), ("grasp", C_GRASP),
            ("target_idx", C_TARGET))
ATTR_NAMES = ["action", "pose", "point_state", "target_idx", "reward", "terminal", "timestep", "returns", "state_pose",
              "image_state", "collide", "grasp", "perturb_flags", "goal", "expert_flags", "expert_action"]


def replay_config(args=None, uniform_num_pts=None, episode_max_len=None, gamma=None, buffer_start_idx=None, RL=None,
                  save_data_name=None):
    """The fields ``BaseMemory.__init__`` takes from the reference's cfg (replay_memory.py:28-32: every ``RL_TRAIN`` key
    becomes an attribute, plus ``RL_MAX_STEP`` and ``RL_SAVE_DATA_NAME``); explicit keyword arguments win, and without a
    cfg the defaults of experiments/config.py apply.  Unsupported switches fail loudly instead of being ignored."""
    tr = getattr(args, "RL_TRAIN", None) if args is not None else None
    has = lambda key: tr is not None and key in tr  # noqa: E731
    pick = lambda v, key, default: v if v is not None else (tr[key] if has(key) else default)  # noqa: E731
    if has("use_image") and tr["use_image"]:
        raise NotImplementedError("ReplayMemoryB200: use_image = True (image observations) is outside the point-cloud update path")
    if has("self_supervision") and tr["self_supervision"]:
        raise NotImplementedError("ReplayMemoryB200: self_supervision (set_onpolicy_goal) is not built")
    return dict(uniform_num_pts=int(pick(uniform_num_pts, "uniform_num_pts", 1024)), gamma=float(pick(gamma, "gamma", 0.95)),
                buffer_start_idx=int(pick(buffer_start_idx, "buffer_start_idx", 0)), RL=bool(pick(RL, "RL", True)),
                episode_max_len=int(episode_max_len if episode_max_len is not None else getattr(args, "RL_MAX_STEP", 20)),
                save_data_name=save_data_name if save_data_name is not None else getattr(args, "RL_SAVE_DATA_NAME", "data_buffer.npz"))


class ReplayBatch(dict):
    """What ``ReplayMemoryB200.sample`` returns: the reference's minibatch dict, assembled lazily.

    Handed straight to ``DDPGB200/BCB200.update_parameters`` it is never materialised: ``prepare_data`` asks the memory to
    gather the sampled rows directly into the agent's input buffers (one launch pair, no intermediate copy, no
    per-field copies).  Any dict access (``batch["point_state_batch"]``, ``keys()``, ``items()`` ...) materialises the
    full dict of device tensors first, so code written against ``BaseMemory.sample`` keeps working."""

    def __init__(self, memory, batch_idx):
        super().__init__()
        self.memory, self.batch_idx, self.materialised = memory, np.asarray(batch_idx), False
        reg = getattr(memory, "_lazy", None)
        if reg is not None:
            reg[id(self)] = weakref.ref(self)     # a write to the buffer snapshots this batch first (BaseMemory.sample copies at sample time)

    def materialise(self, snapshot=False):
        if not self.materialised:
            self.materialised = True
            getattr(self.memory, "_lazy", {}).pop(id(self), None)
            d = self.memory.gather(self.batch_idx)
            if snapshot:   # the per-batch-size output buffers are reused by the next gather: keep private copies
                d = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in d.items()}
            dict.update(self, d)
        return self

    def __getitem__(self, k):
        if not dict.__contains__(self, k):      # keys the caller attached itself do not force the gather
            self.materialise()
        return dict.__getitem__(self, k)

    def __contains__(self, k):
        return dict.__contains__(self.materialise(), k)

    def __iter__(self):
        return dict.__iter__(self.materialise())

    def __len__(self):
        return dict.__len__(self.materialise())

    def get(self, k, default=None):
        return dict.get(self.materialise(), k, default)

    def keys(self):
        return dict.keys(self.materialise())

    def values(self):
        return dict.values(self.materialise())

    def items(self):
        return dict.items(self.materialise())


class ReplayMemoryB200:
    def __init__(self, buffer_size, args=None, name="expert", device=None, uniform_num_pts=None, episode_max_len=None,
                 gamma=None, buffer_start_idx=None, RL=None, channels=4, save_data_name=None):
        """``args`` may be the reference's cfg (RL_TRAIN / RL_MAX_STEP / RL_SAVE_DATA_NAME, replay_memory.py:20-32);
        explicit keyword arguments win over it."""
        if not torch.cuda.is_available():
            raise RuntimeError("ReplayMemoryB200 keeps the replay in GPU memory: no CUDA device, no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.cur_idx, self.total_env_step, self.is_full, self.name = 0, 0, False, name
        c = replay_config(args, uniform_num_pts=uniform_num_pts, episode_max_len=episode_max_len, gamma=gamma,
                          buffer_start_idx=buffer_start_idx, RL=RL, save_data_name=save_data_name)
        for k, v in c.items():
            setattr(self, k, v)
        self.buffer_size, self.channels = int(buffer_size), channels
        self.attr_names = ATTR_NAMES[:]
        self.init_buffer()

    # ---- storage (replay_memory.py:359-384) ---------------------------------------------------------------
    def init_buffer(self):
        n, dev = self.buffer_size, self.device
        self.row = (self.channels, self.uniform_num_pts + 6)
        self.point_state = torch.zeros((n,) + self.row, dtype=torch.float32, device=dev)       # HBM, the only copy
        self.records = torch.zeros(n, REC_W, dtype=torch.float32, device=dev)
        self.episode_map_dev = torch.zeros(n, dtype=torch.int32, device=dev)
        self._rec_host = torch.zeros(n, REC_W, dtype=torch.float32).pin_memory()
        self._rec = self._rec_host.numpy()
        self._emap_host = torch.zeros(n, dtype=torch.int32).pin_memory()
        self.episode_map = np.zeros((n,), dtype=np.uint32)
        # numpy views with the reference's attribute names (all float32 there too)
        self.action = self._rec[:, C_ACTION:C_ACTION + 6]
        self.expert_action = self._rec[:, C_EXPERT_ACTION:C_EXPERT_ACTION + 6]
        self.goal = self._rec[:, C_GOAL:C_GOAL + 7]
        for nm, c in _SCALARS:
            setattr(self, nm, self._rec[:, c])
        self.pose = np.zeros((n, 64), dtype=np.float32)
        self.state_pose = np.zeros((n, 4, 4), dtype=np.float32)
        self.image_state = np.zeros((n, 1), dtype=np.uint16)
        self._dirty = []       # disjoint [lo, hi) ranges of host record / episode-map rows not yet mirrored to the device
        self._out = {}
        self._lazy = {}                  # id -> weakref of ReplayBatch objects handed out by sample() and not gathered yet

    def upper_idx(self):
        return max(self.cur_idx, 1) if not self.is_full else self.buffer_size

    def __len__(self):
        return self.upper_idx()

    def get_cur_idx(self):
        return self.cur_idx

    def get_total_env_step(self):
        return self.total_env_step

    def reset(self):
        self.cur_idx, self.is_full = 0, False

    def _mark(self, lo, hi):
        """Add [lo, hi) to the dirty set, kept as a few disjoint ranges (a ring that wraps dirties its tail and its head,
        not the whole table)."""
        r = sorted(self._dirty + [(int(lo), int(hi))])
        out = [r[0]]
        for a, b in r[1:]:
            if a <= out[-1][1]:
                out[-1] = (out[-1][0], max(out[-1][1], b))
            else:
                out.append((a, b))
        while len(out) > 4:      # bound the bookkeeping: fuse the two closest ranges
            k = min(range(len(out) - 1), key=lambda i: out[i + 1][0] - out[i][1])
            out[k: k + 2] = [(out[k][0], out[k + 1][1])]
        self._dirty = out

    def _settle(self):
        """Called before the buffer is written: minibatches that were sampled but not consumed yet are gathered NOW, so
        they keep the contents they had at sample time (replay_memory.py:166-176 copies in ``sample``)."""
        if self._lazy:
            for ref in list(self._lazy.values()):
                b = ref()
                if b is not None:
                    b.materialise(snapshot=True)
            self._lazy.clear()

    def mark_dirty(self, lo=0, hi=None):
        """Call after editing the host-side arrays (``returns``, ``episode_map``, ``reward`` ...) in place for slots
        [lo, hi): the next ``sample`` mirrors them to the device record table.  ``push`` / ``add_episode`` / ``load`` /
        ``recompute_return_with_gamma`` do this themselves."""
        self._settle()
        self._mark(lo, self.buffer_size if hi is None else hi)

    def _flush(self):
        """Mirror the host-side records / episode map of the slots touched since the last sample to the device."""
        for lo, hi in self._dirty:
            self._emap_host.numpy()[lo:hi] = self.episode_map[lo:hi].astype(np.int64).astype(np.int32)  # keep the uint32 bits
            # blocking copies: the host arrays are live (the next push writes into them), the ranges are small
            self.records[lo:hi].copy_(self._rec_host[lo:hi])
            self.episode_map_dev[lo:hi].copy_(self._emap_host[lo:hi])
        self._dirty = []

    # ---- writers (replay_memory.py:178-232) -----------------------------------------------------------------
    def push(self, step_dict, _pending=None):
        self._settle()
        store_idx = self.cur_idx % self.buffer_size
        ps = np.asarray(step_dict["point_state"])
        if ps.shape[1] < 100 or ps.sum() == 0:
            return
        if ps.shape != self.row:
            raise ValueError("point_state has shape %s, the buffer stores %s" % (ps.shape, self.row))
        for name in self.attr_names:
            if name in ("image_state", "point_state") or name not in step_dict:
                continue
            getattr(self, name)[store_idx] = step_dict[name]
        # float64 -> float32 here (the reference converts the same values at sample time, agent.py:221-222)
        ps32 = np.ascontiguousarray(ps, dtype=np.float32)
        if _pending is None:
            self.point_state[store_idx].copy_(torch.from_numpy(ps32), non_blocking=False)
        else:
            _pending.append((store_idx, ps32))       # add_episode uploads runs of consecutive slots with one copy each
        self._mark(store_idx, store_idx + 1)
        if self.cur_idx >= self.buffer_size - 1:
            self.is_full = True
        self.cur_idx += 1
        self.total_env_step += 1
        if self.cur_idx >= self.buffer_size or self.cur_idx < self.buffer_start_idx:
            self.cur_idx = self.buffer_start_idx

    def add_episode(self, episode, explore=False, test=False):
        n = len(episode)
        if (not self.RL) and episode[-1]["reward"] < 0.5 and not explore:
            return
        pending = []
        for transition in episode:
            self.push(transition, pending)
        i = 0
        while i < len(pending):                      # one H2D per run of consecutive slots (two when the ring wraps)
            j = i
            while j + 1 < len(pending) and pending[j + 1][0] == pending[j][0] + 1:
                j += 1
            self.point_state[pending[i][0]: pending[j][0] + 1].copy_(torch.from_numpy(np.stack([c for _, c in pending[i: j + 1]])))
            i = j + 1
        if self.cur_idx - n >= 0 and n > 0:
            cost_to_go = 0
            for i in range(n):
                self.returns[self.cur_idx - i - 1] = self.reward[self.cur_idx - i - 1] + self.gamma ** i * cost_to_go
                cost_to_go = self.returns[self.cur_idx - 1 - i]
            self.episode_map[self.cur_idx - n: self.cur_idx] = self.cur_idx - 1
            self._mark(self.cur_idx - n, self.cur_idx)

    def recompute_return_with_gamma(self):
        self._settle()
        ends = np.sort(np.unique(self.episode_map))
        out = self.returns.copy()
        for k in range(len(ends) - 1):
            start, end = int(ends[k]), int(ends[k + 1])
            cost_to_go = 0
            for i in range(end - start):
                cur = end + 1
                out[cur - i - 1] = self.reward[cur - i - 1] + self.gamma ** i * cost_to_go
                cost_to_go = out[cur - i - 1]
        self.returns[:] = out
        self._mark(0, self.buffer_size)

    # ---- sampling (replay_memory.py:166-176, 109-127, 251-272) -------------------------------------------------
    def draw_indices(self, batch_size):
        batch_idx = np.random.randint(self.episode_max_len, self.upper_idx(), batch_size)
        np.random.shuffle(batch_idx)
        return batch_idx

    def _buffers(self, B):
        o = self._out.get(B)
        if o is None:
            dev = self.device
            o = dict(state=torch.empty((B,) + self.row, dtype=torch.float32, device=dev),
                     next=torch.empty((B,) + self.row, dtype=torch.float32, device=dev),
                     rec=torch.empty(B, 2 * REC_W, dtype=torch.float32, device=dev),
                     inc=torch.empty(B, dtype=torch.int32, device=dev), idx=torch.empty(B, dtype=torch.int32, device=dev),
                     idx_host=torch.empty(max(B, 1), dtype=torch.int32).pin_memory()[:B], copied=torch.cuda.Event())
            self._out[B] = o
        return o

    def gather(self, batch_idx):
        """One minibatch for the given slots: a dict of DEVICE tensors under the reference's keys.  The tensors are
        views of per-batch-size output buffers that the next ``gather`` of the same size overwrites (the update step
        consumes them before that)."""
        batch_idx = np.asarray(batch_idx)
        B = int(batch_idx.shape[0])
        if B and (batch_idx.min() < 0 or batch_idx.max() >= self.buffer_size):
            raise IndexError("replay index out of range [0, %d)" % self.buffer_size)
        self._flush()
        o = self._buffers(B)
        with torch.cuda.device(self.device):
            self._upload_indices(o, batch_idx)
        row_floats = self.row[0] * self.row[1]
        if B:
          with torch.cuda.device(self.device):
            lib.gaddpg_replay_gather(self.point_state.data_ptr(), row_floats, self.records.data_ptr(), REC_W, C_TIMESTEP,
                                     self.episode_map_dev.data_ptr(), self.buffer_size, o["idx"].data_ptr(), B, o["state"].data_ptr(),
                                     o["next"].data_ptr(), o["rec"].data_ptr(), o["inc"].data_ptr(), None, None, current_stream())
        r, n = o["rec"][:, :REC_W], o["rec"][:, REC_W:]
        data = {
            "point_state_batch": o["state"], "next_point_state_batch": o["next"],
            "action_batch": r[:, C_ACTION:C_ACTION + 6], "expert_action_batch": r[:, C_EXPERT_ACTION:C_EXPERT_ACTION + 6],
            "goal_batch": r[:, C_GOAL:C_GOAL + 7], "reward_batch": r[:, C_REWARD], "return_batch": r[:, C_RETURN],
            "mask_batch": r[:, C_TERMINAL], "time_batch": r[:, C_TIMESTEP], "expert_flag_batch": r[:, C_EXPERT],
            "perturb_flag_batch": r[:, C_PERTURB], "collide_batch": r[:, C_COLLIDE], "grasp_batch": r[:, C_GRASP],
            "next_action_batch": n[:, C_ACTION:C_ACTION + 6], "next_expert_action_batch": n[:, C_EXPERT_ACTION:C_EXPERT_ACTION + 6],
            "next_goal_batch": n[:, C_GOAL:C_GOAL + 7], "next_return_batch": n[:, C_RETURN],
            # host-side leftovers of the reference dict that the update never reads (agent.py:211-240 keeps them as-is)
            "image_state_batch": self.image_state[batch_idx].astype(np.float32), "next_image_state_batch": None,
            "state_pose_batch": self.state_pose[batch_idx], "grasp_sample_batch": np.zeros([0, 4, 4]),
            "batch_idx": np.uint8(batch_idx), "increment_idx": o["inc"],
        }
        return data

    def sample(self, batch_size):
        return ReplayBatch(self, self.draw_indices(batch_size))

    def _upload_indices(self, o, batch_idx):
        o["copied"].synchronize()          # the previous minibatch's index upload must have left the pinned buffer
        o["idx_host"].numpy()[:] = batch_idx
        o["idx"].copy_(o["idx_host"], non_blocking=True)
        o["copied"].record()

    def gather_into(self, batch_idx, cloud, next_cloud, vec, vec_layout):
        """Gather a minibatch straight into an agent's input buffers: ``cloud`` / ``next_cloud`` (B, C, N+6) and the
        field-major vector ``vec`` whose layout is ``vec_layout`` = {field: (offset, width)} with the agent's field
        names (action, expert_action, goal, reward, ret, done, time, expert_flag, perturb_flag)."""
        batch_idx = np.asarray(batch_idx)
        B = int(batch_idx.shape[0])
        if tuple(cloud.shape) != (B,) + self.row or (next_cloud is not None and tuple(next_cloud.shape) != (B,) + self.row):
            raise ValueError("agent cloud buffers %s do not match the stored rows %s" % (tuple(cloud.shape), (B,) + self.row))
        if B and (batch_idx.min() < 0 or batch_idx.max() >= self.buffer_size):
            raise IndexError("replay index out of range [0, %d)" % self.buffer_size)
        self._flush()
        o = self._buffers(B)
        key = tuple(sorted(vec_layout.items()))
        if o.get("soa_key") != key:
            cols = dict(action=(C_ACTION, 6), expert_action=(C_EXPERT_ACTION, 6), goal=(C_GOAL, 7), reward=(C_REWARD, 1),
                        ret=(C_RETURN, 1), done=(C_TERMINAL, 1), time=(C_TIMESTEP, 1), expert_flag=(C_EXPERT, 1),
                        perturb_flag=(C_PERTURB, 1))
            m = np.full((2, REC_W), -1, dtype=np.int32)
            for name, (c0, w) in cols.items():
                if name in vec_layout:
                    off, width = vec_layout[name]
                    assert width == w, (name, width, w)
                    m[0, c0:c0 + w] = off + np.arange(w)
                    m[1, c0:c0 + w] = w
            o["soa_map"], o["soa_key"] = torch.from_numpy(m).to(self.device), key
        with torch.cuda.device(self.device):
            self._upload_indices(o, batch_idx)
        if B:
          with torch.cuda.device(self.device):
            nxt = next_cloud if next_cloud is not None else o["next"]
            lib.gaddpg_replay_gather(self.point_state.data_ptr(), self.row[0] * self.row[1], self.records.data_ptr(), REC_W, C_TIMESTEP,
                                     self.episode_map_dev.data_ptr(), self.buffer_size, o["idx"].data_ptr(), B, cloud.data_ptr(),
                                     nxt.data_ptr(), None, o["inc"].data_ptr(), o["soa_map"].data_ptr(), vec.data_ptr(),
                                     current_stream())

    # ---- persistence (replay_memory.py:274-357) ---------------------------------------------------------------
    def save(self, save_dir="."):
        """Same ``np.savez`` file as the reference (float64 ``point_state``), so either implementation loads it."""
        os.makedirs(save_dir, exist_ok=True)
        d = {name: np.array(getattr(self, name)) for name in self.attr_names if name != "point_state"}
        d["point_state"] = self.point_state.cpu().numpy().astype(np.float64)
        d.update(episode_map=self.episode_map, is_full=self.is_full, cur_idx=self.cur_idx, total_env_step=self.total_env_step)
        np.savez(os.path.join(save_dir, self.save_data_name), **d)

    def load(self, data_dir, buffer_size=100000, trusted=False, **kwargs):
        """replay_memory.py:316-357.  The file holds numeric arrays only, so pickled objects are refused unless the caller
        vouches for the file (``trusted=True``: the reference itself loads with ``allow_pickle=True``)."""
        path = os.path.join(data_dir, self.save_data_name)
        if not os.path.exists(path):
            return
        self._settle()
        data = np.load(path, allow_pickle=bool(trusted), mmap_mode="r")
        n = int(np.amax(data["episode_map"]))
        for name in self.attr_names + ["episode_map"]:
            if name == "image_state" or name not in data:
                continue
            if name == "point_state":
                self.point_state[:n].copy_(torch.from_numpy(np.ascontiguousarray(data[name][:n], dtype=np.float32)))
            else:
                getattr(self, name)[:n] = data[name][:n]
        self.cur_idx = n
        self.total_env_step = int(data["total_env_step"])
        self.is_full = bool(data["is_full"]) and self.cur_idx >= self.buffer_size - 1
        self.cur_idx = self.upper_idx()
        self.recompute_return_with_gamma()
        self._mark(0, self.buffer_size)
