"""CPU restatement of the replay buffer's producer side of the update path (oracle; TEST INFRASTRUCTURE ONLY).

Follows /root/reference/core/replay_memory.py (class BaseMemory) for what feeds ``Agent.update_parameters``:
``init_buffer`` :359-384, ``push`` :178-207, ``add_episode`` :209-232, ``upper_idx`` :129-130, ``__getitem__`` :109-127,
``sample`` :166-176, ``post_process_batch`` :251-272, ``recompute_return_with_gamma`` :152-164, ``reset`` :148-150.
Default configuration only (``use_image = False``, ``self_supervision = False``: experiments/config.py:105,113), i.e.
``set_onpolicy_goal`` is not restated.  Plain numpy, same dtypes as the reference (float64 clouds, float32 records,
uint32 episode map), same use of numpy's GLOBAL random state in ``sample`` so that ``np.random.seed`` reproduces the
reference's minibatch indices.

Pinned: ``oracle/make_golden.py::replay_fixture`` feeds identical synthetic episodes to the UNMODIFIED reference class
and to this one and requires every key of every sampled minibatch to be bit-identical before it writes
``tests/golden/replay_n128.npz``.
"""
import numpy as np

ATTR_NAMES = ["action", "pose", "point_state", "target_idx", "reward", "terminal", "timestep", "returns", "state_pose",
              "image_state", "collide", "grasp", "perturb_flags", "goal", "expert_flags", "expert_action"]


class OracleMemory:
    def __init__(self, buffer_size, uniform_num_pts=1024, episode_max_len=20, gamma=0.95, buffer_start_idx=0, RL=True,
                 name="expert"):
        self.cur_idx, self.total_env_step, self.is_full, self.name = 0, 0, False, name
        self.buffer_size, self.uniform_num_pts, self.episode_max_len = buffer_size, uniform_num_pts, episode_max_len
        self.gamma, self.buffer_start_idx, self.RL = gamma, buffer_start_idx, RL
        self.init_buffer()

    def init_buffer(self):                                                     # :359-384 (use_image = False)
        n = self.buffer_size
        self.image_state = np.zeros((n, 1), dtype=np.uint16)
        self.action = np.zeros((n, 6), dtype=np.float32)
        self.expert_action = np.zeros((n, 6), dtype=np.float32)
        self.terminal = np.zeros((n,), dtype=np.float32)
        self.timestep = np.zeros((n,), dtype=np.float32)
        self.reward = np.zeros((n,), dtype=np.float32)
        self.returns = np.zeros((n,), dtype=np.float32)
        self.pose = np.zeros((n, 64), dtype=np.float32)
        self.point_state = np.zeros([n, 4, self.uniform_num_pts + 6])
        self.collide = np.zeros((n,), dtype=np.float32)
        self.grasp = np.zeros((n,), dtype=np.float32)
        self.state_pose = np.zeros((n, 4, 4), dtype=np.float32)
        self.target_idx = np.zeros((n,), dtype=np.float32)
        self.goal = np.zeros((n, 7), dtype=np.float32)
        self.episode_map = np.zeros((n,), dtype=np.uint32)
        self.expert_flags = np.zeros((n,), dtype=np.float32)
        self.perturb_flags = np.zeros((n,), dtype=np.float32)

    def upper_idx(self):                                                       # :129-130
        return max(self.cur_idx, 1) if not self.is_full else len(self.point_state)

    def __len__(self):
        return self.upper_idx()

    def reset(self):                                                           # :148-150
        self.cur_idx, self.is_full = 0, False

    def push(self, step_dict):                                                 # :178-207
        store_idx = self.cur_idx % len(self.point_state)
        if step_dict["point_state"].shape[1] < 100 or step_dict["point_state"].sum() == 0:
            return
        for name in ATTR_NAMES:
            if name == "image_state":
                continue                                                       # use_image = False
            if name in step_dict:
                getattr(self, name)[store_idx] = step_dict[name]
        if self.cur_idx >= len(self.episode_map) - 1:
            self.is_full = True
        self.cur_idx += 1
        self.total_env_step += 1
        if self.cur_idx >= len(self.point_state) or self.cur_idx < self.buffer_start_idx:
            self.cur_idx = self.buffer_start_idx

    def add_episode(self, episode, explore=False, test=False):                 # :209-232 (reward bookkeeping omitted)
        n = len(episode)
        if (not self.RL) and episode[-1]["reward"] < 0.5 and not explore:
            return
        for transition in episode:
            self.push(transition)
        if self.cur_idx - n >= 0 and n > 0:
            cost_to_go = 0
            for i in range(n):
                self.returns[self.cur_idx - i - 1] = self.reward[self.cur_idx - i - 1] + self.gamma ** i * cost_to_go
                cost_to_go = self.returns[self.cur_idx - 1 - i]
            self.episode_map[self.cur_idx - n: self.cur_idx] = self.cur_idx - 1

    def recompute_return_with_gamma(self):                                     # :152-164
        ends = np.sort(np.unique(self.episode_map))
        out = self.returns.copy()
        for k in range(len(ends) - 1):
            start, end = ends[k], ends[k + 1]
            cost_to_go = 0
            for i in range(end - start):
                cur = end + 1
                out[cur - i - 1] = self.reward[cur - i - 1] + self.gamma ** i * cost_to_go
                cost_to_go = out[cur - i - 1]
        self.returns = out

    def save(self, save_dir=".", save_data_name="data_buffer.npz"):            # :336-356
        import os
        os.makedirs(save_dir, exist_ok=True)
        d = {name: getattr(self, name) for name in ATTR_NAMES + ["episode_map", "is_full", "cur_idx", "total_env_step", "target_idx"]}
        np.savez(os.path.join(save_dir, save_data_name), **d)

    def load(self, data_dir, save_data_name="data_buffer.npz"):                # :274-334 (use_image = False)
        import os
        path = os.path.join(data_dir, save_data_name)
        if not os.path.exists(path):
            return
        data = np.load(path, allow_pickle=True, mmap_mode="r")
        n = np.amax(data["episode_map"])           # sic: the slice [:n] leaves the last stored transition out
        for name in ATTR_NAMES + ["episode_map", "target_idx"]:
            if name == "image_state" or name not in data:
                continue
            getattr(self, name)[:n] = data[name][:n]
        self.cur_idx = n
        self.total_env_step = int(data["total_env_step"])
        self.is_full = bool(data["is_full"]) and self.cur_idx >= self.buffer_size - 1
        self.cur_idx = self.upper_idx()
        self.recompute_return_with_gamma()

    def draw_indices(self, batch_size):                                        # :169-172
        batch_idx = np.random.randint(self.episode_max_len, self.upper_idx(), batch_size)
        np.random.shuffle(batch_idx)
        return batch_idx

    def gather(self, batch_idx):                                               # :109-127 + :251-272
        inc = np.minimum(self.episode_map[batch_idx], batch_idx + 1).astype(int)
        f32 = np.float32
        data = {
            "image_state_batch": self.image_state[batch_idx].astype(f32).copy(),  # process_image_output, 2-D => cast only
            "expert_action_batch": f32(self.expert_action[batch_idx]),
            "action_batch": f32(self.action[batch_idx]),
            "reward_batch": f32(self.reward[batch_idx]),
            "return_batch": f32(self.returns[batch_idx]),
            "mask_batch": f32(self.terminal[batch_idx]),
            "time_batch": f32(self.timestep[batch_idx]),
            "state_pose_batch": f32(self.state_pose[batch_idx]),
            "collide_batch": f32(self.collide[batch_idx]),
            "grasp_batch": f32(self.grasp[batch_idx]),
            "goal_batch": f32(self.goal[batch_idx]),
        }
        data["grasp_sample_batch"] = np.zeros([0, 4, 4])
        data["next_image_state_batch"] = self.image_state[inc].astype(f32).copy()
        data["next_goal_batch"] = f32(self.goal[inc])
        data["next_expert_action_batch"] = f32(self.expert_action[inc])
        data["next_action_batch"] = f32(self.action[inc])
        data["next_point_state_batch"] = self.point_state[inc]
        data["next_return_batch"] = self.returns[inc]
        data["point_state_batch"] = self.point_state[batch_idx]
        data["time_batch"] = f32(self.timestep[self.episode_map[batch_idx]]) + 1 - data["time_batch"]   # remaining steps
        data["expert_flag_batch"] = f32(self.expert_flags[batch_idx])
        data["perturb_flag_batch"] = f32(self.perturb_flags[batch_idx])
        data["batch_idx"] = np.uint8(batch_idx)                                # sic: wraps modulo 256
        return data

    def sample(self, batch_size):                                              # :166-176
        return self.gather(self.draw_indices(batch_size))
