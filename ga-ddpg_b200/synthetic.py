"""Seeded synthetic replay minibatches with the layout ``BaseMemory.sample`` produces
(/root/reference/core/replay_memory.py:166-176,251-272) — SURVEY.md §8(d) "Synthetic distributions".

Clouds are (B, C, N+6): six leading hand-finger columns (/root/reference/core/utils.py:38-40, mask
channel = 1) followed by N object points on the surface of a random box (mask channel = 0); the
encoder drops the 6 hand columns whenever the last dim is not 1024 (networks.py:234-235).
"""
import numpy as np

HAND_FINGER_POINT = np.array(
    [[0.0, 0.0, 0.0, -0.0, 0.0, -0.0],
     [0.0, 0.0, 0.053, -0.053, 0.053, -0.053],
     [0.0, 0.0, 0.075, 0.075, 0.105, 0.105]], dtype=np.float64)

ACTION_HIGH = np.array([0.06, 0.06, 0.06, np.pi / 6, np.pi / 6, np.pi / 6])
ACTION_LOW = -ACTION_HIGH


def _box_surface(rs, n, centre, half):
    """n points uniform on the surface of an axis-aligned box."""
    area = np.array([half[1] * half[2], half[0] * half[2], half[0] * half[1]])
    axis = rs.choice(3, size=n, p=area / area.sum())
    pts = rs.uniform(-1.0, 1.0, size=(n, 3)) * half
    sign = rs.choice([-1.0, 1.0], size=n)
    pts[np.arange(n), axis] = sign * half[axis]
    return pts + centre


def make_cloud(rs, B, N, channels=4, centre=None, half=None, zero_frac=0.005, dtype=np.float32):
    cloud = np.zeros((B, channels, N + 6), dtype=np.float64)
    if centre is None:
        centre = np.stack([rs.uniform(-0.05, 0.05, B), rs.uniform(-0.05, 0.05, B), rs.uniform(0.15, 0.45, B)], 1)
    if half is None:
        half = rs.uniform(0.03, 0.10, size=(B, 3))
    for b in range(B):
        pts = _box_surface(rs, N, centre[b], half[b]) + rs.normal(0.0, 1e-3, size=(N, 3))
        cloud[b, :3, :6] = HAND_FINGER_POINT
        cloud[b, 3, :6] = 1.0
        cloud[b, :3, 6:] = pts.T
        if channels > 4:
            cloud[b, 4:, :] = rs.normal(0.0, 1.0, size=(channels - 4, N + 6))
        nz = int(round(zero_frac * N))
        if nz > 0:
            cols = rs.choice(N, size=nz, replace=False) + 6
            cloud[b, :, cols] = 0.0
    return cloud.astype(dtype), centre, half


def make_batch(B, N, step=0, channels=4, dtype=np.float32, seed=1234, compact=False):
    """One replay minibatch dict (keys = what Agent.prepare_data consumes, agent.py:211-240).  ``compact``: 2 cm-cube
    objects — every SA1 ball (r = 0.02) then holds >= 64 distinct points and all 32 SA2 centroids lie within one SA2
    radius (0.04) of each other, i.e. the ball-query groups have no duplicate slots to fold (the dense worst case)."""
    rs = np.random.RandomState(seed + step)
    action = rs.uniform(ACTION_LOW, ACTION_HIGH, size=(B, 6)).astype(np.float32)
    cloud, centre, half = make_cloud(rs, B, N, channels, dtype=dtype, half=np.full((B, 3), 0.01) if compact else None,
                                     zero_frac=0.0 if compact else 0.005)
    next_cloud, _, _ = make_cloud(rs, B, N, channels, centre=centre - action[:, :3].astype(np.float64), half=half, dtype=dtype)
    ret = (rs.rand(B) < 0.6).astype(np.float32) * rs.uniform(0.2, 1.0, B).astype(np.float32)
    if not (ret > 0).any():
        ret[rs.randint(B)] = 1.0  # an empty goal mask gives NaN losses in the reference (loss.py:17-23)
    expert = (rs.rand(B) < 0.7).astype(np.float32)
    if not (expert >= 1).any():
        expert[rs.randint(B)] = 1.0

    def goal():
        q = rs.normal(size=(B, 4))
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        return np.concatenate([q, rs.uniform(-0.1, 0.3, size=(B, 3))], 1).astype(np.float32)

    batch = {
        "point_state_batch": cloud,
        "next_point_state_batch": next_cloud,
        "image_state_batch": np.zeros((B, 1), np.float32),
        "next_image_state_batch": np.zeros((B, 1), np.float32),
        "action_batch": action,
        "next_action_batch": rs.uniform(ACTION_LOW, ACTION_HIGH, size=(B, 6)).astype(np.float32),
        "expert_action_batch": rs.uniform(ACTION_LOW, ACTION_HIGH, size=(B, 6)).astype(np.float32),
        "next_expert_action_batch": rs.uniform(ACTION_LOW, ACTION_HIGH, size=(B, 6)).astype(np.float32),
        "reward_batch": (rs.rand(B) < 0.1).astype(np.float32),
        "return_batch": ret,
        "next_return_batch": ret.copy(),
        "mask_batch": (rs.rand(B) < 0.05).astype(np.float32),
        "time_batch": rs.randint(1, 21, size=B).astype(np.float32),
        "expert_flag_batch": expert,
        "perturb_flag_batch": (rs.rand(B) < 0.1).astype(np.float32),
        "goal_batch": goal(),
        "next_goal_batch": goal(),
        "grasp_sample_batch": np.zeros([0, 4, 4]),
        "batch_idx": np.arange(B).astype(np.uint8),
    }
    if not (batch["perturb_flag_batch"] < 1).any():
        batch["perturb_flag_batch"][0] = 0.0
    return batch


def make_episode(length, N, seed=0, success=True, expert=True):
    """One rollout as the list of per-step dicts ``BaseMemory.add_episode`` consumes (replay_memory.py:209-232; keys =
    ``attr_names`` :33-50 as the environment loop fills them): float64 ``point_state`` (4, N+6), 6-D actions, the
    sparse terminal reward, 1-based ``timestep``, unit-quaternion + translation ``goal``."""
    rs = np.random.RandomState(10_000 + seed)
    centre = np.array([[rs.uniform(-0.05, 0.05), rs.uniform(-0.05, 0.05), rs.uniform(0.15, 0.45)]])
    half = rs.uniform(0.03, 0.10, size=(1, 3))
    episode = []
    for t in range(length):
        action = rs.uniform(ACTION_LOW, ACTION_HIGH).astype(np.float32)
        cloud, _, _ = make_cloud(rs, 1, N, 4, centre=centre, half=half, dtype=np.float64)
        centre = centre - action[None, :3].astype(np.float64)
        q = rs.normal(size=4)
        q /= np.linalg.norm(q)
        pose = np.eye(4, dtype=np.float32)
        pose[:3, 3] = rs.uniform(-0.3, 0.3, 3)
        last = t == length - 1
        episode.append({
            "action": action, "expert_action": rs.uniform(ACTION_LOW, ACTION_HIGH).astype(np.float32),
            "pose": rs.normal(size=64).astype(np.float32), "point_state": cloud[0], "target_idx": float(seed % 7),
            "reward": float(success and last), "terminal": float(last), "timestep": float(t + 1), "state_pose": pose,
            "collide": 0.0, "grasp": float(last), "perturb_flags": float(rs.rand() < 0.1), "expert_flags": float(expert),
            "goal": np.concatenate([q, rs.uniform(-0.1, 0.3, 3)]).astype(np.float32), "target_name": "box%d" % (seed % 7),
        })
    return episode
