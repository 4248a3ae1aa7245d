"""CPU restatement of the two point-matching losses of the update path (oracle; test infrastructure).

Follows /root/reference/core/loss.py:17-31 and the geometry helpers it calls in
/root/reference/core/utils.py: get_control_point_tensor :814-831, transform_control_points :834-874,
tc_rotation_matrix :890-911, control_points_from_rot_and_trans :926-937, qrot :940-958, rotZ :624-634.
"""
import numpy as np
import torch

# six gripper control points in the hand frame (utils.py:819-824)
CONTROL_POINTS = np.array(
    [[0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [0.053, -0.0, 0.075], [-0.053, 0.0, 0.075], [0.053, -0.0, 0.105],
     [-0.053, 0.0, 0.105]], dtype=np.float32)


def control_points(rotz: bool) -> torch.Tensor:
    """(6,3) float32.  rotz=True right-multiplies by Rz(pi/2) in float64 numpy, then casts (utils.py:826-829)."""
    cp = CONTROL_POINTS
    if rotz:
        a = np.pi / 2
        rz = np.array([[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]])
        cp = np.matmul(cp, rz)
    return torch.tensor(cp).float()


def quat_rotate(q, v):
    """utils.py:940-958: v + 2*(w*(u x v) + u x (u x v)), q = (w, u)."""
    u = q[..., 1:]
    uv = torch.cross(u, v, dim=-1)
    uuv = torch.cross(u, uv, dim=-1)
    return v + 2 * (q[..., :1] * uv + uuv)


def grasp_control_points(qt):
    """(M,7) [quat wxyz, trans] -> (M,6,3) control points (transform_control_points, mode 'qt', rotz=True)."""
    cp = control_points(True).unsqueeze(0).expand(qt.shape[0], -1, -1)
    q = qt[:, None, :4].expand(-1, 6, -1)
    return quat_rotate(q, cp) + qt[:, None, 4:]


def goal_pred_loss(pred, goal):
    """loss.py:17-23.  Empty input -> mean over nothing = NaN, exactly like the reference."""
    return torch.mean(torch.abs(grasp_control_points(pred) - grasp_control_points(goal)).sum(-1))


def euler_rotation(az, el, th):
    """utils.py:890-911 (batched): Rz(th) @ Ry(el) @ Rx(az)."""
    cx, cy, cz = torch.cos(az), torch.cos(el), torch.cos(th)
    sx, sy, sz = torch.sin(az), torch.sin(el), torch.sin(th)
    one, zero = torch.ones_like(cx), torch.zeros_like(cx)
    rx = torch.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], -1).reshape(-1, 3, 3)
    ry = torch.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], -1).reshape(-1, 3, 3)
    rz = torch.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], -1).reshape(-1, 3, 3)
    return torch.matmul(rz, torch.matmul(ry, rx))


def action_control_points(a):
    """(M,6) [trans, euler] -> (M,6,3): cp @ R^T + t (control_points_from_rot_and_trans, no rotz)."""
    rot = euler_rotation(a[:, 3], a[:, 4], a[:, 5])
    cp = control_points(False).unsqueeze(0).expand(a.shape[0], -1, -1)
    return torch.matmul(cp, rot.permute(0, 2, 1)) + a[:, None, :3]


def pose_bc_loss(pi, expert):
    """loss.py:25-31."""
    return torch.mean(torch.abs(action_control_points(pi) - action_control_points(expert)).sum(-1))
