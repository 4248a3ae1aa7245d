"""CPU restatement of ``regularize_pc_point_count`` (oracle; TEST INFRASTRUCTURE ONLY).

Follows /root/reference/core/utils.py:784-812 with the FPS / gather of oracle/pointnet2_ops_cpu (the restated
``pointnet2_ops``) in place of the CUDA extension.  Pinned in tests/test_regularize_cpu.py: when the reference tree is
present the UNMODIFIED reference function (running on the same CPU ops through oracle/refstack.py) must return
identical arrays for all three branches."""
import numpy as np
import torch

from .pointnet2_ops_cpu import pointnet2_utils as U


def regularize_pc_point_count(pc, npoints, use_farthest_point=False):
    if pc.shape[0] > npoints:
        if use_farthest_point:
            t = torch.from_numpy(pc)[None].float()
            idx = U.fps_raw(t[..., :3].contiguous(), npoints)
            new = torch.gather(t.transpose(1, 2), 2, idx.long().unsqueeze(1).expand(-1, t.shape[2], -1))
            pc = new[0].T.contiguous().numpy()
        else:
            pc = pc[np.random.choice(range(pc.shape[0]), size=npoints, replace=False), :]
    else:
        required = npoints - pc.shape[0]
        if required > 0:
            pc = np.concatenate((pc, pc[np.random.choice(range(pc.shape[0]), size=required), :]), axis=0)
    return pc
