"""Candidate pose grid (reference vision_3d/obj_pose_opt.py:8-54).  pytorch3d is not a dependency
here: euler_angles_to_matrix(., 'XYZ') is spelled out (Rx @ Ry @ Rz)."""
import math

import torch

# per scene_type: x, y, z offsets from scene_centre and the Euler range (obj_pose_opt.py:16-36)
_BOUNDS = {
    0: ((-0.12, 0.04), (-0.10, 0.06), (0.00, 0.085), (0.0, 0.0)),            # pool table
    1: ((-0.15, 0.20), (0.40, 0.44), (0.04, 0.41), (-math.pi, math.pi / 2)),  # shelf (6-DoF)
    3: ((-0.19, 0.15), (-0.25, 0.10), (0.00, 0.14), (0.0, 0.0)),              # shopping
}


def euler_angles_to_matrix_xyz(eulers: torch.Tensor) -> torch.Tensor:
    c, s = torch.cos(eulers), torch.sin(eulers)
    one, zero = torch.ones_like(c[:, 0]), torch.zeros_like(c[:, 0])
    rx = torch.stack([one, zero, zero, zero, c[:, 0], -s[:, 0], zero, s[:, 0], c[:, 0]], -1).view(-1, 3, 3)
    ry = torch.stack([c[:, 1], zero, s[:, 1], zero, one, zero, -s[:, 1], zero, c[:, 1]], -1).view(-1, 3, 3)
    rz = torch.stack([c[:, 2], -s[:, 2], zero, s[:, 2], c[:, 2], zero, zero, zero, one], -1).view(-1, 3, 3)
    return torch.matmul(torch.matmul(rx, ry), rz)


def sample_poses_grid(task_model, sample_res=[40, 40, 1, 1, 1, 1], scene_type=0):
    """Absolute world-frame poses, [N,16] flattened homogeneous matrices, x slowest / z-rotation fastest."""
    scene_model = task_model.scene_model
    device = scene_model.device
    if scene_type not in _BOUNDS:
        raise NotImplementedError("scene_type %d not implemented" % scene_type)
    bx, by, bz, brot = _BOUNDS[scene_type]
    centre = scene_model.scene_centre
    axes = []
    for i, b in enumerate((bx, by, bz)):
        lo, hi = torch.tensor(b) + centre[i]
        axes.append(torch.linspace(lo, hi, sample_res[i]).to(device))
    for i in range(3):
        lo, hi = torch.tensor(brot)
        axes.append(torch.linspace(lo, hi, sample_res[3 + i]).to(device))
    combos = torch.cartesian_prod(*axes)
    pose_batch = torch.eye(4).repeat(combos.shape[0], 1, 1).to(device)
    pose_batch[:, :3, 3] = combos[:, :3]
    pose_batch[:, :3, :3] = euler_angles_to_matrix_xyz(combos[:, 3:])
    return pose_batch.reshape(-1, 16)
