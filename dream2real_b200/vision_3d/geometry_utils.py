"""Score heat-map smoothing (reference vision_3d/geometry_utils.py:252-269), torch only."""
import math

import torch
import torch.nn.functional as F


def _gaussian_kernel_3x3(sigma: float, dtype, device):
    # torchvision _get_gaussian_kernel1d(kernel_size=3): x = linspace(-1, 1, 3); pdf = exp(-0.5 (x/sigma)^2); normalise
    x = torch.linspace(-1.0, 1.0, 3, dtype=dtype, device=device)
    k = torch.exp(-0.5 * (x / sigma) ** 2)
    k = k / k.sum()
    return torch.outer(k, k)


def spatially_smooth_heatmap(pose_scores, sample_res, sigma=0.7):
    """zeros -> min non-zero, view as [rest, 1, X, Y], pad 1 with the minimum, 3x3 Gaussian, zeros restored.
    (The reference pads with the constant and then lets torchvision reflect-pad the padded image; the
    reflect ring is cropped away again, so a constant-padded 3x3 convolution is identical.)"""
    s = pose_scores.clone()
    min_nonzero = torch.min(s[s != 0]).item()
    zero_idxs = torch.nonzero(s == 0, as_tuple=True)
    s[zero_idxs] = min_nonzero
    rest = sample_res[2] * sample_res[3] * sample_res[4] * sample_res[5]
    img = s.view(sample_res[0] * sample_res[1], rest).swapaxes(0, 1).reshape(rest, 1, sample_res[0], sample_res[1])
    img = F.pad(img, (1, 1, 1, 1), mode="constant", value=min_nonzero)
    k = _gaussian_kernel_3x3(float(sigma), img.dtype, img.device)
    out = F.conv2d(img, k[None, None])
    out = out.reshape(rest, sample_res[0] * sample_res[1]).swapaxes(0, 1).reshape(-1)
    out[zero_idxs] = 0
    return out.contiguous()
