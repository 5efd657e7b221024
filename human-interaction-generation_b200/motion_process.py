"""Sampled motion -> 3-D joints on the GPU (mirror of the reference's codes/utils/motion_process.py, two-person part).

  recover_from_ric2(data1, data2, joints_num)   same name / argument meaning as utils/motion_process.py:418-456:
                                                data* [B, T, 263] with the init-state row LAST (what plot_t2m2 passes,
                                                tools/visualization.py:55), returns two [B, T-1, joints_num, 3] tensors
  joints_from_samples(x, mean, std, ...)        the whole tail of tools/visualization.py:149-155 + recover_from_ric2 on
                                                the sampler's own layout (x [2B, T, 263], init state in row 0), one kernel
  mpjpe(a, b)                                   mean per-joint position error between two joint tensors

Everything runs in ONE launch of hig_recover_joints (csrc/joints.cu) on the tensors' device; there is no CPU path.
"""
import torch

from . import ops


def recover_from_ric2(data1, data2, joints_num=22):
    if data1.shape != data2.shape or data1.dim() != 3:
        raise ValueError("recover_from_ric2: data1 / data2 must both be [B, T, C]")
    B = data1.shape[0]
    x = torch.cat([data1, data2], dim=0).to(torch.float32).contiguous()
    j = ops.recover_joints(x, joints_num=joints_num, init_row=-1)
    return j[:B], j[B:]


def joints_from_samples(x, mean=None, std=None, init_mean=None, init_std=None, length=None, joints_num=22):
    """x [S, T, C] fp32 CUDA as p_sample_loop returns it -> [S, T-1, joints_num, 3]; rows >= length[s] give zeros."""
    return ops.recover_joints(x.to(torch.float32).contiguous(), mean, std, init_mean, init_std, length=length,
                              joints_num=joints_num, init_row=0)


def mpjpe(a, b, length=None):
    """Mean Euclidean joint distance over (sequence, frame, joint); length [S] (rows incl. the init row) restricts the
    mean to each sequence's valid frames."""
    d = (a.double() - b.double()).pow(2).sum(-1).sqrt()          # [S, F, J]
    if length is None:
        return d.mean()
    F = d.shape[1]
    valid = (torch.arange(F, device=d.device)[None, :] < (torch.as_tensor(length, device=d.device).reshape(-1, 1) - 1))
    w = valid[..., None].expand_as(d).double()
    return (d * w).sum() / w.sum().clamp(min=1)
