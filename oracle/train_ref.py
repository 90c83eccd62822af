"""CPU restatement of one training step of the heat-map network (TEST INFRASTRUCTURE; SURVEY.md 8a row a12,
BASELINE configs[3]).  The CUDA train path is not built yet -- this is the pinned target it will be held to.

* ``generate_target``         -- ``libs/common/img_proc.py:347-409``: un-normalised Gaussian dots (sigma, 3-sigma
  support, centre = int(joint / stride + 0.5)) and the visibility weights.
* ``train_forward_backward``  -- the loop body of ``libs/trainer/trainer.py:183-198`` without the optimiser:
  ``model.train()`` forward (``hrnet.py:563-614`` with BatchNorm2d in training mode: batch statistics, running
  statistics updated with momentum 0.1 and the unbiased variance), the heat-map loss
  (``JointsMSELoss.forward`` ``libs/loss/function.py:28-46`` / ``calc_hm_loss`` ``:95-111``) and
  ``loss.backward()``.  The forward is ``oracle.hrnet_ref.hrnet_forward`` with ``ctx=Train`` (same torch-CPU ops
  the reference module executes), the backward is autograd on that graph.

Pinned by ``tests/golden/make_golden.py::golden_train`` against the reference module itself in train mode.
"""
from collections import OrderedDict

import numpy as np
import torch

from . import hrnet_ref


class Train(hrnet_ref.Exact):
    """Reference arithmetic in training mode (see hrnet_ref._cb)."""
    training = True


def generate_target(joints, joints_vis, num_joints, input_size, heatmap_size, sigma):
    """joints [K,3] (crop pixels), joints_vis [K]; input_size / heatmap_size as numpy arrays [w, h] -- upstream
    indexes ``heatmap_size[0]`` as the number of ROWS when allocating (square maps in every shipped config)."""
    input_size, heatmap_size = np.asarray(input_size), np.asarray(heatmap_size)
    target_weight = np.ones((num_joints, 1), dtype=np.float32)
    target_weight[:, 0] = joints_vis
    target = np.zeros((num_joints, heatmap_size[0], heatmap_size[1]), dtype=np.float32)
    tmp_size = sigma * 3
    for k in range(num_joints):
        if target_weight[k] <= 0.5:
            continue
        feat_stride = input_size / heatmap_size
        mu_x = int(joints[k][0] / feat_stride[0] + 0.5)
        mu_y = int(joints[k][1] / feat_stride[1] + 0.5)
        ul = [int(mu_x - tmp_size), int(mu_y - tmp_size)]
        br = [int(mu_x + tmp_size + 1), int(mu_y + tmp_size + 1)]
        if ul[0] >= heatmap_size[1] or ul[1] >= heatmap_size[0] or br[0] < 0 or br[1] < 0:
            target_weight[k] = 0
            continue
        size = 2 * tmp_size + 1
        x = np.arange(0, size, 1, np.float32)
        y = x[:, np.newaxis]
        x0 = y0 = size // 2
        g = np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * sigma ** 2))
        g_x = max(0, -ul[0]), min(br[0], heatmap_size[1]) - ul[0]
        g_y = max(0, -ul[1]), min(br[1], heatmap_size[0]) - ul[1]
        img_x = max(0, ul[0]), min(br[0], heatmap_size[1])
        img_y = max(0, ul[1]), min(br[1], heatmap_size[0])
        target[k][img_y[0]:img_y[1], img_x[0]:img_x[1]] = g[g_y[0]:g_y[1], g_x[0]:g_x[1]]
    return target, target_weight


def joints_mse_loss(pred, target, target_weight=None):
    """torch form of JointsMSELoss.forward (function.py:28-46): mean over joints of 0.5 * MSE(w*pred, w*gt)."""
    B, K = pred.shape[:2]
    p = pred.reshape(B, K, -1)
    t = target.reshape(B, K, -1)
    loss = 0
    for k in range(K):
        pk, tk = p[:, k], t[:, k]
        if target_weight is not None:
            pk, tk = pk * target_weight[:, k], tk * target_weight[:, k]
        loss = loss + 0.5 * torch.mean((pk - tk) ** 2)
    return loss / K


def train_forward_backward(sd, cfgs, x, target, target_weight=None, coords_gt=None, coor_weight=0.1):
    """One forward + backward in training mode.  sd: state dict (not modified).
    ``coords_gt`` ([B,K,2] in (0,1), coordinate head only) adds the shipped composite loss' second term,
    ``coor_weight * mean |coords - coords_gt|`` (``JointsCompositeLoss`` ``function.py:170-202`` with
    ``loss_spec_list: ['mse', 'l1', ...]``, ``loss_weight_list: [1.0, 0.1, ...]``, KITTI_train_IGRs.yml:88-89).
    Returns (loss float, grads {name: tensor}, new_sd with the updated BN running statistics)."""
    work = OrderedDict()
    for k, v in sd.items():
        v = v.detach().clone()
        if v.is_floating_point() and 'running_' not in k:
            v.requires_grad_(True)
        work[k] = v
    with torch.enable_grad():
        out = hrnet_ref.hrnet_forward.__wrapped__(work, cfgs, x, ctx=Train)
        maps = out[0] if isinstance(out, tuple) else out
        loss = joints_mse_loss(maps, target, target_weight)
        if coords_gt is not None:
            loss = loss + coor_weight * torch.mean(torch.abs(out[1] - coords_gt))
        loss.backward()
    grads = OrderedDict((k, v.grad) for k, v in work.items() if v.requires_grad and v.grad is not None)
    new_sd = OrderedDict()
    for k, v in work.items():
        new_sd[k] = v.detach() + 1 if k.endswith('num_batches_tracked') else v.detach()
    return float(loss.detach()), grads, new_sd
