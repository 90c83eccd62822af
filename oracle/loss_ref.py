"""CPU restatement of the reference's heat-map MSE loss (TEST INFRASTRUCTURE).

``joints_mse_loss`` -- ``JointsMSELoss.forward`` ``libs/loss/function.py:28-46`` (and, with
``target_weight=None``, ``JointsCompositeLoss.calc_hm_loss`` ``:95-111``): per joint
``0.5 * mean((w*pred - w*gt)^2)`` over batch and pixels, averaged over joints.  The gradient is the
analytic derivative (what autograd gives upstream).
"""
import numpy as np


def joints_mse_loss(pred, target, target_weight=None):
    pred = np.asarray(pred, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    B, K = pred.shape[:2]
    w = np.ones((B, K, 1, 1)) if target_weight is None else np.asarray(target_weight, dtype=np.float64).reshape(B, K, 1, 1)
    d = w * pred - w * target
    n = pred.size
    loss = 0.0
    for k in range(K):
        loss += 0.5 * np.mean(d[:, k] ** 2)
    loss /= K
    grad = w * d / n
    return loss, grad
