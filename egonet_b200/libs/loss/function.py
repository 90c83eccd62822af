"""B200-native drop-in for the heat-map loss of the reference's ``libs/loss/function.py``.

``JointsMSELoss`` keeps the upstream constructor and ``forward(output, target, target_weight, meta)``
signature [function.py:22-46]; ``calc_hm_loss`` is the same quantity as
``JointsCompositeLoss.calc_hm_loss`` [function.py:95-111].  Forward and the gradient w.r.t. the
predicted heat-maps come out of one fused kernel (``egn_mse_hm_fwd_bwd``); the losses are autograd
nodes, so ``loss.backward()`` continues into the native training engine of HC (trainer.py:183-198).
"""
import torch
import torch.nn as nn

from ... import _native as N


def mse_hm_fwd_bwd(output, target, target_weight=None, want_grad=True):
    """(loss scalar tensor, grad [B,K,H,W] or None) on the device of ``output``."""
    if not output.is_cuda:
        raise RuntimeError('native heat-map loss has no CPU path')
    out = output.detach().float().contiguous()
    tgt = target.detach().float().contiguous().to(out.device)
    B, K, H, W = out.shape
    w = None
    if target_weight is not None:
        w = target_weight.detach().float().reshape(B, K).contiguous().to(out.device)
    loss = torch.empty((), device=out.device, dtype=torch.float32)
    grad = torch.empty_like(out) if want_grad else None
    ws = torch.empty(1, device=out.device, dtype=torch.float64)
    with torch.cuda.device(out.device):
        N.check(N.lib().egn_mse_hm_fwd_bwd(N.ptr(out), N.ptr(tgt), N.ptr(w), B, K, H, W, N.ptr(loss), N.ptr(grad),
                                           N.ptr(ws), N.current_stream()))
    return loss, grad


class _HeatmapMSE(torch.autograd.Function):
    """loss = 0.5 * mean over joints of MSE(w * pred, w * gt); backward hands out the gradient the same kernel
    launch already produced, scaled by the incoming gradient."""

    @staticmethod
    def forward(ctx, output, target, target_weight):
        need = output.requires_grad
        loss, grad = mse_hm_fwd_bwd(output, target, target_weight, want_grad=need)
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        return (ctx.grad * grad_loss if ctx.grad is not None else None), None, None


class JointsMSELoss(nn.Module):
    def __init__(self, use_target_weight):
        super().__init__()
        self.use_target_weight = use_target_weight

    def forward(self, output, target, target_weight, meta=None):
        return _HeatmapMSE.apply(output, target, target_weight if self.use_target_weight else None)


def calc_hm_loss(output, target):
    return _HeatmapMSE.apply(output, target, None)
