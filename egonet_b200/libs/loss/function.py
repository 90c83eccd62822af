"""B200-native drop-in for the training losses of the reference's ``libs/loss/function.py``.

``JointsCompositeLoss`` [function.py:61-202] keeps the upstream constructor, attributes (``cr_indices``,
``target_cr``, ``apply_cr_loss``) and ``forward(output, target, target_weight, meta)``: heat-map term through
``egn_mse_hm_fwd_bwd``, coordinate (L1 / smooth-L1 / MSE) and cross-ratio terms through ``egn_coord_loss_fwd_bwd``.

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


_KINDS = {'mse': 0, 'sl1': 1, 'l1': 2}


def coord_loss_fwd_bwd(coords_pred, coords_gt_px, img_size, coor_kind, coor_weight, cr_indices=None, cr_kind='sl1',
                       cr_weight=0.0, target_cr=1.0, cr_threshold=0.15, want_grad=True):
    """(losses [3] = total | coor | cr, grad [B,K,2] or None) on the device of ``coords_pred``."""
    import numpy as np
    if not coords_pred.is_cuda:
        raise RuntimeError('native coordinate loss has no CPU path')
    p = coords_pred.detach().float().contiguous()
    g = torch.as_tensor(coords_gt_px, dtype=torch.float32).to(p.device).contiguous()
    B, K = p.shape[0], p.shape[1]
    if tuple(g.shape) != (B, K, 2) or p.shape[2] != 2:
        raise ValueError('coords_pred and coords_gt must both be [B,K,2]')
    idx, L = None, 0
    if cr_indices is not None and cr_weight != 0.0:
        idx = torch.as_tensor(np.asarray(cr_indices, dtype=np.int32)).to(p.device).contiguous()
        L = idx.shape[0]
        if idx.dim() != 2 or idx.shape[1] != 4 or int(idx.min()) < 0 or int(idx.max()) >= K:
            raise ValueError('cr_indices must be [L,4] point indices below K')
    losses = torch.empty(3, device=p.device, dtype=torch.float32)
    grad = torch.empty_like(p) if want_grad else None
    with torch.cuda.device(p.device):
        N.check(N.lib().egn_coord_loss_fwd_bwd(N.ptr(p), N.ptr(g), B, K, float(img_size[0]), float(img_size[1]),
                                               _KINDS[coor_kind], float(coor_weight), N.ptr(idx), L, _KINDS[cr_kind],
                                               float(cr_weight), float(target_cr), float(cr_threshold), N.ptr(losses),
                                               N.ptr(grad), N.current_stream()))
    return losses, grad


class _CoordLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coords_pred, coords_gt_px, cfg):
        losses, grad = coord_loss_fwd_bwd(coords_pred, coords_gt_px, want_grad=coords_pred.requires_grad, **cfg)
        ctx.grad = grad
        return losses[0]

    @staticmethod
    def backward(ctx, grad_loss):
        return (ctx.grad * grad_loss if ctx.grad is not None else None), None, None


class JointsCompositeLoss(nn.Module):
    """[function.py:61-202] L = w_hm * L_hm + w_coor * L_2d (+ w_cr * L_cr once ``apply_cr_loss`` is set)."""

    def __init__(self, spec_list, img_size, hm_size, loss_weights=[1, 1, 1], target_cr=None, cr_loss_thres=0.15,
                 use_target_weight=False):
        super().__init__()
        self.comp_dict = {}
        for name, spec, w in zip(('hm', 'coor', 'cr'), spec_list, loss_weights):
            if spec != 'None':
                if spec not in _KINDS:
                    raise KeyError(spec)
                self.comp_dict[name] = (spec, w)
        if 'hm' in self.comp_dict and self.comp_dict['hm'][0] != 'mse':
            raise NotImplementedError('the native heat-map term is the MSE of every shipped config')
        self.img_size = img_size
        self.hm_size = hm_size
        self.target_cr = target_cr
        self.use_target_weight = use_target_weight
        self.apply_cr_loss = False
        self.cr_loss_thres = cr_loss_thres

    def calc_hm_loss(self, output, target):
        return calc_hm_loss(output, target)

    def forward(self, output, target, target_weight=None, meta=None):
        import numpy as np
        heatmaps_pred, coordinates_pred = output if type(output) is tuple else (output, None)
        total = 0
        if 'hm' in self.comp_dict:
            if len(heatmaps_pred) != len(target):            # heat-maps of unlabeled data at the tail of the batch
                heatmaps_pred = heatmaps_pred[:len(target)]
            total = total + self.calc_hm_loss(heatmaps_pred, target) * self.comp_dict['hm'][1]
        use_cr = 'cr' in self.comp_dict and self.comp_dict['cr'][1] != 'None' and self.apply_cr_loss
        if 'coor' in self.comp_dict or use_cr:
            if coordinates_pred is None:
                raise NotImplementedError('coordinate / cross-ratio terms need the coordinate head '
                                          "(head_type='coordinates'); soft-argmax back-propagation is not built")
            gt = np.asarray(meta['transformed_joints'])[:, :, :2].astype(np.float32)
            coor_w = self.comp_dict['coor'][1] if 'coor' in self.comp_dict else 0.0
            n = len(gt)
            pred_fs = coordinates_pred[:n] if len(coordinates_pred) != n else coordinates_pred
            if use_cr and len(coordinates_pred) != n:
                raise NotImplementedError('cross-ratio term over unlabeled extra samples is not built')
            cfg = dict(img_size=self.img_size, coor_kind=self.comp_dict['coor'][0] if 'coor' in self.comp_dict else 'l1',
                       coor_weight=coor_w, cr_indices=getattr(self, 'cr_indices', None) if use_cr else None,
                       cr_kind=self.comp_dict['cr'][0] if use_cr else 'sl1',
                       cr_weight=self.comp_dict['cr'][1] if use_cr else 0.0,
                       target_cr=self.target_cr if use_cr else 1.0, cr_threshold=self.cr_loss_thres)
            total = total + _CoordLoss.apply(pred_fs, gt, cfg)
        return total
