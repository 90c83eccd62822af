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


_CRIT = {
    'mse': (lambda d: d * d, lambda d: 2 * d),
    'sl1': (lambda d: np.where(np.abs(d) < 1, 0.5 * d * d, np.abs(d) - 0.5), lambda d: np.where(np.abs(d) < 1, d, np.sign(d))),
    'l1': (lambda d: np.abs(d), lambda d: np.sign(d)),
}


def cr_mask(coords, cr_indices, threshold):
    """``JointsCompositeLoss.get_cr_mask`` ``function.py:140-153``: a line of four points counts when its smallest
    non-zero pairwise distance exceeds the threshold (fore-shortened edges are dropped)."""
    coords = np.asarray(coords)
    mask = np.zeros((len(coords), len(cr_indices)))
    for b in range(len(coords)):
        for l, idx in enumerate(cr_indices):
            p = coords[b][idx].astype(np.float64)
            dm = np.sqrt(((p[:, None] - p[None]) ** 2).sum(-1))
            nz = dm[np.nonzero(dm)]
            mask[b, l] = 1.0 if len(nz) and nz.min() > threshold else 0.0
    return mask


def composite_coord_terms(coords_pred, joints_px, img_size, coor_kind='l1', coor_weight=0.1, cr_indices=None,
                          cr_kind='sl1', cr_weight=0.0, target_cr=4 / 3, cr_threshold=0.15):
    """Coordinate and cross-ratio terms of ``JointsCompositeLoss`` (``calc_coor_loss`` ``function.py:159-168``,
    ``calc_cross_ratio_loss`` ``:113-138`` with ``appro_cr`` ``img_proc.py:709-720``) and their analytic gradient
    w.r.t. the predicted coordinates.  Returns (total, coor, cr, grad [B,K,2]) in float64."""
    p = np.asarray(coords_pred, dtype=np.float64)
    gt = np.asarray(joints_px, dtype=np.float64)[:, :, :2].astype(np.float32).astype(np.float64)
    gt = gt / np.array([img_size[0], img_size[1]], dtype=np.float64)
    f, df = _CRIT[coor_kind]
    d = p - gt
    coor = f(d).mean()
    grad = coor_weight * df(d) / d.size
    cr = 0.0
    if cr_indices is not None and cr_weight:
        mask = cr_mask(np.asarray(coords_pred, dtype=np.float32), cr_indices, cr_threshold)
        total = mask.sum()
        fc, dfc = _CRIT[cr_kind]
        if total > 0:
            for b in range(len(p)):
                for l, idx in enumerate(cr_indices):
                    if not mask[b, l]:
                        continue
                    A, B_, C, D = p[b][idx]
                    AC, BD, BC, AD = C - A, D - B_, C - B_, D - A
                    a, bb, c, dd = AC @ AC, BD @ BD, BC @ BC, AD @ AD
                    v = a * bb / (c * dd) / target_cr ** 2
                    cr += float(fc(np.float64(v - 1))) / total
                    k = cr_weight * float(dfc(np.float64(v - 1))) * v / total
                    grad[b, idx[0]] += k * (-2 * AC / a + 2 * AD / dd)
                    grad[b, idx[1]] += k * (-2 * BD / bb + 2 * BC / c)
                    grad[b, idx[2]] += k * (2 * AC / a - 2 * BC / c)
                    grad[b, idx[3]] += k * (2 * BD / bb - 2 * AD / dd)
    return coor_weight * coor + cr_weight * cr, coor, cr, grad
