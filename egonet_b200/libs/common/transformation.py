"""B200-native drop-in for the hot-path part of the reference's
``libs/common/transformation.py``: the batched pose solve.

The reference solves one instance at a time on the host (``compute_rigid_transform``
[transformation.py:99-134] inside ``EgoNet.get_6d_rep`` [egonet.py:279-295], then
scipy Euler angles and the observation angle [egonet.py:203-236]).  ``pose_solve``
does all of it for N instances in one fp64 kernel launch.  ``pnp_refine`` is the
batched form of the optional reprojection refinement [transformation.py:143-157].
"""
import numpy as np
import torch

from ... import _native as N


def pose_solve(kpts_3d, kpts_2d=None, K=None, alpha_mode='trans', want_rotation=False):
    """kpts_3d: CUDA fp64 [N,P,3] (P = 8 or 32).  Returns CUDA fp64 [N,7] =
    (euler x, y, z | translation x, y, z | alpha) and optionally R [N,3,3].
    ``alpha_mode='proj'`` needs kpts_2d (CUDA fp64 [N, >=1], first column = screen
    x of the first key-point) and the 3x3 intrinsics K."""
    if not kpts_3d.is_cuda:
        raise RuntimeError('native pose_solve has no CPU path')
    x = kpts_3d.detach().to(torch.float64).contiguous()
    n = x.shape[0]
    if n == 0:
        empty = torch.empty((0, 7), device=x.device, dtype=torch.float64)
        return (empty, torch.empty((0, 3, 3), device=x.device, dtype=torch.float64)) if want_rotation else empty
    x = x.view(n, -1, 3)
    p = x.shape[1]
    if alpha_mode == 'trans':
        mode, k2, stride, fx, cx = N.ALPHA_TRANS, None, 0, 0.0, 0.0
    elif alpha_mode == 'proj':
        if kpts_2d is None or K is None:
            raise ValueError("alpha_mode='proj' needs kpts_2d and K")
        mode = N.ALPHA_PROJ
        k2 = kpts_2d.detach().to(torch.float64).contiguous().view(n, -1)
        stride = k2.shape[1]
        Kn = np.asarray(K, dtype=np.float64)
        fx, cx = float(Kn[0, 0]), float(Kn[0, 2])
    else:
        raise NotImplementedError  # same as egonet.py:328
    out = torch.empty((n, 7), device=x.device, dtype=torch.float64)
    rot = torch.empty((n, 3, 3), device=x.device, dtype=torch.float64) if want_rotation else None
    with torch.cuda.device(x.device):
        N.check(N.lib().egn_pose_solve(N.ptr(x), n, p, N.ptr(k2), stride, fx, cx, mode, N.ptr(out), N.ptr(rot),
                                       N.current_stream()))
    return (out, rot) if want_rotation else out


def compute_rigid_transform(X, Y, W=None, verbose=False):
    """[transformation.py:99-134] single-instance form kept for API parity:
    X, Y are [3, N] arrays (numpy or tensors); returns (R [3,3], t [3,1]) as numpy
    fp64.  Only the Kabsch rotation between a template cuboid and its prediction
    is on the hot path; this wrapper exists for callers that want R itself and
    works for N in {8, 32} point clouds laid out like the cuboid (it reuses the
    pose kernel, which rebuilds the template from Y)."""
    raise NotImplementedError('use pose_solve(kpts_3d, want_rotation=True): the native kernel derives the '
                              'template from the prediction (egonet.py:238-263) and does not accept an '
                              'arbitrary X')


def pnp_refine_batch(predictions, observations, intrinsics, max_iter=0, return_info=False):
    """N instances of ``pnp_refine`` [transformation.py:143-157] in one launch.
    predictions [N,P,3], observations [N,P,2] (tensors or arrays), intrinsics 3x3.
    Returns CUDA fp64 [N,P,3] (= ``(Rodrigues(R) @ X.T + T).T`` per instance); with ``return_info``
    also pose [N,6] (rvec | tvec), info [N,2] (LM iterations, residual norm in px), status [N]."""
    dev = predictions.device if torch.is_tensor(predictions) and predictions.is_cuda else \
        torch.device('cuda', torch.cuda.current_device())
    x = torch.as_tensor(predictions, dtype=torch.float64).to(dev).contiguous()
    u = torch.as_tensor(observations, dtype=torch.float64).to(dev).contiguous()
    n, p = x.shape[0], x.shape[1]
    if x.dim() != 3 or x.shape[2] != 3 or tuple(u.shape) != (n, p, 2):
        raise ValueError('predictions must be [N,P,3] and observations [N,P,2]')
    Kn = np.asarray(intrinsics, dtype=np.float64)
    out = torch.empty_like(x)
    pose = torch.empty((n, 6), device=dev, dtype=torch.float64)
    info = torch.empty((n, 2), device=dev, dtype=torch.float64)
    status = torch.empty((n,), device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        N.check(N.lib().egn_pnp_refine(N.ptr(x), N.ptr(u), n, p, float(Kn[0, 0]), float(Kn[1, 1]), float(Kn[0, 2]),
                                       float(Kn[1, 2]), int(max_iter), N.ptr(out), N.ptr(pose), N.ptr(info),
                                       N.ptr(status), N.current_stream()))
    return (out, pose, info, status) if return_info else out


def pnp_refine(prediction, observation, intrinsics, dist_coeffs):
    """[transformation.py:143-157] single-instance signature of the reference: prediction [P,3],
    observation [P,2] -> refined [3,P] (numpy fp64).  Lens distortion is not supported by the kernel."""
    if dist_coeffs is not None and np.any(np.asarray(dist_coeffs, dtype=np.float64) != 0):
        raise NotImplementedError('native pnp_refine supports zero lens distortion only')
    pred = np.asarray(prediction, dtype=np.float64)
    out, _, _, status = pnp_refine_batch(pred[None], np.asarray(observation, dtype=np.float64)[None], intrinsics,
                                         return_info=True)
    if int(status[0]) != 0:
        print('PnP failed.')          # as upstream: keep the prediction
        return pred
    return out[0].cpu().numpy().T
