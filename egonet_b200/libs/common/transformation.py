"""B200-native drop-in for the hot-path part of the reference's
``libs/common/transformation.py``: the batched pose solve and the point-set alignment helpers
(``compute_rigid_transform``, ``procrustes_transform``, ``compute_similarity_transform``), plus the
reprojection refinement and its caller ``refine_with_predicted_bbox`` (tools/inference_legacy.py:518-547).

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


def _points(a, dev):
    """[3,P] / [N,3,P] array-like (the reference's layout) -> CUDA fp64 [N,P,3]."""
    t = torch.as_tensor(np.asarray(a.detach().cpu() if torch.is_tensor(a) else a), dtype=torch.float64)
    if t.dim() == 2:
        t = t.unsqueeze(0)
    if t.dim() != 3 or t.shape[1] != 3:
        raise ValueError('point sets must be [3, P] (or batched [N, 3, P]); got %s' % (tuple(t.shape),))
    return t.transpose(1, 2).contiguous().to(dev)


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError('the native geometry kernels have no CPU path')
    return torch.device('cuda', torch.cuda.current_device())


def rigid_transform_batch(X, Y, W=None, want_aligned=False):
    """N instances of ``compute_rigid_transform`` [transformation.py:99-134] in one launch.
    X, Y: CUDA fp64 [N,P,3]; W: None, [N,P] (diagonal weights) or [N,P,P].
    Returns R [N,3,3], t [N,3] (and ``R X + t`` [N,P,3] with ``want_aligned``), all CUDA fp64."""
    n, p = X.shape[0], X.shape[1]
    if tuple(Y.shape) != (n, p, 3) or X.shape[2] != 3:
        raise ValueError('X and Y must both be [N,P,3]')
    mode = 0
    if W is not None:
        W = W.to(torch.float64).contiguous()
        mode = {2: 1, 3: 2}.get(W.dim())
        if mode is None or W.shape[0] != n or any(d != p for d in W.shape[1:]):
            raise ValueError('W must be [N,P] or [N,P,P]')
    R = torch.empty((n, 3, 3), device=X.device, dtype=torch.float64)
    t = torch.empty((n, 3), device=X.device, dtype=torch.float64)
    al = torch.empty((n, p, 3), device=X.device, dtype=torch.float64) if want_aligned else None
    with torch.cuda.device(X.device):
        N.check(N.lib().egn_rigid_transform(N.ptr(X.contiguous()), N.ptr(Y.contiguous()), N.ptr(W), mode, n, p,
                                            N.ptr(R), N.ptr(t), N.ptr(al), N.current_stream()))
    return (R, t, al) if want_aligned else (R, t)


def compute_rigid_transform(X, Y, W=None, verbose=False):
    """[transformation.py:99-134] least-squares rigid transform by SVD, reference signature:
    X, Y [d=3, N] (numpy or tensors), W optional [N] or [N,N]; returns (R [3,3], t [3,1]) numpy fp64."""
    assert len(X) == len(Y)
    dev = _device()
    Xd, Yd = _points(X, dev), _points(Y, dev)
    Wd = None
    if W is not None:
        Wn = np.asarray(W, dtype=np.float64)
        assert len(Wn.shape) in [1, 2]
        Wd = torch.as_tensor(Wn).unsqueeze(0).to(dev)
    R, t = rigid_transform_batch(Xd, Yd, Wd)
    return R[0].cpu().numpy(), t[0].cpu().numpy().reshape(3, 1)


def procrustes_transform(X, Y):
    """[transformation.py:136-141] rigid transform from X to Y applied to X; [3,N] in, [3,N] out."""
    dev = _device()
    _, _, al = rigid_transform_batch(_points(X, dev), _points(Y, dev), want_aligned=True)
    return al[0].cpu().numpy().T


def compute_similarity_transform(X, Y, compute_optimal_scale=False):
    """[transformation.py:48-97] MATLAB-style procrustes.  X targets [N,3], Y inputs [N,3].
    Returns d, Z [N,3], T [3,3], b, c [3] as the reference does (numpy fp64)."""
    dev = _device()
    Xd = torch.as_tensor(np.asarray(X, dtype=np.float64)).unsqueeze(0).contiguous().to(dev)
    Yd = torch.as_tensor(np.asarray(Y, dtype=np.float64)).unsqueeze(0).contiguous().to(dev)
    if Xd.dim() != 3 or Xd.shape[2] != 3 or Xd.shape != Yd.shape:
        raise ValueError('X and Y must both be [N,3]')
    d, Z, T, b, c = similarity_transform_batch(Xd, Yd, compute_optimal_scale)
    return (float(d[0]), Z[0].cpu().numpy(), T[0].cpu().numpy(), 1 if not compute_optimal_scale else float(b[0]),
            c[0].cpu().numpy())


def similarity_transform_batch(X, Y, compute_optimal_scale=False):
    """Batched form: X, Y CUDA fp64 [N,P,3] -> d [N], Z [N,P,3], T [N,3,3], b [N], c [N,3]."""
    n, p = X.shape[0], X.shape[1]
    d = torch.empty((n,), device=X.device, dtype=torch.float64)
    b = torch.empty((n,), device=X.device, dtype=torch.float64)
    Z = torch.empty((n, p, 3), device=X.device, dtype=torch.float64)
    T = torch.empty((n, 3, 3), device=X.device, dtype=torch.float64)
    c = torch.empty((n, 3), device=X.device, dtype=torch.float64)
    with torch.cuda.device(X.device):
        N.check(N.lib().egn_similarity_transform(N.ptr(X.contiguous()), N.ptr(Y.contiguous()), n, p,
                                                 1 if compute_optimal_scale else 0, N.ptr(d), N.ptr(b), N.ptr(Z),
                                                 N.ptr(T), N.ptr(c), N.current_stream()))
    return d, Z, T, b, c


def refine_with_predicted_bbox_batch(preds, observations, intrinsics, threshold=5., max_iter=0):
    """N instances of ``refine_with_predicted_bbox`` [tools/inference_legacy.py:518-547] in one launch.
    preds [N,P,3] (points 1.. relative to point 0), observations [N,P,2].
    Returns ok (CUDA bool [N]) and refined (CUDA fp64 [N,P,3], absolute coordinates)."""
    dev = preds.device if torch.is_tensor(preds) and preds.is_cuda else _device()
    x = torch.as_tensor(preds, dtype=torch.float64).to(dev).contiguous()
    u = torch.as_tensor(observations, dtype=torch.float64).to(dev).contiguous()
    n, p = x.shape[0], x.shape[1]
    if x.dim() != 3 or x.shape[2] != 3 or tuple(u.shape) != (n, p, 2):
        raise ValueError('preds must be [N,P,3] and observations [N,P,2]')
    Kn = np.asarray(intrinsics, dtype=np.float64)
    out = torch.empty_like(x)
    ok = torch.empty((n,), device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        N.check(N.lib().egn_refine_with_bbox(N.ptr(x), N.ptr(u), n, p, float(Kn[0, 0]), float(Kn[1, 1]),
                                             float(Kn[0, 2]), float(Kn[1, 2]), float(threshold), int(max_iter),
                                             N.ptr(out), N.ptr(ok), None, N.current_stream()))
    return ok.bool(), out


def refine_with_predicted_bbox(pred, observation, intrinsics, dist_coeffs, gts=None, threshold=5., ax=None):
    """[tools/inference_legacy.py:518-547] reference signature: pred [P,3], observation [P,2] ->
    (True, refined [3,P]) or (False, None) when the refined root moved more than ``threshold``."""
    if dist_coeffs is not None and np.any(np.asarray(dist_coeffs, dtype=np.float64) != 0):
        raise NotImplementedError('native pnp_refine supports zero lens distortion only')
    if ax is not None:
        raise NotImplementedError('plotting is outside the native path')
    ok, out = refine_with_predicted_bbox_batch(np.asarray(pred, dtype=np.float64)[None],
                                               np.asarray(observation, dtype=np.float64)[None], intrinsics, threshold)
    if not bool(ok[0]):
        return False, None
    return True, out[0].cpu().numpy().T


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
