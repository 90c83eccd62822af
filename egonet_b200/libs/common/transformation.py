"""B200-native drop-in for the hot-path part of the reference's
``libs/common/transformation.py``: the batched pose solve.

The reference solves one instance at a time on the host (``compute_rigid_transform``
[transformation.py:99-134] inside ``EgoNet.get_6d_rep`` [egonet.py:279-295], then
scipy Euler angles and the observation angle [egonet.py:203-236]).  ``pose_solve``
does all of it for N instances in one fp64 kernel launch.
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
