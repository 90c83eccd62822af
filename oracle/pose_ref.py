"""CPU restatement of the per-instance pose solve (TEST INFRASTRUCTURE).

* ``get_template``            -- ``EgoNet.get_template`` ``libs/model/egonet.py:238-263``
  with ``interp_dict['bbox12']`` from ``libs/dataset/KITTI/car_instance.py:63-70``.
* ``compute_rigid_transform`` -- ``libs/common/transformation.py:99-134`` (Kabsch;
  ``numpy.linalg.svd`` = LAPACK gesdd is third-party, same call as upstream).
* ``procrustes_transform`` / ``compute_similarity_transform`` -- ``transformation.py:136-141`` / ``:48-97``.
* ``euler_yxz``               -- ``EgoNet.kpts_to_euler`` ``egonet.py:265-277``:
  upstream calls ``scipy.spatial.transform.Rotation.from_matrix(R).as_euler('yxz')``
  (third-party; scipy 1.5.2 pinned upstream, 1.18.1 here) and reorders to
  [x, y, z]; the closed form of the extrinsic y-x-z decomposition
  ``R = Rz(c) Rx(b) Ry(a)`` is restated here and pinned against scipy by the
  golden vectors.
* ``get_6d_rep``              -- ``egonet.py:279-295``.
* ``observation_angle_trans`` / ``observation_angle_proj`` -- ``egonet.py:203-236``.
"""
import math

import numpy as np

# car_instance.py:63-70 (1-based indices of the 12 cuboid edges: 4 along h, 4 along l, 4 along w)
BBOX12_PARENTS = np.array([1, 3, 5, 7, 1, 2, 3, 4, 1, 2, 5, 6])
BBOX12_CHILDREN = np.array([2, 4, 6, 8, 5, 6, 7, 8, 3, 4, 7, 8])


def get_template(prediction, interp_coef=(0.332, 0.667)):
    """prediction [P,3] (P = 8 or 32) -> canonical cuboid [3,P]."""
    prediction = np.asarray(prediction)
    lines = prediction[BBOX12_PARENTS - 1] - prediction[BBOX12_CHILDREN - 1]
    lines = np.sqrt(np.sum(lines ** 2, axis=1))
    h, l, w = np.sum(lines[:4]) / 4, np.sum(lines[4:8]) / 4, np.sum(lines[8:]) / 4
    x = np.array([l, l, l, l, 0, 0, 0, 0], dtype=np.float64) - np.float32(l) / 2
    y = np.array([0, h, 0, h, 0, h, 0, h], dtype=np.float64) - np.float32(h)
    z = np.array([w, w, 0, 0, w, w, 0, 0], dtype=np.float64) - np.float32(w) / 2
    corners = np.array([x, y, z])
    if len(prediction) == 32:
        par, chi = corners[:, BBOX12_PARENTS - 1], corners[:, BBOX12_CHILDREN - 1]
        seg = chi - par
        corners = np.hstack([corners] + [par + c * seg for c in interp_coef])
    return corners


def compute_rigid_transform(X, Y, W=None):
    """Least-squares R, t with R X + t ~ Y for [3,N] point sets; W optional [N] or [N,N] weights
    (``transformation.py:112-122``: unweighted centroids, ``H = Xm W Ym^T``)."""
    cX = np.mean(X, axis=1, keepdims=True)
    cY = np.mean(Y, axis=1, keepdims=True)
    if W is None:
        H = (X - cX) @ (Y - cY).T
    else:
        W = np.asarray(W, dtype=np.float64)
        H = (X - cX) @ (np.diag(W) if W.ndim == 1 else W) @ (Y - cY).T
    U, S, Vt = np.linalg.svd(H)
    R = Vt.T @ U.T
    if np.linalg.det(R) < 0:
        Vt[-1, :] *= -1
        R = Vt.T @ U.T
    return R, -R @ cX + cY


def procrustes_transform(X, Y):
    """``transformation.py:136-141``: the rigid transform from X to Y applied to X ([3,N])."""
    R, t = compute_rigid_transform(X, Y)
    return R @ X + t


def compute_similarity_transform(X, Y, compute_optimal_scale=False):
    """``transformation.py:48-97`` (MATLAB procrustes): X targets [N,M], Y inputs [N,M] ->
    d, Z, T, b, c."""
    muX, muY = X.mean(0), Y.mean(0)
    X0, Y0 = X - muX, Y - muY
    ssX, ssY = (X0 ** 2.).sum(), (Y0 ** 2.).sum()
    normX, normY = np.sqrt(ssX), np.sqrt(ssY)
    X0, Y0 = X0 / normX, Y0 / normY
    U, s, Vt = np.linalg.svd(X0.T @ Y0, full_matrices=False)
    V = Vt.T
    sign = np.sign(np.linalg.det(V @ U.T))
    V[:, -1] *= sign
    s[-1] *= sign
    T = V @ U.T
    tr = s.sum()
    if compute_optimal_scale:
        b, d, Z = tr * normX / normY, 1 - tr ** 2, normX * tr * (Y0 @ T) + muX
    else:
        b, d, Z = 1, 1 + ssY / ssX - 2 * tr * normY / normX, normY * (Y0 @ T) + muX
    return d, Z, T, b, muX - b * (muY @ T)


def euler_yxz(R):
    """[x, y, z] angles of the extrinsic 'yxz' decomposition (away from gimbal lock)."""
    b = math.asin(max(-1.0, min(1.0, R[2, 1])))
    a = math.atan2(-R[2, 0], R[2, 2])
    c = math.atan2(-R[0, 1], R[1, 1])
    return np.array([b, a, c])


def get_6d_rep(predictions):
    predictions = np.asarray(predictions, dtype=np.float64)
    predictions = predictions.reshape(len(predictions), -1, 3)
    angles = np.zeros((len(predictions), 3))
    for i, pred in enumerate(predictions):
        R, _ = compute_rigid_transform(get_template(pred), pred.T)
        angles[i] = euler_yxz(R)
    return angles, predictions[:, 0, :]


def _wrap(alpha):
    while alpha > math.pi:
        alpha -= math.pi * 2
    while alpha < -math.pi:
        alpha += math.pi * 2
    return alpha


def observation_angle_trans(euler_angles, translations):
    out = euler_angles[:, 1].copy()
    for i in range(len(euler_angles)):
        out[i] = _wrap(euler_angles[i][1] - math.atan2(-translations[i][2], translations[i][0]) - 0.5 * math.pi)
    return out


def observation_angle_proj(euler_angles, kpts, K):
    """kpts: list of per-instance [1, 2J] (or [J,2]) arrays; uses element [0,0]."""
    f, cx = K[0, 0], K[0, 2]
    out = euler_angles[:, 1].copy()
    for i in range(len(euler_angles)):
        x3d = np.asarray(kpts[i]).reshape(-1)[0] - cx
        out[i] = _wrap(euler_angles[i][1] - math.atan2(-f, x3d) - 0.5 * math.pi)
    return out
