"""CPU restatement of the reprojection refinement (TEST INFRASTRUCTURE, never imported by the product).

``pnp_refine`` -- ``libs/common/transformation.py:143-157``:

    (success, R, T) = cv2.solvePnP(prediction, observation, intrinsics, dist_coeffs, flags=cv2.SOLVEPNP_ITERATIVE)
    refined = cv2.Rodrigues(R)[0] @ prediction.T + T                  # [3, P]

``cv2.solvePnP`` is third-party arithmetic that is not under /root/reference (OpenCV 3.4.2 pinned upstream in
``docs/spec-list.txt``, 4.13.0 in this image).  Its published algorithm (modules/calib3d:
``cvFindExtrinsicCameraParams2`` driving ``CvLevMarq``; unchanged between those versions) is restated in numpy:

  1. normalise the image points with the intrinsics (zero distortion);
  2. planarity test: singular values of the scatter matrix of the object points, planar if W[2]/W[1] < 1e-3
     (the planar homography branch is NOT restated -- ``solve_pnp_iterative`` raises for planar input);
  3. DLT: rows ``[X 1 0 -xX -x]`` / ``[0 X 1 -yX -y]``, right singular vector of the smallest singular value
     of ``L^T L``, sign so that det(RR) > 0, rotation = polar factor ``U V^T``, translation scaled by
     ``|R|_F / |RR|_F``, rotation vector by the matrix->vector branch of ``cvRodrigues2``;
  4. Levenberg-Marquardt (``CvLevMarq``): pixel residuals, ``lambda = 10^k`` with k0 = -3, the diagonal of
     ``J^T J`` multiplied by ``1 + lambda``, SVD solve, the step is re-taken with k += 1 (up to 16) while the
     residual norm grows, k -= 1 on acceptance, at most 20 accepted steps, stop when
     ``|p - p_prev| / |p_prev| < FLT_EPSILON``.  The Jacobian w.r.t. the rotation vector is written with the
     left Jacobian of SO(3): ``d(R X)/dr = -[R X]_x J_l(r)`` (identical to OpenCV's analytic dR/dr).

Pinned by ``tests/golden/make_golden.py::golden_pnp`` against the reference's own ``pnp_refine`` (cv2 executed).
Where cv2 stops on its 20-iteration cap without having converged (wild DLT starts), the iteration is chaotic and
two correct implementations agree only loosely; the golden file records cv2's convergence so tests can tell.
"""
import numpy as np

FLT_EPSILON = float(np.finfo(np.float32).eps)


def hat(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.]])


def rodrigues(r):
    th = np.linalg.norm(r)
    if th < np.finfo(np.float64).eps:
        return np.eye(3)
    k = r / th
    c, s = np.cos(th), np.sin(th)
    return c * np.eye(3) + (1 - c) * np.outer(k, k) + s * hat(k)


def left_jacobian(r):
    th = np.linalg.norm(r)
    if th < 1e-8:
        return np.eye(3) + 0.5 * hat(r)
    Rx = hat(r)
    return np.eye(3) + (1 - np.cos(th)) / th ** 2 * Rx + (th - np.sin(th)) / th ** 3 * Rx @ Rx


def rotation_to_vector(R):
    """matrix -> vector branch of cvRodrigues2."""
    r = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = np.sqrt((r @ r) * 0.25)
    c = min(1.0, max(-1.0, (np.trace(R) - 1) * 0.5))
    th = np.arccos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        rx = np.sqrt(max((R[0, 0] + 1) * 0.5, 0))
        ry = np.sqrt(max((R[1, 1] + 1) * 0.5, 0)) * (-1 if R[0, 1] < 0 else 1)
        rz = np.sqrt(max((R[2, 2] + 1) * 0.5, 0)) * (-1 if R[0, 2] < 0 else 1)
        if abs(rx) < abs(ry) and abs(rx) < abs(rz) and (R[1, 2] > 0) != (ry * rz > 0):
            rz = -rz
        v = np.array([rx, ry, rz])
        return v * (th / np.linalg.norm(v))
    return r * (0.5 / s * th)


def project(X, r, t, K, want_jac=False):
    R = rodrigues(r)
    Y = X @ R.T
    Xc = Y + t
    z = Xc[:, 2]
    x, y = Xc[:, 0] / z, Xc[:, 1] / z
    uv = np.stack([K[0, 0] * x + K[0, 2], K[1, 1] * y + K[1, 2]], 1)
    if not want_jac:
        return uv, None
    Jl = left_jacobian(r)
    J = np.zeros((2 * len(X), 6))
    for i in range(len(X)):
        dp = np.array([[K[0, 0] / z[i], 0, -K[0, 0] * x[i] / z[i]], [0, K[1, 1] / z[i], -K[1, 1] * y[i] / z[i]]])
        J[2 * i:2 * i + 2, :3] = dp @ (-hat(Y[i])) @ Jl
        J[2 * i:2 * i + 2, 3:] = dp
    return uv, J


def is_planar(X):
    Mc = X.mean(0)
    W = np.linalg.svd((X - Mc).T @ (X - Mc), compute_uv=False)
    return W[2] / W[1] < 1e-3


def dlt_init(X, uv, K):
    xn = (uv - K[:2, 2]) * np.array([1 / K[0, 0], 1 / K[1, 1]])
    P = len(X)
    L = np.zeros((2 * P, 12))
    for i in range(P):
        x, y = -xn[i]
        M = X[i]
        L[2 * i] = [M[0], M[1], M[2], 1, 0, 0, 0, 0, x * M[0], x * M[1], x * M[2], x]
        L[2 * i + 1] = [0, 0, 0, 0, M[0], M[1], M[2], 1, y * M[0], y * M[1], y * M[2], y]
    _, V = np.linalg.eigh(L.T @ L)
    RRt = V[:, 0].reshape(3, 4).copy()
    if np.linalg.det(RRt[:, :3]) < 0:
        RRt = -RRt
    RR, tt = RRt[:, :3], RRt[:, 3]
    U, _, Vt = np.linalg.svd(RR)
    R = U @ Vt
    return rotation_to_vector(R), tt * (np.linalg.norm(R) / np.linalg.norm(RR))


def solve_pnp_iterative(X, uv, K, max_iter=20, eps=FLT_EPSILON):
    """-> (rvec [3], tvec [3], accepted iterations, final residual norm in pixels)."""
    X, uv, K = np.asarray(X, np.float64), np.asarray(uv, np.float64), np.asarray(K, np.float64)
    if len(X) < 6:
        raise ValueError('DLT needs at least 6 points')
    if is_planar(X):
        raise NotImplementedError('planar object points: OpenCV switches to a homography start (not restated)')
    p = np.concatenate(dlt_init(X, uv, K))
    lam10, iters, prev_err = -3, 0, None
    while True:
        uvp, J = project(X, p[:3], p[3:], K, True)
        err = (uvp - uv).reshape(-1)
        if iters == 0:
            prev_err = np.linalg.norm(err)
        JtJ, JtE, prev = J.T @ J, J.T @ err, p.copy()
        while True:
            A = JtJ.copy()
            A[np.diag_indices(6)] *= 1 + 10.0 ** lam10
            p = prev - np.linalg.lstsq(A, JtE, rcond=None)[0]
            err_n = np.linalg.norm(project(X, p[:3], p[3:], K)[0] - uv)
            if err_n > prev_err:
                lam10 += 1
                if lam10 <= 16:
                    continue
            break
        lam10 = max(lam10 - 1, -16)
        iters += 1
        if iters >= max_iter or np.linalg.norm(p - prev) / np.linalg.norm(prev) < eps:
            break
        prev_err = err_n
    return p[:3], p[3:], iters, err_n


def pnp_refine(prediction, observation, intrinsics, dist_coeffs=None):
    """transformation.py:143-157 -> refined [3, P]."""
    if dist_coeffs is not None and np.any(np.asarray(dist_coeffs) != 0):
        raise NotImplementedError('lens distortion is not restated')
    r, t, _, _ = solve_pnp_iterative(prediction, observation, intrinsics)
    return rodrigues(r) @ np.asarray(prediction, np.float64).T + t[:, None]


def refine_with_predicted_bbox(pred, observation, intrinsics, threshold=5.):
    """``tools/inference_legacy.py:518-547``: pred [P,3] with points 1.. relative to point 0 ->
    (ok, refined [3,P] or None)."""
    box = np.array(pred, dtype=np.float64)
    box[1:, :] += box[0, :].reshape(1, 3)
    refined = pnp_refine(box, observation, intrinsics, None)
    dist = np.sqrt(np.sum((refined[:, 0] - box[0, :]) ** 2))
    return (False, None) if dist > threshold else (True, refined)


def synth_cases(n, seed, noise_3d=0.02, noise_px=0.4, offset=0.8, points=9):
    """Seeded (prediction [n,P,3], observation [n,P,2]) pairs shaped like the reference's use: a cuboid
    (centre + 8 corners, or the 32-point cuboid + centre) seen by a KITTI camera; the prediction is the true
    box displaced and slightly deformed, the observation its noisy projection."""
    from .egonet_ref import KITTI_K
    g = np.random.Generator(np.random.PCG64(seed))
    preds, obs = [], []
    for _ in range(n):
        l, h, w = g.uniform(3, 5), g.uniform(1.3, 2), g.uniform(1.4, 2)
        c = np.array([[l / 2, 0, w / 2], [l / 2, -h, w / 2], [l / 2, 0, -w / 2], [l / 2, -h, -w / 2],
                      [-l / 2, 0, w / 2], [-l / 2, -h, w / 2], [-l / 2, 0, -w / 2], [-l / 2, -h, -w / 2]])
        if points == 33:
            par = np.array([1, 3, 5, 7, 1, 2, 3, 4, 1, 2, 5, 6]) - 1
            chi = np.array([2, 4, 6, 8, 5, 6, 7, 8, 3, 4, 7, 8]) - 1
            seg = c[chi] - c[par]
            c = np.vstack([c, c[par] + 0.332 * seg, c[par] + 0.667 * seg])
        c = np.vstack([c.mean(0), c])
        Rg = rodrigues(np.array([g.uniform(-0.1, 0.1), g.uniform(-np.pi, np.pi), g.uniform(-0.1, 0.1)]))
        tg = np.array([g.uniform(-15, 15), g.uniform(0.5, 2.5), g.uniform(6, 60)])
        Xgt = c @ Rg.T + tg
        preds.append(Xgt + g.normal(0, noise_3d, Xgt.shape) + g.normal(0, offset, 3))
        uv = Xgt @ KITTI_K.T
        obs.append(uv[:, :2] / uv[:, 2:3] + g.normal(0, noise_px, (len(c), 2)))
    return np.array(preds), np.array(obs)
