"""CPU restatement of the crop geometry around ``HC`` (TEST INFRASTRUCTURE).

* ``enlarge_bbox`` / ``resize_bbox`` / ``modify_bbox`` -- ``libs/common/img_proc.py:411-459``.
* ``get_affine_transform`` -- ``libs/common/img_proc.py:26-64`` (+ ``get_dir``
  :84-91, ``get_3rd_point`` :80-82).  Upstream ends in ``cv2.getAffineTransform``
  on three float32 point pairs (third-party, OpenCV 3.4.2 pinned upstream,
  4.13 in this image); OpenCV's algorithm (imgproc/imgwarp.cpp,
  ``getAffineTransform``: a 6x6 linear system solved in double precision) is
  restated here with ``numpy.linalg.solve``.  Pinned against cv2 in
  ``tests/golden/make_golden.py``.
* ``affine_transform_modified`` -- ``libs/common/img_proc.py:71-78``.
* ``local_to_screen`` -- the per-instance loop of ``EgoNet.get_keypoints``
  ``libs/model/egonet.py:436-453``.
"""
import numpy as np

SIZE = 200.0  # img_proc.py:14


def enlarge_bbox(left, top, right, bottom, enlarge):
    width, height = right - left, bottom - top
    nw, nh = width * enlarge[0], height * enlarge[1]
    cx, cy = (left + right) / 2, (top + bottom) / 2
    return [cx - 0.5 * nw, cy - 0.5 * nh, cx + 0.5 * nw, cy + 0.5 * nh]


def resize_bbox(left, top, right, bottom, target_ar=1.):
    width, height = right - left, bottom - top
    cx, cy = (left + right) / 2, (top + bottom) / 2
    if height / width > target_ar:
        nw = height * (1 / target_ar)
        l, r, t, b = cx - 0.5 * nw, cx + 0.5 * nw, top, bottom
    else:
        nh = width * target_ar
        l, r, t, b = left, right, cy - 0.5 * nh, cy + 0.5 * nh
    return {'bbox': [l, t, r, b], 'c': np.array([cx, cy]),
            's': np.array([(r - l) / SIZE, (b - t) / SIZE])}


def modify_bbox(bbox, target_ar, enlarge=1.1):
    lb = enlarge_bbox(bbox[0], bbox[1], bbox[2], bbox[3], [enlarge, enlarge])
    return resize_bbox(lb[0], lb[1], lb[2], lb[3], target_ar=target_ar)


def _third(a, b):
    d = a - b
    return b + np.array([-d[1], d[0]], dtype=np.float32)


def _solve_affine(src, dst):
    """2x3 M with M @ [src_i, 1] = dst_i for 3 float32 point pairs (fp64 solve)."""
    A = np.zeros((6, 6), dtype=np.float64)
    b = np.zeros(6, dtype=np.float64)
    for i in range(3):
        A[i, 0:3] = [src[i, 0], src[i, 1], 1.0]
        A[i + 3, 3:6] = [src[i, 0], src[i, 1], 1.0]
        b[i], b[i + 3] = dst[i, 0], dst[i, 1]
    return np.linalg.solve(A, b).reshape(2, 3)


def get_affine_transform(center, scale, rot, output_size, inv=0):
    center = np.asarray(center, dtype=np.float64)
    scale_tmp = np.asarray(scale, dtype=np.float64) * SIZE
    src_w = scale_tmp[0]                       # only scale[0] is used (img_proc.py:42)
    dst_h, dst_w = output_size
    rot_rad = np.pi * rot / 180
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    p = [0, src_w * -0.5]
    src_dir = np.array([p[0] * cs - p[1] * sn, p[0] * sn + p[1] * cs])
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center
    src[1, :] = center + src_dir
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir
    src[2, :] = _third(src[0, :], src[1, :])
    dst[2, :] = _third(dst[0, :], dst[1, :])
    return _solve_affine(dst, src) if inv else _solve_affine(src, dst)


def affine_transform_modified(pts, t):
    new_pts = np.hstack([pts, np.ones((len(pts), 1))]).T
    return (t @ new_pts)[:2, :].T


def local_to_screen(coords, centers, scales, rots, resolution):
    """coords [N,K,2] float32 in (0,1) -> list of N fp64 [K,2] screen key-points."""
    width, height = resolution
    local = np.array(coords, dtype=np.float32, copy=True)
    local *= np.array(resolution).reshape(1, 1, 2)           # egonet.py:438 (stays float32)
    out = []
    for i in range(len(local)):
        t_inv = get_affine_transform(centers[i], scales[i], rots[i], (height, width), inv=1)
        out.append(affine_transform_modified(local[i], t_inv))
    return out
