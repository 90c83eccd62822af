"""CPU restatement of the crop front-end (TEST INFRASTRUCTURE, never imported by the product).

What the reference does per box (``EgoNet.crop_single_instance`` ``libs/model/egonet.py:68-95``,
called from ``crop_instances`` ``:105-155``):

    trans    = get_affine_transform(c, s, 0, (height, width))           img_proc.py:26-64
    instance = cv2.warpAffine(img, trans, (width, height), flags=cv2.INTER_LINEAR)   egonet.py:85-89
    instance = pth_trans(instance)      # transforms.ToTensor() + Normalize(mean, std)   car_instance.py:522-531

``cv2.warpAffine`` is third-party arithmetic that is NOT under /root/reference: OpenCV (3.4.2 pinned
upstream ``docs/spec-list.txt``, 4.13.0 in this image, built with ALGO_HINT_ACCURATE).  Its published
algorithm for 8-bit images (modules/imgproc/src/imgwarp.cpp: ``warpAffine`` -> ``WarpAffineInvoker`` ->
``remap`` with ``INTER_LINEAR``, unchanged between those versions) is restated here:

  * the 2x3 matrix is inverted in double precision (no WARP_INVERSE_MAP);
  * per destination column x: ``adelta = cvRound(M[0]*x*1024)``, ``bdelta = cvRound(M[3]*x*1024)``;
    per row y: ``X0 = cvRound((M[1]*y + M[2])*1024) + 16``, ``Y0 = cvRound((M[4]*y + M[5])*1024) + 16``
    (AB_BITS = 10, round_delta = AB_SCALE / INTER_TAB_SIZE / 2);
  * ``X = (X0 + adelta) >> 5``: integer source column ``X >> 5`` (saturated to int16) and a 5-bit
    fraction ``X & 31``; same for Y;
  * the four taps are blended with ``BilinearTab_i`` -- 15-bit fixed-point weights
    ``saturate_cast<short>((1-fy)(1-fx) * 32768)`` ... whose sum is forced to 32768 -- as
    ``(sum w_i * p_i + (1 << 14)) >> 15``; taps outside the image read BORDER_CONSTANT 0.
    For this 2x2 table every weight is an exact integer ``(32-fy)(32-fx)*32`` etc.; the only entry that
    saturates is (0,0): 32768 -> 32767, and the table's sum fix-up gives the missing 1 to w11.

``ToTensor`` on a uint8 HWC array is ``permute -> float32 -> / 255``; ``Normalize`` is
``(x - mean) / std`` with float32 mean/std (torchvision/transforms/functional.py, _functional_tensor.py).

Pinned by ``tests/golden/make_golden.py::golden_crop`` against cv2 + torchvision executed through the
reference's own ``crop_single_instance`` (``tests/golden/crop.npz``).
"""
import numpy as np

from . import affine_ref

AB_BITS = 10
AB_SCALE = 1 << AB_BITS
INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
REMAP_COEF_BITS = 15
ROUND_DELTA = AB_SCALE // INTER_TAB_SIZE // 2


def bilinear_tab():
    """cv::initInterTab2D(INTER_LINEAR, fixpt=true) -> int32 [32(fy), 32(fx), 2(dy), 2(dx)].

    Restated with OpenCV's float arithmetic and its sum fix-up loop (imgwarp.cpp).  With ksize = 2 the
    loop scans flat indices 3..6 of the entry: index 3 is w11, 4..6 belong to the NEXT (still zero) entry,
    so a deficit (diff < 0) is added to w11 while a surplus would be written into the next entry and
    later overwritten (it never occurs for the bilinear table)."""
    tab1 = np.zeros((INTER_TAB_SIZE, 2), np.float32)
    scale = np.float32(1.0 / INTER_TAB_SIZE)
    for i in range(INTER_TAB_SIZE):
        x = np.float32(i) * scale
        tab1[i] = (np.float32(1.0) - x, x)
    itab = np.zeros((INTER_TAB_SIZE, INTER_TAB_SIZE, 2, 2), np.int32)
    for i in range(INTER_TAB_SIZE):
        for j in range(INTER_TAB_SIZE):
            isum = 0
            for k1 in range(2):
                for k2 in range(2):
                    v = np.float32(tab1[i, k1] * tab1[j, k2]) * np.float32(1 << REMAP_COEF_BITS)
                    iv = int(min(32767, max(-32768, int(np.rint(v)))))
                    itab[i, j, k1, k2] = iv
                    isum += iv
            diff = isum - (1 << REMAP_COEF_BITS)
            if diff < 0 or (diff > 0 and itab[i, j, 1, 1] <= 0):
                itab[i, j, 1, 1] -= diff
    return itab


_ITAB = None


def invert_affine(M):
    """The in-place inversion at the top of cv::warpAffine (double precision, this operation order)."""
    M = np.asarray(M, np.float64)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    m0, m1, m3, m4 = A11, M[0, 1] * (-D), M[1, 0] * (-D), A22
    b1 = -m0 * M[0, 2] - m1 * M[1, 2]
    b2 = -m3 * M[0, 2] - m4 * M[1, 2]
    return np.array([[m0, m1, b1], [m3, m4, b2]])


def warp_positions(M, dsize):
    """Fixed-point source positions for every destination pixel: (sx, sy, fx, fy) int arrays [H, W]."""
    W, H = dsize
    Mi = invert_affine(M)
    x = np.arange(W, dtype=np.float64)
    y = np.arange(H, dtype=np.float64)
    adelta = np.rint(Mi[0, 0] * x * AB_SCALE).astype(np.int64)
    bdelta = np.rint(Mi[1, 0] * x * AB_SCALE).astype(np.int64)
    X0 = np.rint((Mi[0, 1] * y + Mi[0, 2]) * AB_SCALE).astype(np.int64) + ROUND_DELTA
    Y0 = np.rint((Mi[1, 1] * y + Mi[1, 2]) * AB_SCALE).astype(np.int64) + ROUND_DELTA
    X = (X0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    sx = np.clip(X >> INTER_BITS, -32768, 32767)
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    return sx, sy, X & (INTER_TAB_SIZE - 1), Y & (INTER_TAB_SIZE - 1)


def warp_affine(img, M, dsize):
    """cv2.warpAffine(img, M, dsize, flags=cv2.INTER_LINEAR) for a uint8 [h, w, C] image."""
    global _ITAB
    if _ITAB is None:
        _ITAB = bilinear_tab()
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim == 3
    ih, iw = img.shape[:2]
    sx, sy, fx, fy = warp_positions(M, dsize)
    w = _ITAB[fy, fx].astype(np.int64)                       # [H, W, 2, 2]
    acc = np.full(sx.shape + (img.shape[2],), 1 << (REMAP_COEF_BITS - 1), np.int64)
    for dy in range(2):
        for dx in range(2):
            yy, xx = sy + dy, sx + dx
            ok = (yy >= 0) & (yy < ih) & (xx >= 0) & (xx < iw)
            p = img[np.clip(yy, 0, ih - 1), np.clip(xx, 0, iw - 1)].astype(np.int64)
            acc += np.where(ok[..., None], p, 0) * w[..., dy, dx][..., None]
    return np.clip(acc >> REMAP_COEF_BITS, 0, 255).astype(np.uint8)


def to_tensor_normalize(crop_u8, mean, std):
    """transforms.Compose([ToTensor(), Normalize(mean, std)]) on a uint8 HWC crop -> float32 [C, H, W]."""
    x = np.ascontiguousarray(crop_u8.transpose(2, 0, 1)).astype(np.float32) / np.float32(255)
    m = np.asarray(mean, np.float32)[:, None, None]
    s = np.asarray(std, np.float32)[:, None, None]
    return ((x - m) / s).astype(np.float32)


def crop_single_instance(img, bbox, resolution, mean=None, std=None):
    """``EgoNet.crop_single_instance`` (egonet.py:68-95) without the xy maps.
    resolution = (width, height).  Returns (uint8 crop [h, w, 3], float32 [3, h, w] or None, c, s)."""
    width, height = resolution
    ret = affine_ref.modify_bbox(np.asarray(bbox, np.float64), height / width)
    trans = affine_ref.get_affine_transform(ret['c'], ret['s'], 0., (height, width))
    crop = warp_affine(img, trans, (int(width), int(height)))
    norm = to_tensor_normalize(crop, mean, std) if mean is not None else None
    return crop, norm, ret['c'], ret['s']


def synth_image(height, width, seed):
    """Deterministic image with edges, gradients and texture (PCG64; no external data)."""
    g = np.random.Generator(np.random.PCG64(seed))
    yy, xx = np.mgrid[0:height, 0:width]
    img = np.zeros((height, width, 3), np.float64)
    for c in range(3):
        img[..., c] = 96 + 80 * np.sin(xx / (7.0 + 3 * c) + c) * np.cos(yy / (5.0 + 2 * c))
    for _ in range(40):                                         # opaque rectangles: hard edges
        x0, y0 = int(g.integers(0, width)), int(g.integers(0, height))
        w, h = int(g.integers(4, max(5, width // 4))), int(g.integers(4, max(5, height // 3)))
        img[y0:y0 + h, x0:x0 + w] = g.integers(0, 256, 3)
    img += g.normal(0, 12, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)
