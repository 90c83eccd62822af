"""B200-native drop-ins for the hot-path functions of the reference's
``libs/common/img_proc.py``: the three heat-map decoders and the crop geometry.

Decoders take a CUDA fp32 tensor ``[B, K, H, W]`` and return CUDA tensors (the
reference's ``get_max_preds`` / ``soft_arg_max_np`` take numpy arrays on the host
after a full D2H copy of the maps; here the maps never leave the device).
"""
import numpy as np
import torch

from ... import _native as N

SIZE = 200.0  # [img_proc.py:14]


def _maps(batch_heatmaps):
    if not isinstance(batch_heatmaps, torch.Tensor):
        raise TypeError('batch_heatmaps must be a CUDA torch.Tensor (no host path; the reference '
                        'version of this function takes numpy)')
    if batch_heatmaps.dim() != 4:
        raise AssertionError('batch_images should be 4-ndim')
    if not batch_heatmaps.is_cuda:
        raise RuntimeError('native decoders have no CPU path: input must be a CUDA tensor')
    return batch_heatmaps.detach().float().contiguous()


def get_max_preds(batch_heatmaps, return_index=False):
    """[img_proc.py:608-637] hard arg-max -> (preds [B,K,2] float32, maxvals [B,K,1])."""
    hm = _maps(batch_heatmaps)
    B, K, H, W = hm.shape
    preds = torch.empty((B, K, 2), device=hm.device, dtype=torch.float32)
    maxvals = torch.empty((B, K, 1), device=hm.device, dtype=torch.float32)
    idx = torch.empty((B, K), device=hm.device, dtype=torch.int32)
    with torch.cuda.device(hm.device):
        N.check(N.lib().egn_argmax2d(N.ptr(hm), B, K, H, W, N.ptr(idx), N.ptr(preds), N.ptr(maxvals),
                                     N.current_stream()))
    return (preds, maxvals, idx) if return_index else (preds, maxvals)


def _soft(batch_heatmaps, mode):
    hm = _maps(batch_heatmaps)
    B, K, H, W = hm.shape
    preds = torch.empty((B, K, 2), device=hm.device, dtype=torch.float32)
    maxvals = torch.empty((B, K, 1), device=hm.device, dtype=torch.float32)
    with torch.cuda.device(hm.device):
        N.check(N.lib().egn_soft_argmax2d(N.ptr(hm), B, K, H, W, mode, N.ptr(preds), N.ptr(maxvals),
                                          N.current_stream()))
    return preds, maxvals


def soft_arg_max(batch_heatmaps):
    """[img_proc.py:678-707] soft-max normalised expectation, raw-map max, no mask."""
    return _soft(batch_heatmaps, N.SOFTARGMAX_SOFTMAX)


def soft_arg_max_np(batch_heatmaps):
    """[img_proc.py:639-676] sum-normalised expectation, zeroed where max <= 0
    (does not modify its argument, unlike the reference's in-place division)."""
    return _soft(batch_heatmaps, N.SOFTARGMAX_SUM)


# ---- crop geometry (host, float64 numpy; a few flops per box) -----------------
def enlarge_bbox(left, top, right, bottom, enlarge):
    """[img_proc.py:437-451]"""
    w, h = (right - left) * enlarge[0], (bottom - top) * enlarge[1]
    cx, cy = (left + right) / 2, (top + bottom) / 2
    return [cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h]


def resize_bbox(left, top, right, bottom, target_ar=1.):
    """[img_proc.py:411-435]"""
    w, h = right - left, bottom - top
    cx, cy = (left + right) / 2, (top + bottom) / 2
    if h / w > target_ar:
        nw = h * (1 / target_ar)
        box = [cx - 0.5 * nw, top, cx + 0.5 * nw, bottom]
    else:
        nh = w * target_ar
        box = [left, cy - 0.5 * nh, right, cy + 0.5 * nh]
    return {'bbox': box, 'c': np.array([cx, cy]),
            's': np.array([(box[2] - box[0]) / SIZE, (box[3] - box[1]) / SIZE])}


def modify_bbox(bbox, target_ar, enlarge=1.1):
    """[img_proc.py:453-459]"""
    b = enlarge_bbox(bbox[0], bbox[1], bbox[2], bbox[3], [enlarge, enlarge])
    return resize_bbox(b[0], b[1], b[2], b[3], target_ar=target_ar)


def modify_bbox_batch(bboxes, target_ar, enlarge=1.1):
    """``modify_bbox`` [img_proc.py:411-459] for an [n,4] array of boxes at once (same float64 operations in
    the same order, so every row equals the per-box function bit for bit).
    Returns (bbox_resize [n,4], centers [n,2], scales [n,2])."""
    b = np.asarray(bboxes).reshape(-1, 4)          # arithmetic in the boxes' own dtype, as the per-box function
    left, top, right, bottom = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    w, h = (right - left) * enlarge, (bottom - top) * enlarge
    cx, cy = (left + right) / 2, (top + bottom) / 2
    left, top, right, bottom = cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h
    w, h = right - left, bottom - top
    cx, cy = (left + right) / 2, (top + bottom) / 2
    tall = h / w > target_ar
    nw, nh = h * (1 / target_ar), w * target_ar
    box = np.where(tall[:, None], np.stack([cx - 0.5 * nw, top, cx + 0.5 * nw, bottom], 1),
                   np.stack([left, cy - 0.5 * nh, right, cy + 0.5 * nh], 1))
    return box, np.stack([cx, cy], 1), np.stack([(box[:, 2] - box[:, 0]) / SIZE, (box[:, 3] - box[:, 1]) / SIZE], 1)


def get_affine_transform(center, scale, rot, output_size, shift=None, inv=0):
    """[img_proc.py:26-64] host-side crop affine for ``cv2.warpAffine`` (crop_instances
    is host code upstream as well).  Three float32 reference points per side exactly
    as upstream builds them, then the exact 3-point affine in float64 (what
    ``cv2.getAffineTransform`` computes).  The device path uses
    ``egn_local_to_screen`` for the inverse transform instead."""
    center = np.asarray(center, dtype=np.float64)
    src_w = float(np.asarray(scale, dtype=np.float64)[0]) * SIZE
    dst_h, dst_w = output_size
    ang = np.pi * rot / 180
    direction = np.array([src_w * 0.5 * np.sin(ang), src_w * -0.5 * np.cos(ang)])
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0] = center
    src[1] = center + direction
    dst[0] = [dst_w * 0.5, dst_h * 0.5]
    dst[1] = np.array([dst_w * 0.5, dst_h * 0.5]) + np.array([0, dst_w * -0.5], np.float32)
    for pts in (src, dst):
        d = pts[0] - pts[1]
        pts[2] = pts[1] + np.array([-d[1], d[0]], dtype=np.float32)
    a, b = (dst, src) if inv else (src, dst)
    A = np.hstack([a.astype(np.float64), np.ones((3, 1))])
    return np.linalg.solve(A, b.astype(np.float64)).T


def local_to_screen(coords, centers, scales, resolution, rots=None):
    """Batched form of the per-instance loop in ``EgoNet.get_keypoints``
    [egonet.py:436-453]: ``get_affine_transform(..., inv=1)`` [img_proc.py:26-64]
    followed by ``affine_transform_modified`` [img_proc.py:71-78].

    coords: CUDA fp32 [N,K,2] in (0,1); centers/scales: [N,2] fp64 (tensor or
    array); resolution = (width, height).  Returns CUDA fp64 [N,K,2]."""
    if not coords.is_cuda:
        raise RuntimeError('native local_to_screen has no CPU path')
    dev = coords.device
    c = coords.detach().float().contiguous()
    n, k = c.shape[0], c.shape[1]
    ce = torch.as_tensor(np.asarray(centers, dtype=np.float64) if not torch.is_tensor(centers) else centers,
                         dtype=torch.float64).to(dev).contiguous()
    sc = torch.as_tensor(np.asarray(scales, dtype=np.float64) if not torch.is_tensor(scales) else scales,
                         dtype=torch.float64).to(dev).contiguous()
    ro = None
    if rots is not None:
        ro = torch.as_tensor(np.asarray(rots, dtype=np.float64), dtype=torch.float64).to(dev).contiguous()
    out = torch.empty((n, k, 2), device=dev, dtype=torch.float64)
    with torch.cuda.device(dev):
        N.check(N.lib().egn_local_to_screen(N.ptr(c), N.ptr(ce), N.ptr(sc), N.ptr(ro), n, k,
                                            int(resolution[0]), int(resolution[1]), N.ptr(out),
                                            N.current_stream()))
    return out


def normalize_params(pth_trans):
    """(mean, std) of the ``Compose([ToTensor(), Normalize(mean, std)])`` the dataset hands to the
    model [car_instance.py:522-531]; (0, 1) per channel when the pipeline has no Normalize.
    Anything else has no device implementation and raises (there is no host fallback)."""
    mean, std = [0., 0., 0.], [1., 1., 1.]
    steps = getattr(pth_trans, 'transforms', None)
    if steps is None:
        raise NotImplementedError('pth_trans must be a Compose of ToTensor / Normalize (got %r)' % (pth_trans,))
    saw_tensor = False
    for t in steps:
        name = type(t).__name__
        if name == 'ToTensor':
            saw_tensor = True
        elif name == 'Normalize':
            mean, std = [float(v) for v in t.mean], [float(v) for v in t.std]
        else:
            raise NotImplementedError('transform %s has no device implementation' % name)
    if not saw_tensor:
        raise NotImplementedError('pth_trans without ToTensor has no device implementation')
    return mean, std


def crop_instances_device(images, image_of_crop, centers, scales, resolution, mean=None, std=None,
                          return_u8=False):
    """Every crop of a batch in one launch: the device form of ``EgoNet.crop_single_instance``
    [egonet.py:68-95] = ``get_affine_transform`` [img_proc.py:26-64] + ``cv2.warpAffine(INTER_LINEAR)``
    [egonet.py:85-89] + ``ToTensor``/``Normalize`` [car_instance.py:522-531].

    images: list of CUDA uint8 [H,W,3] RGB tensors (rows may be strided); image_of_crop: [N] ints;
    centers/scales: [N,2] fp64 (``modify_bbox`` output); resolution = (width, height).
    Returns CUDA fp32 [N,3,height,width] (and the uint8 [N,height,width,3] warp result with ``return_u8``)."""
    import ctypes
    if not images or not all(torch.is_tensor(i) and i.is_cuda for i in images):
        raise RuntimeError('native crop front-end needs CUDA uint8 images (there is no CPU path)')
    dev = images[0].device
    width, height = int(resolution[0]), int(resolution[1])
    table = (N.Image * len(images))()
    for k, im in enumerate(images):
        if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3 or im.stride(2) != 1 or im.stride(1) != 3:
            raise ValueError('image %d must be uint8 [H,W,3] with packed pixels' % k)
        table[k] = N.Image(im.data_ptr(), im.shape[0], im.shape[1], im.stride(0), 3)
    idx_host = np.asarray(image_of_crop, dtype=np.int32).reshape(-1)
    n = len(idx_host)
    if n and (idx_host.min() < 0 or idx_host.max() >= len(images)):
        raise ValueError('image_of_crop out of range')
    if torch.is_tensor(centers) and centers.is_cuda and torch.is_tensor(scales) and scales.is_cuda:
        ce = centers.to(torch.float64).contiguous().view(-1, 2)
        sc = scales.to(torch.float64).contiguous().view(-1, 2)
        ce_h = sc_h = np.zeros((0, 2))
    else:
        ce_h = np.ascontiguousarray(np.asarray(centers.cpu() if torch.is_tensor(centers) else centers, dtype=np.float64)).reshape(-1, 2)
        sc_h = np.ascontiguousarray(np.asarray(scales.cpu() if torch.is_tensor(scales) else scales, dtype=np.float64)).reshape(-1, 2)
        ce = sc = None
    if (ce is not None and (ce.shape[0] != n or sc.shape[0] != n)) or (ce is None and (len(ce_h) != n or len(sc_h) != n)):
        raise ValueError('centers / scales / image_of_crop disagree on the number of crops')
    # one host buffer, ONE H2D copy for the whole call: [centers | scales | image table | image_of_crop]
    tbytes = bytes(table)
    meta = np.zeros(ce_h.nbytes + sc_h.nbytes + len(tbytes) + idx_host.nbytes, dtype=np.uint8)
    o_sc, o_tab = ce_h.nbytes, ce_h.nbytes + sc_h.nbytes
    o_idx = o_tab + len(tbytes)                        # table entries are 24 bytes: 4-byte alignment holds
    meta[:o_sc] = ce_h.view(np.uint8).reshape(-1)
    meta[o_sc:o_tab] = sc_h.view(np.uint8).reshape(-1)
    meta[o_tab:o_idx] = np.frombuffer(tbytes, dtype=np.uint8)
    meta[o_idx:] = idx_host.view(np.uint8)
    meta_dev = torch.from_numpy(meta).to(dev, non_blocking=True)
    base = meta_dev.data_ptr()
    import ctypes as _ct
    p_ce = N.ptr(ce) if ce is not None else _ct.c_void_p(base)
    p_sc = N.ptr(sc) if sc is not None else _ct.c_void_p(base + o_sc)
    p_tab, p_idx = _ct.c_void_p(base + o_tab), _ct.c_void_p(base + o_idx)
    out = torch.empty((n, 3, height, width), device=dev, dtype=torch.float32)
    u8 = torch.empty((n, height, width, 3), device=dev, dtype=torch.uint8) if return_u8 else None
    m3 = (ctypes.c_float * 3)(*(mean if mean is not None else [0., 0., 0.]))
    s3 = (ctypes.c_float * 3)(*(std if std is not None else [1., 1., 1.]))
    with torch.cuda.device(dev):
        N.check(N.lib().egn_crop_instances(p_tab, len(images), p_idx, p_ce, p_sc, n,
                                           width, height, m3, s3, N.ptr(out), N.ptr(u8), N.current_stream()))
    return (out, u8) if return_u8 else out


def generate_target_batch(joints, joints_vis, parameters):
    """N samples of ``generate_target`` [img_proc.py:347-409] in one launch.
    joints [N,K,3] (crop pixels), joints_vis [N,K]; parameters as upstream (num_joints, target_type,
    input_size, heatmap_size, sigma).  Returns CUDA fp32 (target [N,K,hs[0],hs[1]], target_weight [N,K,1])."""
    import numpy as np
    if parameters['target_type'] != 'gaussian':
        raise AssertionError('Only support gaussian map now!')          # as upstream :365
    if parameters.get('use_different_joints_weight'):
        raise NotImplementedError('use_different_joints_weight has no device implementation')
    dev = torch.device('cuda', torch.cuda.current_device())
    j = torch.as_tensor(np.asarray(joints, dtype=np.float64) if not torch.is_tensor(joints) else joints,
                        dtype=torch.float64).to(dev).contiguous()
    v = torch.as_tensor(np.asarray(joints_vis, dtype=np.float32) if not torch.is_tensor(joints_vis) else joints_vis,
                        dtype=torch.float32).to(dev).contiguous()
    n, k = j.shape[0], j.shape[1]
    if j.dim() != 3 or j.shape[2] != 3 or tuple(v.shape) != (n, k) or k != parameters['num_joints']:
        raise ValueError('joints must be [N,K,3] and joints_vis [N,K] with K = num_joints')
    ins, hs = parameters['input_size'], parameters['heatmap_size']
    target = torch.empty((n, k, int(hs[0]), int(hs[1])), device=dev, dtype=torch.float32)
    weight = torch.empty((n, k), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        N.check(N.lib().egn_generate_target(N.ptr(j), N.ptr(v), n, k, int(ins[0]), int(ins[1]), int(hs[0]), int(hs[1]),
                                            float(parameters['sigma']), N.ptr(target), N.ptr(weight), N.current_stream()))
    return target, weight.unsqueeze(-1)


def to_npy(tensor):
    """[img_proc.py:722-728]"""
    return tensor if isinstance(tensor, np.ndarray) else tensor.data.cpu().numpy()
