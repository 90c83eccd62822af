"""Overlaps of the KITTI object evaluation on the device.

The reference ships the official C++ evaluator (``tools/kitti-eval/evaluate_object_3d_offline.cpp``), which walks
every (detection, ground truth) pair of every frame on one host thread through Boost.Geometry polygon operations
[``groundBoxOverlap`` :293-314, ``box3DOverlap`` :317-344, ``imageBoxOverlap`` :224-262].  ``box_overlaps`` /
``image_box_overlaps`` compute the full [D, G] matrices of a frame in one launch each (``egn_box_overlaps``,
``egn_image_box_overlaps``); the matching / AP bookkeeping around them stays host code.
"""
import numpy as np
import torch

from ... import _native as N


def _dev(a, cols):
    t = torch.as_tensor(np.asarray(a.cpu() if torch.is_tensor(a) else a, dtype=np.float64) if not (torch.is_tensor(a) and a.is_cuda) else a,
                        dtype=torch.float64)
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError('native evaluation overlaps have no CPU path')
        t = t.cuda()
    t = t.contiguous().view(-1, cols)
    return t


def box_overlaps(dets, gts, criterion=-1, want=('ground', 'box3d')):
    """dets [D,7], gts [G,7] rows = (ry, h, w, l, x, y_bottom, z) -> dict of CUDA fp64 [D,G] overlap matrices."""
    d, g = _dev(dets, 7), _dev(gts, 7)
    D, G = d.shape[0], g.shape[0]
    out = {k: torch.empty((D, G), device=d.device, dtype=torch.float64) for k in want}
    with torch.cuda.device(d.device):
        N.check(N.lib().egn_box_overlaps(N.ptr(d), N.ptr(g), D, G, int(criterion), N.ptr(out.get('ground')),
                                         N.ptr(out.get('box3d')), N.current_stream()))
    return out


def image_box_overlaps(dets, gts, criterion=-1):
    """dets [D,4], gts [G,4] = (x1, y1, x2, y2) -> CUDA fp64 [D,G]."""
    d, g = _dev(dets, 4), _dev(gts, 4)
    out = torch.empty((d.shape[0], g.shape[0]), device=d.device, dtype=torch.float64)
    with torch.cuda.device(d.device):
        N.check(N.lib().egn_image_box_overlaps(N.ptr(d), N.ptr(g), d.shape[0], g.shape[0], int(criterion), N.ptr(out),
                                               N.current_stream()))
    return out


def boxes_from_annotations(annots):
    """Rows (ry, h, w, l, x, y, z) from the instance dictionaries ``csv_read_annot`` returns
    (``dimensions`` = [l, h, w], ``locations`` = [x, y, z], car_instance.py:803-826)."""
    return np.array([[a['rot_y'], a['dimensions'][1], a['dimensions'][2], a['dimensions'][0]] + list(a['locations'])
                     for a in annots], dtype=np.float64).reshape(-1, 7)
