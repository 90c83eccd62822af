"""CPU restatement of the per-crop pipeline ``EgoNet.forward`` minus image I/O
(TEST INFRASTRUCTURE): crops + boxes -> screen key-points -> 3D cuboid ->
Euler angles / translation / alpha.

Follows ``libs/model/egonet.py``: ``get_keypoints`` :424-467, ``lift_2d_to_3d``
:469-486, ``gather_lifting_results`` :297-339 (``get_6d_rep`` :279-295,
``get_observation_angle_{trans,proj}`` :203-236).  Image reading / cv2 warping
(``crop_instances`` :105-155) is outside the hot path;
the synthetic workload supplies already-cropped tensors plus the box geometry.
"""
import numpy as np
import torch

from . import affine_ref, hrnet_ref, lifter_ref, pose_ref

from egonet_b200.synth import KITTI_K  # noqa: E402,F401


def synth_crops(n, cfgs, seed=0):
    from egonet_b200 import synth
    return synth.crops(n, cfgs, seed)


def synth_boxes(n, cfgs, seed=2, enlarge=1.2):
    """Same recipe as ``egonet_b200.synth.boxes`` but through the ORACLE's own modify_bbox, so the
    two implementations of the box geometry are cross-checked by the golden pipeline test."""
    W, H = cfgs['heatmapModel']['input_size']
    target_ar = H / W
    rng = np.random.Generator(np.random.PCG64(seed))
    cx, cy = rng.uniform(50, 1190, n), rng.uniform(120, 330, n)
    bw, bh = rng.uniform(30, 400, n), rng.uniform(25, 250, n)
    records = []
    for i in range(n):
        box = [cx[i] - bw[i] / 2, cy[i] - bh[i] / 2, cx[i] + bw[i] / 2, cy[i] + bh[i] / 2]
        box = np.array(affine_ref.modify_bbox(box, target_ar=1.0, enlarge=enlarge)['bbox'])
        ret = affine_ref.modify_bbox(box, target_ar)
        records.append({'center': ret['c'], 'scale': ret['s'], 'rotation': 0.0,
                        'bbox': box, 'bbox_resize': ret['bbox']})
    return records


def run_pipeline(hc_sd, l_sd, stats, cfgs, crops, records, K=KITTI_K, ctx=hrnet_ref.Exact):
    """Returns dict of fp64 arrays: kpts_2d [N,2J], kpts_3d [N,J-1,3], euler [N,3],
    translation [N,3], alpha_trans [N], alpha_proj [N]; plus fp32 coords/maps."""
    out = hrnet_ref.hrnet_forward(hc_sd, cfgs, crops, ctx=ctx)
    maps, coords = out
    coords = coords.numpy()
    kpts = affine_ref.local_to_screen(coords, [r['center'] for r in records],
                                      [r['scale'] for r in records],
                                      [r['rotation'] for r in records],
                                      cfgs['heatmapModel']['input_size'])
    kpts_2d = np.concatenate([k.reshape(1, -1) for k in kpts], axis=0)
    kpts_3d = lifter_ref.lift_2d_to_3d(l_sd, cfgs, stats, kpts_2d)
    euler, trans = pose_ref.get_6d_rep(kpts_3d)
    a_trans = pose_ref.observation_angle_trans(euler, trans)
    a_proj = pose_ref.observation_angle_proj(euler, [k.reshape(1, -1) for k in kpts], K)
    return {'maps': maps.numpy(), 'coords': coords, 'kpts_2d': kpts_2d, 'kpts_3d': kpts_3d,
            'euler': euler, 'translation': trans, 'alpha_trans': a_trans, 'alpha_proj': a_proj}
