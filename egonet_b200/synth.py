"""Synthetic workload generators shared by bench.py, smoke() and the tests.

No checkpoints or datasets are reachable offline, so the measured workload is
seeded: model-section config dicts restating the reference YAML files, PCG64
weights for ``HC`` / ``L`` (numpy streams are stable across machines and torch
versions, unlike ``torch.manual_seed``), lifter statistics, N(0,1) crops and
KITTI-shaped detector boxes.
"""
import copy
from collections import OrderedDict

import numpy as np
import torch

# typical KITTI P2 intrinsics (upstream debug constant car_instance.py:933-935)
KITTI_K = np.array([[707.0493, 0.0, 604.0814], [0.0, 707.0493, 180.5066], [0.0, 0.0, 1.0]])


# ----------------------------------------------------------------------------- configs
def _stage(modules, channels):
    n = len(channels)
    return {'num_modules': modules, 'num_branches': n, 'block': 'basic',
            'num_blocks': [4] * n, 'num_channels': list(channels), 'fuse_method': 'sum'}


def make_cfgs(widths=(48, 96, 192, 384), input_size=(256, 256), heatmap_size=(64, 64), modules=(1, 4, 3),
              head_type='coordinates', num_joints=33, neurons=1024, lifter_blocks=2):
    """Model blocks of a reference YAML (configs/KITTI_inference:demo.yml:61-151 for the defaults)."""
    return {
        'FCModel': {'name': 'lifter', 'refine_3d': False, 'norm_twoD': False, 'num_blocks': lifter_blocks,
                    'input_size': 2 * num_joints, 'output_size': 3 * (num_joints - 1),
                    'num_neurons': neurons, 'dropout': 0.5, 'leaky': False},
        'heatmapModel': {
            'name': 'hrnet', 'add_xy': False, 'input_size': list(input_size), 'head_type': head_type,
            'pixel_shuffle': False, 'heatmap_size': list(heatmap_size), 'init_weights': True,
            'num_joints': num_joints, 'pretrained': '',
            'extra': {'pretrained_layers': ['*'], 'final_conv_kernel': 1,
                      'stage2': _stage(modules[0], widths[:2]), 'stage3': _stage(modules[1], widths[:3]),
                      'stage4': _stage(modules[2], widths[:4])},
        },
    }


def demo_cfgs(head_type='coordinates'):
    """HRNet-W48, 256x256 crops, 64x64 maps, 33 joints: every shipped inference config."""
    return make_cfgs(head_type=head_type)


def tiny_cfgs(head_type='coordinates'):
    """Same topology shrunk (W16, 128x128 crops, one module per stage) for fast tests."""
    return make_cfgs(widths=(16, 32, 64, 128), input_size=(128, 128), heatmap_size=(32, 32),
                     modules=(1, 1, 1), head_type=head_type, neurons=256, lifter_blocks=2)


def ped_cfgs():
    """W32, 192(w)x256(h) crops, 48x64 maps: the shape family of configs/KITTI_train_IGRs_Ped.yml."""
    return make_cfgs(widths=(32, 64, 128, 256), input_size=(192, 256), heatmap_size=(48, 64), modules=(1, 4, 3))


def clone(cfgs):
    return copy.deepcopy(cfgs)


# ----------------------------------------------------------------------------- weights
def hc_weights(spec, seed=1):
    """Seeded ``HC`` state dict for an ordered {key: shape} inventory (e.g. ``model.state_dict()``).

    Conv weights are uniform with variance 1/fan_in; every BatchNorm gets randomised affine
    parameters and running statistics so that folding is exercised; block-tail and fuse BNs are
    damped so activations stay O(1) through the ~90 sequential layers (logits std ~1, no sigmoid
    saturation)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = OrderedDict()
    for key, shape in spec.items():
        shape = tuple(shape.shape) if hasattr(shape, 'shape') else tuple(shape)
        owner, leaf = key.rsplit('.', 1)
        if leaf == 'num_batches_tracked':
            sd[key] = torch.tensor(0, dtype=torch.long)
            continue
        if len(shape) == 4:
            b = np.sqrt(3.0 / (shape[1] * shape[2] * shape[3]))
            arr = rng.uniform(-b, b, size=shape)
        elif leaf == 'running_var':
            arr = rng.uniform(0.5, 1.5, size=shape)
        elif leaf == 'running_mean':
            arr = rng.normal(0.0, 0.1, size=shape)
        elif leaf == 'weight':  # BN gamma
            damp = owner.endswith('bn2') or owner.endswith('bn3')
            is_block_tail = damp and ('branches' in owner or 'layer1' in owner or 'head2' in owner)
            gain = 0.35 if is_block_tail else (0.5 if 'fuse_layers' in owner else 1.0)
            arr = rng.uniform(0.5, 1.5, size=shape) * gain
        elif leaf == 'bias':
            arr = rng.normal(0.0, 0.1, size=shape)
        else:
            raise KeyError(key)
        sd[key] = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32))
    return sd


def lifter_weights(spec, seed=11):
    """Seeded ``L`` state dict for an ordered {key: shape} inventory."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = OrderedDict()
    for key, shape in spec.items():
        shape = tuple(shape.shape) if hasattr(shape, 'shape') else tuple(shape)
        leaf = key.rsplit('.', 1)[1]
        if leaf == 'num_batches_tracked':
            sd[key] = torch.tensor(0, dtype=torch.long)
            continue
        if len(shape) == 2:
            b = np.sqrt(3.0 / shape[1])
            arr = rng.uniform(-b, b, size=shape)
        elif leaf == 'running_var':
            arr = rng.uniform(0.5, 1.5, size=shape)
        elif leaf == 'running_mean':
            arr = rng.normal(0.0, 0.1, size=shape)
        elif leaf == 'weight':
            arr = rng.uniform(0.5, 1.5, size=shape)
        else:
            arr = rng.normal(0.0, 0.1, size=shape)
        sd[key] = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32))
    return sd


def lifter_stats(cfgs, seed=12, image_size=(1242, 375)):
    """Synthetic ``LS`` statistics dict (fp64 [1,n] arrays, keys as in LS.npy)."""
    fc = cfgs['FCModel']
    rng = np.random.Generator(np.random.PCG64(seed))
    nin, nout = fc['input_size'], fc['output_size']
    mean_in = np.empty((1, nin))
    mean_in[0, 0::2] = rng.uniform(0, image_size[0], nin // 2)
    mean_in[0, 1::2] = rng.uniform(0, image_size[1], nin // 2)
    return {'mean_in': mean_in, 'std_in': rng.uniform(20, 200, (1, nin)),
            'mean_out': rng.normal(0, 1, (1, nout)), 'std_out': rng.uniform(0.2, 2, (1, nout))}


def weights_digest(sd):
    """Order-dependent fp64 checksum of a state dict (pins regenerated weights)."""
    acc = 0.0
    for i, (k, v) in enumerate(sd.items()):
        if v.dtype.is_floating_point:
            a = v.double().flatten()
            w = torch.arange(1, a.numel() + 1, dtype=torch.float64) % 97 + 1.0
            acc += float((a * w).sum()) * ((i % 13) + 1)
    return acc


# ----------------------------------------------------------------------------- inputs
def crops(n, cfgs, seed=0):
    """Seeded N(0,1) crops [n,3,H,W] float32 (ImageNet-normalised images are ~N(0,1))."""
    W, H = cfgs['heatmapModel']['input_size']
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(rng.standard_normal((n, 3, H, W), dtype=np.float32))


def boxes(n, cfgs, seed=2, enlarge=1.2):
    """KITTI-shaped detector boxes -> per-instance records with center/scale derived as the reference
    does (tools/inference.py:113-116 enlarges by 1.2, then egonet.py:141-142 by 1.1 + aspect fix)."""
    from .libs.common.img_proc import modify_bbox
    W, H = cfgs['heatmapModel']['input_size']
    target_ar = H / W
    rng = np.random.Generator(np.random.PCG64(seed))
    cx, cy = rng.uniform(50, 1190, n), rng.uniform(120, 330, n)
    bw, bh = rng.uniform(30, 400, n), rng.uniform(25, 250, n)
    records = []
    for i in range(n):
        box = [cx[i] - bw[i] / 2, cy[i] - bh[i] / 2, cx[i] + bw[i] / 2, cy[i] + bh[i] / 2]
        box = np.array(modify_bbox(box, target_ar=1.0, enlarge=enlarge)['bbox'])
        ret = modify_bbox(box, target_ar)
        records.append({'center': ret['c'], 'scale': ret['s'], 'rotation': 0.0,
                        'bbox': box, 'bbox_resize': ret['bbox']})
    return records
