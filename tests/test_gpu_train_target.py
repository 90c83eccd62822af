"""GPU parity of the Gaussian target generator (``egn_generate_target``) against targets made by the reference's
``generate_target`` (tests/golden/train_tiny.npz) and against the CPU oracle on the BASELINE configs[3] shape."""
import numpy as np
import pytest
import torch

from oracle import configs, train_ref

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from egonet_b200.libs.common import img_proc


def test_generate_target_vs_reference_golden(golden):
    g = golden('train_tiny.npz')
    hm = configs.tiny_cfgs('heatmap')['heatmapModel']
    params = {'num_joints': hm['num_joints'], 'target_type': 'gaussian', 'input_size': hm['input_size'],
              'heatmap_size': hm['heatmap_size'], 'sigma': int(g['sigma'])}
    tgt, wgt = img_proc.generate_target_batch(g['joints'], g['vis'], params)
    tgt, wgt = tgt.cpu().numpy(), wgt.cpu().numpy()
    np.testing.assert_array_equal(wgt, g['target_weight'])                 # visibility weights: exact
    np.testing.assert_array_equal(tgt == 0, g['target'] == 0)             # dot geometry: exact
    np.testing.assert_allclose(tgt, g['target'], rtol=1e-6, atol=0)       # CUDA expf (2 ulp) vs numpy float32 exp (1 ulp)


def test_generate_target_config3_shape_vs_oracle():
    """128 samples x 33 joints on 64 x 64 maps, sigma 1 (BASELINE configs[3]); joints outside, on the border,
    invisible; N = 0."""
    rng = np.random.Generator(np.random.PCG64(81))
    n, k = 128, 33
    joints = np.concatenate([rng.uniform(-40, 296, (n, k, 2)), np.ones((n, k, 1))], 2)
    joints[0, 0, :2] = (0.0, 0.0)
    joints[0, 1, :2] = (255.9, 255.9)
    joints[0, 2, :2] = (-1.9, 128.0)            # int() truncates towards zero
    vis = (rng.uniform(0, 1, (n, k)) > 0.15).astype(np.float32)
    vis[0, :3] = 1.0
    params = {'num_joints': k, 'target_type': 'gaussian', 'input_size': [256, 256], 'heatmap_size': [64, 64], 'sigma': 1}
    tgt, wgt = img_proc.generate_target_batch(joints, vis, params)
    tgt, wgt = tgt.cpu().numpy(), wgt.cpu().numpy()
    for b in (0, 1, 77, 127):
        t, w = train_ref.generate_target(joints[b], vis[b], k, params['input_size'], params['heatmap_size'], 1)
        np.testing.assert_array_equal(wgt[b], w)
        np.testing.assert_array_equal(tgt[b] == 0, t == 0)
        np.testing.assert_allclose(tgt[b], t, rtol=1e-6, atol=0)
    assert tgt.max() == 1.0 and (wgt.sum() < vis.sum())
    empty, _ = img_proc.generate_target_batch(np.zeros((0, k, 3)), np.zeros((0, k), np.float32), params)
    assert empty.shape == (0, k, 64, 64)
