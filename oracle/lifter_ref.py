"""CPU restatement of the 2D->3D lifter ``L`` and its (de)normalisation (TEST INFRASTRUCTURE).

* ``state_dict_spec`` / ``lifter_forward`` -- ``libs/model/FCmodel.py:45-105``
  (``FCModel``) and :9-43 (``ResidualBlock``), eval mode (Dropout = identity,
  BatchNorm1d uses running statistics), ReLU only (``leaky: False`` in every
  shipped config).
* ``normalize_1d`` / ``unnormalize_1d`` -- ``libs/dataset/normalization/operations.py:21-52``
  (``individual=False`` branch, the only one the inference path takes).
* ``lift_2d_to_3d`` -- ``EgoNet.lift_2d_to_3d`` ``libs/model/egonet.py:469-486``
  including its dtype chain fp64 -> fp32 -> L -> fp64.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def _bn(spec, p, c):
    for leaf in ('weight', 'bias', 'running_mean', 'running_var'):
        spec[p + '.' + leaf] = (c,)
    spec[p + '.num_batches_tracked'] = ()


def state_dict_spec(cfgs):
    fc = cfgs['FCModel']
    n, nin, nout = fc['num_neurons'], fc['input_size'], fc['output_size']
    spec = OrderedDict()
    spec['w1.weight'] = (n, nin)
    spec['w1.bias'] = (n,)
    _bn(spec, 'batch_norm1', n)
    for i in range(fc['num_blocks']):
        p = 'res_blocks.%d' % i
        spec[p + '.w1.weight'] = (n, n)
        spec[p + '.w1.bias'] = (n,)
        _bn(spec, p + '.batch_norm1', n)
        spec[p + '.w2.weight'] = (n, n)
        spec[p + '.w2.bias'] = (n,)
        _bn(spec, p + '.batch_norm2', n)
    spec['w2.weight'] = (nout, n)
    spec['w2.bias'] = (nout,)
    return spec


def make_weights(cfgs, seed=11):
    from egonet_b200 import synth
    return synth.lifter_weights(state_dict_spec(cfgs), seed)


def make_stats(cfgs, seed=12, image_size=(1242, 375)):
    from egonet_b200 import synth
    return synth.lifter_stats(cfgs, seed, image_size)


def _lin_bn_relu(sd, x, lin, bn):
    y = F.linear(x, sd[lin + '.weight'], sd[lin + '.bias'])
    y = F.batch_norm(y, sd[bn + '.running_mean'], sd[bn + '.running_var'],
                     sd[bn + '.weight'], sd[bn + '.bias'], False, 0.1, BN_EPS)
    return F.relu(y)


@torch.no_grad()
def lifter_forward(sd, cfgs, x):
    """FCModel.forward, eval mode: x [n, in] float32 tensor -> [n, out] float32."""
    y = _lin_bn_relu(sd, x, 'w1', 'batch_norm1')
    for i in range(cfgs['FCModel']['num_blocks']):
        p = 'res_blocks.%d' % i
        t = _lin_bn_relu(sd, y, p + '.w1', p + '.batch_norm1')
        t = _lin_bn_relu(sd, t, p + '.w2', p + '.batch_norm2')
        y = y + t
    return F.linear(y, sd['w2.weight'], sd['w2.bias'])


def normalize_1d(data, mean, std):
    return (data - mean) / std


def unnormalize_1d(data, mean, std):
    return data * std + mean


def lift_2d_to_3d(sd, cfgs, stats, kpts_2d):
    """kpts_2d: fp64 [n, 2J] screen key-points -> fp64 [n, J-1, 3] (egonet.py:473-485)."""
    data = normalize_1d(np.asarray(kpts_2d, dtype=np.float64), stats['mean_in'], stats['std_in'])
    data = torch.from_numpy(data.astype(np.float32))
    pred = lifter_forward(sd, cfgs, data).numpy()
    pred = unnormalize_1d(pred, stats['mean_out'], stats['std_out'])
    return pred.reshape(len(pred), -1, 3)
