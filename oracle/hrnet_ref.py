"""CPU restatement of the reference's HRNet pose network ``HC`` (TEST INFRASTRUCTURE).

Functional (state-dict driven) fp32 forward on torch-CPU that follows
``libs/model/heatmapModel/hrnet.py`` operation by operation:

* ``state_dict_spec``   -- the parameter/buffer inventory the reference module
  registers (names, shapes, order): ``hrnet.py:311-469`` (+ ``_make_layer``
  :512-529, ``_make_transition_layer`` :471-510, ``_make_stage`` :531-561,
  ``HighResolutionModule._make_branches/_make_fuse_layers`` :174-277).
* ``hrnet_forward``     -- ``PoseHighResolutionNet.forward`` ``hrnet.py:563-614``
  with ``BasicBlock.forward`` :76-92, ``Bottleneck.forward`` :113-133 and
  ``HighResolutionModule.forward`` :282-300.

The arithmetic is deliberately the same torch-CPU conv/batch-norm the
reference would execute on CPU, so this file doubles as the "port" CPU baseline
that ``bench.py`` times.  It is pinned against the real reference module by
``tests/golden/make_golden.py`` (run in the build container).
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, used everywhere in hrnet.py


# ----------------------------------------------------------------------------
# parameter inventory
# ----------------------------------------------------------------------------
def _bn(spec, prefix, c):
    spec[prefix + '.weight'] = (c,)
    spec[prefix + '.bias'] = (c,)
    spec[prefix + '.running_mean'] = (c,)
    spec[prefix + '.running_var'] = (c,)
    spec[prefix + '.num_batches_tracked'] = ()


def _conv(spec, key, cout, cin, kh, kw, bias=False):
    spec[key + '.weight'] = (cout, cin, kh, kw)
    if bias:
        spec[key + '.bias'] = (cout,)


def _basic_block(spec, p, cin, cout, downsample):
    # hrnet.py:66-74 registration order: conv1, bn1, conv2, bn2, downsample
    _conv(spec, p + '.conv1', cout, cin, 3, 3)
    _bn(spec, p + '.bn1', cout)
    _conv(spec, p + '.conv2', cout, cout, 3, 3)
    _bn(spec, p + '.bn2', cout)
    if downsample:
        _conv(spec, p + '.downsample.0', cout, cin, 1, 1)
        _bn(spec, p + '.downsample.1', cout)


def _stage_cfgs(cfgs):
    extra = cfgs['heatmapModel']['extra']
    out = []
    for name in ('stage2', 'stage3', 'stage4'):
        sc = extra[name]
        if sc['block'] != 'basic':
            raise NotImplementedError('oracle covers block=basic stages (all shipped configs)')
        out.append(sc)
    return out


def state_dict_spec(cfgs, in_channels=None):
    """Ordered {key: shape} of ``PoseHighResolutionNet(cfgs).state_dict()``."""
    hm = cfgs['heatmapModel']
    if in_channels is None:
        in_channels = 5 if hm.get('add_xy', False) else 3  # hrnet.py:688-689
    spec = OrderedDict()
    _conv(spec, 'conv1', 64, in_channels, 3, 3)
    _bn(spec, 'bn1', 64)
    _conv(spec, 'conv2', 64, 64, 3, 3)
    _bn(spec, 'bn2', 64)
    # layer1 = 4 Bottlenecks, planes 64, expansion 4 (hrnet.py:325)
    inpl = 64
    for k in range(4):
        p = 'layer1.%d' % k
        _conv(spec, p + '.conv1', 64, inpl, 1, 1)
        _bn(spec, p + '.bn1', 64)
        _conv(spec, p + '.conv2', 64, 64, 3, 3)
        _bn(spec, p + '.bn2', 64)
        _conv(spec, p + '.conv3', 256, 64, 1, 1)
        _bn(spec, p + '.bn3', 256)
        if k == 0:
            _conv(spec, p + '.downsample.0', 256, inpl, 1, 1)
            _bn(spec, p + '.downsample.1', 256)
        inpl = 256
    pre = [256]
    stages = _stage_cfgs(cfgs)
    for si, sc in enumerate(stages):
        cur = list(sc['num_channels'])
        # transition (hrnet.py:471-510)
        tp = 'transition%d' % (si + 1)
        for i in range(len(cur)):
            if i < len(pre):
                if cur[i] != pre[i]:
                    _conv(spec, '%s.%d.0' % (tp, i), cur[i], pre[i], 3, 3)
                    _bn(spec, '%s.%d.1' % (tp, i), cur[i])
            else:
                for j in range(i + 1 - len(pre)):
                    cin = pre[-1]
                    cout = cur[i] if j == i - len(pre) else cin
                    _conv(spec, '%s.%d.%d.0' % (tp, i, j), cout, cin, 3, 3)
                    _bn(spec, '%s.%d.%d.1' % (tp, i, j), cout)
        # stage (hrnet.py:531-561)
        nb = sc['num_branches']
        last_stage = si == len(stages) - 1
        for m in range(sc['num_modules']):
            mp = 'stage%d.%d' % (si + 2, m)
            multi = not (last_stage and m == sc['num_modules'] - 1)
            for b in range(nb):
                for k in range(sc['num_blocks'][b]):
                    _basic_block(spec, '%s.branches.%d.%d' % (mp, b, k), cur[b], cur[b], False)
            for i in range(nb if multi else 1):
                for j in range(nb):
                    fp = '%s.fuse_layers.%d.%d' % (mp, i, j)
                    if j > i:
                        _conv(spec, fp + '.0', cur[i], cur[j], 1, 1)
                        _bn(spec, fp + '.1', cur[i])
                    elif j < i:
                        for k in range(i - j):
                            cout = cur[i] if k == i - j - 1 else cur[j]
                            _conv(spec, '%s.%d.0' % (fp, k), cout, cur[j], 3, 3)
                            _bn(spec, '%s.%d.1' % (fp, k), cout)
        pre = cur
    nj = hm['num_joints']
    if hm['head_type'] == 'heatmap':
        k = hm['extra']['final_conv_kernel']
        _conv(spec, 'final_layer', nj, pre[0], k, k, bias=True)
    elif hm['head_type'] == 'coordinates':
        mw, mh = hm['heatmap_size']
        _conv(spec, 'head1.0', nj, pre[0], 1, 1, bias=True)
        cin = nj + 2
        for k in range(4):
            _basic_block(spec, 'head2.%d' % k, cin, 2 * nj, True)
            cin = 2 * nj
        _conv(spec, 'head2.4', 2 * nj, 2 * nj, int(mh / 16), int(mw / 16), bias=True)
    else:
        raise NotImplementedError(hm['head_type'])
    return spec


# ----------------------------------------------------------------------------
# deterministic synthetic weights (no checkpoints are available offline)
# ----------------------------------------------------------------------------
def make_weights(cfgs, seed=1, in_channels=None):
    """Seeded synthetic ``HC`` state dict over this module's own key inventory."""
    from egonet_b200 import synth
    return synth.hc_weights(state_dict_spec(cfgs, in_channels), seed)


def weights_digest(sd):
    from egonet_b200 import synth
    return synth.weights_digest(sd)


# ----------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------
class Exact:
    """Reference arithmetic: unfolded fp32 conv -> batch_norm, no storage rounding."""
    folded = False

    @staticmethod
    def q(t):
        return t

    @staticmethod
    def qw(w):
        return w


class Quantized:
    """Arithmetic model of the fp16 tensor-core path (documented in DESIGN.md):
    BN folded into the conv weights in fp32, folded weights rounded to
    ``dtype``, fp32 accumulation, activations rounded to ``dtype`` at every
    point where the CUDA path stores a tensor to HBM.  The stem conv1 and the
    head2 tail conv keep fp32 weights (CUDA-core kernels)."""
    folded = True

    def __init__(self, dtype=torch.float16):
        self.dtype = dtype

    def q(self, t):
        return t.to(self.dtype).to(torch.float32)

    def qw(self, w):
        return w.to(self.dtype).to(torch.float32)


def fold_bn(sd, conv, bn):
    """Fold eval-mode BN into (weight, bias): y = conv(x, w') + b'."""
    w = sd[conv + '.weight']
    b = sd.get(conv + '.bias')
    if bn is None:
        return w, (b if b is not None else torch.zeros(w.shape[0]))
    scale = sd[bn + '.weight'] / torch.sqrt(sd[bn + '.running_var'] + BN_EPS)
    shift = sd[bn + '.bias'] - sd[bn + '.running_mean'] * scale
    if b is not None:
        shift = shift + b * scale
    return w * scale.view(-1, 1, 1, 1), shift


def _cb(sd, x, conv, bn, stride=1, pad=0, relu=False, ctx=Exact, fp32_weights=False):
    """conv -> eval-mode BatchNorm -> optional ReLU (no storage rounding here)."""
    if ctx.folded:
        w, b = fold_bn(sd, conv, bn)
        y = F.conv2d(x, w if fp32_weights else ctx.qw(w), b, stride=stride, padding=pad)
    else:
        y = F.conv2d(x, sd[conv + '.weight'], sd.get(conv + '.bias'), stride=stride, padding=pad)
        if bn is not None:
            # ctx.training (oracle/train_ref.py): batch statistics + in-place running-stat update (momentum 0.1)
            y = F.batch_norm(y, sd[bn + '.running_mean'], sd[bn + '.running_var'],
                             sd[bn + '.weight'], sd[bn + '.bias'], getattr(ctx, 'training', False), 0.1, BN_EPS)
    return F.relu(y) if relu else y


def _basic(sd, x, p, stride=1, downsample=False, ctx=Exact):
    # hrnet.py:76-92
    out = ctx.q(_cb(sd, x, p + '.conv1', p + '.bn1', stride, 1, True, ctx))
    if downsample:
        res = ctx.q(_cb(sd, x, p + '.downsample.0', p + '.downsample.1', stride, 0, False, ctx))
    else:
        res = x
    out = _cb(sd, out, p + '.conv2', p + '.bn2', 1, 1, False, ctx)
    return ctx.q(F.relu(out + res))


def _bottleneck(sd, x, p, downsample, ctx=Exact):
    # hrnet.py:113-133
    out = ctx.q(_cb(sd, x, p + '.conv1', p + '.bn1', 1, 0, True, ctx))
    out = ctx.q(_cb(sd, out, p + '.conv2', p + '.bn2', 1, 1, True, ctx))
    if downsample:
        res = ctx.q(_cb(sd, x, p + '.downsample.0', p + '.downsample.1', 1, 0, False, ctx))
    else:
        res = x
    out = _cb(sd, out, p + '.conv3', p + '.bn3', 1, 0, False, ctx)
    return ctx.q(F.relu(out + res))


def _hr_module(sd, xs, mp, nblocks, multi, ctx=Exact):
    # hrnet.py:282-300
    nb = len(xs)
    xs = list(xs)
    for b in range(nb):
        for k in range(nblocks[b]):
            xs[b] = _basic(sd, xs[b], '%s.branches.%d.%d' % (mp, b, k), ctx=ctx)
    outs = []
    for i in range(nb if multi else 1):
        y = None
        for j in range(nb):
            fp = '%s.fuse_layers.%d.%d' % (mp, i, j)
            if j == i:
                t = xs[j]
            elif j > i:
                t = ctx.q(_cb(sd, xs[j], fp + '.0', fp + '.1', 1, 0, False, ctx))
                t = F.interpolate(t, scale_factor=2 ** (j - i), mode='nearest')
            else:
                t = xs[j]
                for k in range(i - j):
                    t = ctx.q(_cb(sd, t, '%s.%d.0' % (fp, k), '%s.%d.1' % (fp, k), 2, 1,
                                  k != i - j - 1, ctx))
            y = t if y is None else y + t
        outs.append(ctx.q(F.relu(y)))
    return outs


def coord_maps(map_width, map_height):
    """hrnet.py:461-467: channel 0 = x (varies along width), 1 = y, float32."""
    x_map = np.tile(np.linspace(0, 1, map_width), (map_height, 1)).reshape(1, 1, map_height, map_width)
    y_map = np.tile(np.linspace(0, 1, map_height).reshape(map_height, 1), (1, map_width))
    y_map = y_map.reshape(1, 1, map_height, map_width)
    return torch.from_numpy(np.concatenate([x_map, y_map], axis=1).astype(np.float32))


@torch.no_grad()
def hrnet_forward(sd, cfgs, x, taps=None, ctx=Exact):
    """``PoseHighResolutionNet.forward`` (hrnet.py:563-614), eval mode.

    x: [B, C, H, W] float32.  Returns ``heatmap`` for ``head_type='heatmap'`` or
    ``(heatmap, coords[B, J, 2])`` for ``'coordinates'``.  If ``taps`` is a dict
    it receives named intermediate activations (for per-stage parity tests).
    ``ctx=Exact`` is the reference's fp32 arithmetic; ``ctx=Quantized()`` models
    the storage roundings of the fp16 tensor-core path.
    """
    hm = cfgs['heatmapModel']
    x = ctx.q(_cb(sd, x, 'conv1', 'bn1', 2, 1, True, ctx, fp32_weights=True))
    if taps is not None:
        taps['stem1'] = x
    x = ctx.q(_cb(sd, x, 'conv2', 'bn2', 2, 1, True, ctx))
    if taps is not None:
        taps['stem2'] = x
    for k in range(4):
        x = _bottleneck(sd, x, 'layer1.%d' % k, k == 0, ctx)
    if taps is not None:
        taps['layer1'] = x
    stages = _stage_cfgs(cfgs)
    pre = [256]
    ys = [x]
    for si, sc in enumerate(stages):
        cur = list(sc['num_channels'])
        tp = 'transition%d' % (si + 1)
        xs = []
        for i in range(len(cur)):
            if i < len(pre):
                if cur[i] != pre[i]:
                    if si != 0:
                        # hrnet.py:583 would feed y_list[-1] (wrong width) here;
                        # no shipped config reaches this branch
                        raise NotImplementedError('channel-changing transition after stage 2')
                    xs.append(ctx.q(_cb(sd, ys[i], '%s.%d.0' % (tp, i), '%s.%d.1' % (tp, i), 1, 1, True, ctx)))
                else:
                    xs.append(ys[i])
            else:
                t = ys[-1]  # hrnet.py:575,583,591
                for j in range(i + 1 - len(pre)):
                    t = ctx.q(_cb(sd, t, '%s.%d.%d.0' % (tp, i, j), '%s.%d.%d.1' % (tp, i, j), 2, 1, True, ctx))
                xs.append(t)
        last_stage = si == len(stages) - 1
        for m in range(sc['num_modules']):
            multi = not (last_stage and m == sc['num_modules'] - 1)
            xs = _hr_module(sd, xs, 'stage%d.%d' % (si + 2, m), sc['num_blocks'], multi, ctx)
            if taps is not None:
                for b, t in enumerate(xs):
                    taps['stage%d.%d.out%d' % (si + 2, m, b)] = t
        ys = xs
        pre = cur
    feat = ys[0]
    if hm['head_type'] == 'heatmap':
        k = hm['extra']['final_conv_kernel']
        return F.conv2d(feat, ctx.qw(sd['final_layer.weight']), sd['final_layer.bias'],
                        padding=1 if k == 3 else 0)
    maps = F.conv2d(feat, ctx.qw(sd['head1.0.weight']), sd['head1.0.bias'])
    mw, mh = hm['heatmap_size']
    cm = coord_maps(mw, mh).repeat(len(maps), 1, 1, 1)
    t = ctx.q(torch.cat([maps, cm], dim=1))
    for k in range(4):
        t = _basic(sd, t, 'head2.%d' % k, stride=2, downsample=True, ctx=ctx)
        if taps is not None:
            taps['head2.%d' % k] = t
    t = F.conv2d(t, sd['head2.4.weight'], sd['head2.4.bias'])
    if taps is not None:
        taps['logits'] = t.reshape(len(t), -1)
    coords = torch.sigmoid(t).view(len(t), -1, 2)
    return maps, coords


def conv_inventory(cfgs, in_channels=3):
    """[(key, cout, cin, kh, kw, stride, out_h, out_w)] for MAC counting."""
    hm = cfgs['heatmapModel']
    W, H = hm['input_size']
    inv = []

    def add(key, cout, cin, k, stride, h, w):
        inv.append((key, cout, cin, k, k, stride, h, w))

    spec = state_dict_spec(cfgs, in_channels)
    # replay the forward symbolically
    h, w = H // 2, W // 2
    add('conv1', 64, in_channels, 3, 2, h, w)
    h, w = h // 2, w // 2
    add('conv2', 64, 64, 3, 2, h, w)
    inpl = 64
    for k in range(4):
        p = 'layer1.%d' % k
        add(p + '.conv1', 64, inpl, 1, 1, h, w)
        add(p + '.conv2', 64, 64, 3, 1, h, w)
        add(p + '.conv3', 256, 64, 1, 1, h, w)
        if k == 0:
            add(p + '.downsample.0', 256, inpl, 1, 1, h, w)
        inpl = 256
    for key, shape in spec.items():
        if not key.endswith('.weight') or len(shape) != 4:
            continue
        name = key[:-7]
        if name.startswith(('conv', 'layer1')):
            continue
        # resolution of the OUTPUT of this conv
        if name.startswith('transition'):
            parts = name.split('.')
            i = int(parts[1])
            if len(parts) == 3:  # same-resolution 3x3
                s, lvl = 1, i
            else:
                s, lvl = 2, i
            add(name, shape[0], shape[1], 3, s, h >> lvl, w >> lvl)
        elif name.startswith('stage'):
            parts = name.split('.')
            if parts[2] == 'branches':
                b = int(parts[3])
                add(name, shape[0], shape[1], 3, 1, h >> b, w >> b)
            else:
                i, j = int(parts[3]), int(parts[4])
                if j > i:
                    add(name, shape[0], shape[1], 1, 1, h >> j, w >> j)
                else:
                    k = int(parts[5])
                    add(name, shape[0], shape[1], 3, 2, h >> (j + k + 1), w >> (j + k + 1))
        elif name.startswith('head1') or name.startswith('final_layer'):
            add(name, shape[0], shape[1], shape[2], 1, h, w)
        elif name.startswith('head2'):
            parts = name.split('.')
            k = int(parts[1])
            if k == 4:
                inv.append((name, shape[0], shape[1], shape[2], shape[3], 1, 1, 1))
            elif parts[2] == 'conv1':
                add(name, shape[0], shape[1], 3, 2, h >> (k + 1), w >> (k + 1))
            elif parts[2] == 'conv2':
                add(name, shape[0], shape[1], 3, 1, h >> (k + 1), w >> (k + 1))
            else:
                add(name, shape[0], shape[1], 1, 2, h >> (k + 1), w >> (k + 1))
    return inv


def macs_per_crop(cfgs, in_channels=3):
    return sum(co * ci * kh * kw * oh * ow for _, co, ci, kh, kw, _, oh, ow in conv_inventory(cfgs, in_channels))
