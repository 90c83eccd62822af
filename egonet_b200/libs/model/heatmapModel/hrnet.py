"""B200-native drop-in for the reference's ``libs/model/heatmapModel/hrnet.py``.

Same public surface (upstream file:line in brackets):

* ``get_pose_net(cfgs, is_train)``                  [hrnet.py:675-690]
* ``PoseHighResolutionNet(cfgs)`` with identical ``state_dict()`` keys / shapes /
  order, ``forward``, ``init_weights``, ``modify_input_channel``,
  ``load_my_state_dict``                            [hrnet.py:309-667]

The module holds the parameters exactly as the reference does (so ``HC.pth``
loads strictly) but ``forward`` does not run torch ops: it hands the input
pointer to the native engine (``egn_hrnet_forward``), which replays the whole
network as fused sm_100a kernels.  BatchNorm folding / repacking happens lazily
before the first forward after the parameters changed.

``forward`` in eval mode is the inference engine; in training mode it runs the
train-mode forward of the native training engine (see ``libs/trainer``).
"""
import ctypes
import logging
import os

import torch
import torch.nn as nn

from .... import _native as N
from .._params import ParamTree, attach, default_init, host_state, is_norm

logger = logging.getLogger(__name__)


def _engine_cfg(cfgs, in_channels, precision, conv_impl, keep_taps):
    hm = cfgs['heatmapModel']
    extra = hm['extra']
    c = N.HRNetCfg()
    c.in_channels = in_channels
    c.input_w, c.input_h = int(hm['input_size'][0]), int(hm['input_size'][1])
    c.heatmap_w, c.heatmap_h = int(hm['heatmap_size'][0]), int(hm['heatmap_size'][1])
    c.num_joints = int(hm['num_joints'])
    head = hm['head_type']
    if head == 'heatmap':
        c.head_type = N.HEAD_HEATMAP
    elif head == 'coordinates':
        c.head_type = N.HEAD_COORDINATES
    else:
        # same exception type the reference raises for unknown heads (hrnet.py:469)
        raise NotImplementedError('head_type %r is not supported by the native engine' % head)
    if hm.get('pixel_shuffle', False):
        raise NotImplementedError('pixel_shuffle up-sampling is not supported by the native engine')
    c.final_conv_kernel = int(extra.get('final_conv_kernel', 1))
    c.num_stages = 3
    for s, name in enumerate(('stage2', 'stage3', 'stage4')):
        sc = extra[name]
        if sc['block'] != 'basic':
            raise NotImplementedError('only block=basic stages are supported (all shipped configs)')
        if sc.get('fuse_method', 'sum') != 'sum':
            raise NotImplementedError('only fuse_method=sum is supported')
        c.stage_modules[s] = int(sc['num_modules'])
        c.stage_branches[s] = int(sc['num_branches'])
        for b in range(int(sc['num_branches'])):
            c.stage_blocks[s][b] = int(sc['num_blocks'][b])
            c.stage_channels[s][b] = int(sc['num_channels'][b])
    c.precision = precision
    c.conv_impl = conv_impl
    c.keep_taps = 1 if keep_taps else 0
    return c


class _TrainFunction(torch.autograd.Function):
    """Train-mode forward / backward of HC through the native training engine (``egn_hrnet_forward_train`` /
    ``egn_hrnet_backward``): autograd sees one node whose inputs are the module's parameters, so the reference's
    ``loss.backward(); optim.step()`` loop [trainer.py:183-198] works unchanged."""

    @staticmethod
    def forward(ctx, module, x, *params):
        maps, coords = module._train_forward(x)
        ctx.module = module
        ctx.batch = x.shape[0]
        ctx.needs = [p.requires_grad for p in params]
        if coords is None:
            return maps
        return maps, coords

    @staticmethod
    def backward(ctx, grad_maps, grad_coords=None):
        grads = ctx.module._train_backward(grad_maps, grad_coords, ctx.batch)
        return (None, None) + tuple(g if need else None for g, need in zip(grads, ctx.needs))


_PRECISIONS = {'fp32': N.PREC_FP32, 'fp16': N.PREC_FP16, 'fp16x2': N.PREC_FP16X2}


class PoseHighResolutionNet(nn.Module):
    """HRNet pose network ``HC`` executed by the native sm_100a engine.

    Extra (optional) config keys, all under ``cfgs['heatmapModel']`` and all with
    defaults so reference YAML files work unchanged:
      ``b200_precision``: ``'fp16x2'`` (default; tcgen05 tensor cores with
      error-compensated split-fp16 operands and fp32 accumulation -- meets the 1e-4
      parity bound against the reference's fp32 results), ``'fp16'`` (opt-in fast mode:
      single fp16 operands, ~1e-3 on coordinates) or ``'fp32'`` (fp32 storage and
      CUDA-core convolutions, the exact comparator).
      ``b200_conv_impl``: ``'auto'`` | ``'simt'``; ``b200_keep_taps``: bool.
    """

    def __init__(self, cfgs, **kwargs):
        super().__init__()
        hm = cfgs['heatmapModel']
        self.cfgs = cfgs
        self.num_joints = hm['num_joints']
        self.head_type = hm['head_type']
        self.pixel_shuffle = hm.get('pixel_shuffle', False)
        self.pretrained_layers = hm['extra'].get('pretrained_layers', ['*'])
        self._precision = _PRECISIONS[kwargs.get('precision', hm.get('b200_precision', 'fp16x2'))]
        impl = kwargs.get('conv_impl', hm.get('b200_conv_impl', 'auto'))
        if os.environ.get('EGN_CONV_IMPL'):        # debugging aid: force 'simt' for a whole process
            impl = os.environ['EGN_CONV_IMPL']
        self._conv_impl = {'auto': N.CONV_AUTO, 'simt': N.CONV_SIMT}[impl]
        self._keep_taps = bool(kwargs.get('keep_taps', hm.get('b200_keep_taps', False)))
        self._in_channels = 3
        self._handle = None
        self._dirty = True
        self._workspace = None
        self._train = None          # native training engine state (created by the first train-mode forward)
        self._build()

    # -- construction -----------------------------------------------------
    def _build(self):
        """(Re)create the engine handle and the parameter tree it describes."""
        self._destroy()
        L = N.lib()
        cfg = _engine_cfg(self.cfgs, self._in_channels, self._precision, self._conv_impl, self._keep_taps)
        handle = ctypes.c_void_p()
        N.check(L.egn_hrnet_create(ctypes.byref(cfg), ctypes.byref(handle)))
        self._handle = handle
        old = {k: v for k, v in self.state_dict().items()} if len(self._modules) else {}
        for name in list(self._modules):
            del self._modules[name]
        shape = (ctypes.c_int64 * 4)()
        for i in range(L.egn_hrnet_num_weights(handle)):
            key = L.egn_hrnet_weight_key(handle, i).decode()
            nd = L.egn_hrnet_weight_shape(handle, i, shape)
            attach(self, key, tuple(shape[d] for d in range(nd)))
        default_init(self)
        if old:  # keep whatever still fits (used by modify_input_channel)
            with torch.no_grad():
                for k, v in self.state_dict().items():
                    if k in old and old[k].shape == v.shape:
                        v.copy_(old[k])
        self._dirty = True

    def _destroy(self):
        if self._handle is not None:
            N.lib().egn_hrnet_destroy(self._handle)
            self._handle = None
        tr = getattr(self, '_train', None)
        if tr is not None:
            N.lib().egn_hrnet_train_destroy(tr['handle'])
            self._train = None

    # -- training ------------------------------------
    def _train_state(self, device):
        """Create the training engine and move every parameter / BatchNorm statistic into ONE flat fp32 device
        buffer (state_dict order, ``egn_hrnet_train_param_offset``); the module's Parameters become views of it,
        so optimisers, ``state_dict()`` and checkpoints keep working while the engine reads and updates them in
        place."""
        L = N.lib()
        tr = self._train
        if tr is None:
            cfg = _engine_cfg(self.cfgs, self._in_channels, N.PREC_FP32, N.CONV_SIMT, False)
            handle = ctypes.c_void_p()
            N.check(L.egn_hrnet_train_create(ctypes.byref(cfg), ctypes.byref(handle)))
            tr = self._train = {'handle': handle, 'flat': None, 'workspace': None, 'entries': None}
        sd = self.state_dict(keep_vars=True)
        if tr['entries'] is None:
            entries = []
            for i, (key, t) in enumerate(sd.items()):
                off = L.egn_hrnet_train_param_offset(tr['handle'], i)
                if off >= 0:
                    entries.append((key, off, t.numel(), tuple(t.shape)))
            tr['entries'] = entries
        flat = tr['flat']
        stale = flat is None or flat.device != device
        if not stale:
            base = flat.data_ptr()
            stale = any(sd[k].data_ptr() != base + 4 * off for k, off, _, _ in tr['entries'])
        if stale:
            flat = torch.zeros(L.egn_hrnet_train_flat_size(tr['handle']), device=device, dtype=torch.float32)
            with torch.no_grad():
                for k, off, n, shape in tr['entries']:
                    view = flat[off:off + n].view(shape)
                    view.copy_(sd[k].detach().to(device=device, dtype=torch.float32))
                    sd[k].data = view
            tr['flat'] = flat
        return tr

    def _train_forward(self, x):
        L = N.lib()
        hm = self.cfgs['heatmapModel']
        with torch.cuda.device(x.device):
            tr = self._train_state(x.device)
            x = x.detach().float().contiguous()
            B = x.shape[0]
            maps = torch.empty((B, self.num_joints, hm['heatmap_size'][1], hm['heatmap_size'][0]), device=x.device,
                               dtype=torch.float32)
            coords = torch.empty((B, self.num_joints, 2), device=x.device, dtype=torch.float32) \
                if self.head_type == 'coordinates' else None
            need = L.egn_hrnet_train_workspace_bytes(tr['handle'], B)
            ws = tr['workspace']
            if ws is None or ws.numel() < need or ws.device != x.device:
                tr['workspace'] = None                      # release before the larger allocation
                ws = tr['workspace'] = torch.empty(need, device=x.device, dtype=torch.uint8)
            N.check(L.egn_hrnet_forward_train(tr['handle'], N.ptr(tr['flat']), N.ptr(x), B, N.ptr(maps), N.ptr(coords),
                                              0.1, 1, N.ptr(ws), ws.numel(), N.current_stream()))
            with torch.no_grad():                           # nn.BatchNorm2d bookkeeping
                torch._foreach_add_([b for n, b in self.named_buffers() if n.endswith('num_batches_tracked')], 1)
        self._dirty = True                                  # the folded inference weights are stale now
        return maps, coords

    def _train_backward(self, grad_maps, grad_coords, batch):
        L = N.lib()
        tr = self._train
        with torch.cuda.device(tr['flat'].device):
            g = grad_maps.detach().float().contiguous() if grad_maps is not None else None
            gc = grad_coords.detach().float().contiguous() if grad_coords is not None else None
            flat_grads = torch.empty_like(tr['flat'])
            ws = tr['workspace']
            N.check(L.egn_hrnet_backward(tr['handle'], N.ptr(tr['flat']), N.ptr(g), N.ptr(gc), batch, N.ptr(flat_grads),
                                         N.ptr(ws), ws.numel(), N.current_stream()))
        tr['last_grads'] = flat_grads                        # FlatOptimizer consumes it without a gather
        views = {k: flat_grads[off:off + n].view(shape) for k, off, n, shape in tr['entries']}
        return [views[name] for name, _ in self.named_parameters()]

    def train_flops_per_sample(self):
        return N.lib().egn_hrnet_train_flops_per_sample(self._train_state(torch.device('cuda', torch.cuda.current_device()))['handle'])

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    # -- nn.Module plumbing: any path that can change parameters marks the engine stale
    def _apply(self, fn, *a, **k):
        self._dirty = True
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._dirty = True
        return super().load_state_dict(*a, **k)

    def refresh(self):
        """Call after editing parameters in place (e.g. ``p.data.mul_``)."""
        self._dirty = True

    def _sync_weights(self):
        L = N.lib()
        for key, t in host_state(self):
            shape = (ctypes.c_int64 * max(1, t.dim()))(*t.shape)
            N.check(L.egn_hrnet_set_weight(self._handle, key.encode(), N.ptr(t), shape, t.dim()))
        N.check(L.egn_hrnet_finalize(self._handle))
        self._dirty = False

    # -- reference API ----------------------------------------------------
    def forward(self, x):
        if self.training:
            if not x.is_cuda:
                raise RuntimeError('the native HC engine has no CPU path: input must be a CUDA tensor')
            return _TrainFunction.apply(self, x, *[p for _, p in self.named_parameters()])
        if not x.is_cuda:
            raise RuntimeError('the native HC engine has no CPU path: input must be a CUDA tensor')
        hm = self.cfgs['heatmapModel']
        if x.dim() != 4 or x.shape[1] != self._in_channels or x.shape[2] != hm['input_size'][1] \
                or x.shape[3] != hm['input_size'][0]:
            raise ValueError('expected input [B,%d,%d,%d], got %s' % (
                self._in_channels, hm['input_size'][1], hm['input_size'][0], tuple(x.shape)))
        with torch.cuda.device(x.device):
            # the engine's folded weights and tensor maps live on the device that was current at the last sync: a
            # handle is bound to one device (and one stream at a time); an input on another GPU re-uploads them
            if self._dirty or getattr(self, '_weights_device', None) != x.device.index:
                self._sync_weights()
                self._weights_device = x.device.index
            x = x.detach().float().contiguous()
            B = x.shape[0]
            maps, coords, _ = self.run(x)
        if self.head_type == 'heatmap':
            return maps
        return maps, coords

    def run(self, x, want_heatmap=True, want_logits=False):
        """Engine call on a contiguous fp32 CUDA tensor; returns (maps, coords, logits)."""
        L = N.lib()
        B = x.shape[0]
        hm = self.cfgs['heatmapModel']
        J = self.num_joints
        maps = torch.empty((B, J, hm['heatmap_size'][1], hm['heatmap_size'][0]), device=x.device,
                           dtype=torch.float32) if want_heatmap else None
        coord_head = self.head_type == 'coordinates'
        coords = torch.empty((B, J, 2), device=x.device, dtype=torch.float32) if coord_head else None
        logits = torch.empty((B, 2 * J), device=x.device, dtype=torch.float32) if (coord_head and want_logits) else None
        need = L.egn_hrnet_workspace_bytes(self._handle, max(B, 1))
        if self._workspace is None or self._workspace.numel() < need or self._workspace.device != x.device:
            self._workspace = torch.empty(need, device=x.device, dtype=torch.uint8)
        N.check(L.egn_hrnet_forward(self._handle, N.ptr(x), B, N.ptr(maps), N.ptr(coords), N.ptr(logits),
                                    N.ptr(self._workspace), self._workspace.numel(), N.current_stream()))
        return maps, coords, logits

    def read_tap(self, name, batch):
        """Intermediate activation of the last forward as fp32 NCHW (needs keep_taps)."""
        L = N.lib()
        dims = (ctypes.c_int * 3)()
        # first call with a scratch buffer large enough for any tap
        hm = self.cfgs['heatmapModel']
        cap = batch * 256 * (hm['input_size'][0] // 2) * (hm['input_size'][1] // 2)
        out = torch.empty(cap, device=self._workspace.device, dtype=torch.float32)
        N.check(L.egn_hrnet_read_tap(self._handle, name.encode(), batch, N.ptr(self._workspace), N.ptr(out),
                                     dims, N.current_stream()))
        C, H, W = dims[0], dims[1], dims[2]
        return out[:batch * C * H * W].view(batch, C, H, W).clone()

    def stats(self):
        L = N.lib()
        return {'macs_per_crop': L.egn_hrnet_macs_per_crop(self._handle),
                'launches': L.egn_hrnet_num_launches(self._handle),
                'tc_launches': L.egn_hrnet_num_tc_launches(self._handle),
                'act_bytes_per_crop': L.egn_hrnet_act_bytes_per_crop(self._handle),
                'weight_bytes': L.egn_hrnet_weight_bytes(self._handle)}

    def init_weights(self, pretrained=''):
        """[hrnet.py:616-647] N(0, 1e-3) conv weights, zero biases, unit BatchNorm,
        then an optional partial load of ``pretrained`` filtered by ``pretrained_layers``."""
        logger.info('=> init weights from normal distribution')
        with torch.no_grad():
            for node in self.modules():
                w = node._parameters.get('weight')
                if w is None:
                    continue
                b = node._parameters.get('bias')
                if is_norm(node):
                    w.fill_(1.0)
                    b.zero_()
                elif w.dim() == 4:
                    w.normal_(std=0.001)
                    if b is not None:
                        b.zero_()
        if os.path.isfile(pretrained):
            sd = torch.load(pretrained, map_location='cpu')
            logger.info('=> loading pretrained model {}'.format(pretrained))
            keep = {k: v for k, v in sd.items()
                    if k.split('.')[0] in self.pretrained_layers or self.pretrained_layers[0] == '*'}
            self.load_state_dict(keep, strict=False)
            logger.info('{:d} modules initialized.'.format(len(keep)))
        elif pretrained:
            logger.error('=> please download pre-trained models first!')
            raise ValueError('{} does not exist!'.format(pretrained))
        self._dirty = True

    def modify_input_channel(self, num_channels):
        """[hrnet.py:649-659] widen conv1 to ``num_channels`` inputs, keeping the RGB filters."""
        if num_channels == self._in_channels:
            return
        old = self.conv1.weight.detach().clone()
        self._in_channels = num_channels
        self._build()
        with torch.no_grad():
            n = min(old.shape[1], num_channels)
            self.conv1.weight[:, :n] = old[:, :n]
        self._dirty = True

    def load_my_state_dict(self, state_dict):
        """[hrnet.py:661-667] copy every entry whose name exists here."""
        own = self.state_dict()
        for name, param in state_dict.items():
            if name in own:
                own[name].copy_(param.data)
        self._dirty = True


def is_freezed(name, freeze_names):
    return any(name.startswith(prefix) for prefix in freeze_names)


def get_pose_net(cfgs, is_train, **kwargs):
    """[hrnet.py:675-690]"""
    model = PoseHighResolutionNet(cfgs, **kwargs)
    if is_train and cfgs['heatmapModel']['init_weights']:
        model.init_weights(cfgs['heatmapModel'].get('pretrained', ''))
    for name, param in model.named_parameters():
        if is_freezed(name, cfgs['heatmapModel']['extra'].get('freeze_layers', [])):
            param.requires_grad = False
            print('{:s} freezed during training.'.format(name))
    if cfgs['heatmapModel'].get('add_xy', False):
        model.modify_input_channel(5)
    return model
