"""B200-native drop-in for the reference's ``libs/model/FCmodel.py`` (the lifter ``L``).

``FCModel`` keeps the reference's constructor arguments and ``state_dict`` layout
[FCmodel.py:45-90] (``w1``, ``batch_norm1``, ``res_blocks.<i>.{w1,batch_norm1,w2,
batch_norm2}``, ``w2``) so ``L.pth`` loads strictly.  ``forward`` [FCmodel.py:92-105]
runs the fused native chain (``egn_lifter_forward``) on a CUDA tensor; the
input/output statistics of ``EgoNet.lift_2d_to_3d`` are fused into that chain by
``EgoNet`` through ``lift`` below.  Inference only (eval-mode BatchNorm, Dropout
is the identity).
"""
import ctypes

import torch
import torch.nn as nn

from ... import _native as N
from ._params import attach, default_init, host_state


class FCModel(nn.Module):
    def __init__(self, stage_id=1, num_neurons=1024, num_blocks=2, p_dropout=0.5, norm_twoD=False,
                 kaiming=False, refine_3d=False, leaky=False, dm=False, input_size=32, output_size=64):
        super().__init__()
        if leaky:
            raise NotImplementedError('leaky ReLU lifter is not supported by the native engine '
                                      '(leaky: False in every shipped config)')
        self.num_neurons, self.p_dropout, self.num_blocks = num_neurons, p_dropout, num_blocks
        self.stage_id, self.refine_3d, self.leaky, self.dm = stage_id, refine_3d, leaky, dm
        self.input_size, self.output_size = input_size, output_size
        n = num_neurons

        def bn(prefix):
            for leaf in ('weight', 'bias', 'running_mean', 'running_var'):
                attach(self, prefix + '.' + leaf, (n,))
            attach(self, prefix + '.num_batches_tracked', ())

        attach(self, 'w1.weight', (n, input_size))
        attach(self, 'w1.bias', (n,))
        bn('batch_norm1')
        for i in range(num_blocks):
            p = 'res_blocks.%d' % i
            attach(self, p + '.w1.weight', (n, n))
            attach(self, p + '.w1.bias', (n,))
            bn(p + '.batch_norm1')
            attach(self, p + '.w2.weight', (n, n))
            attach(self, p + '.w2.bias', (n,))
            bn(p + '.batch_norm2')
        attach(self, 'w2.weight', (output_size, n))
        attach(self, 'w2.bias', (output_size,))
        default_init(self)
        if kaiming:
            nn.init.kaiming_normal_(self.w1.weight.data)
            nn.init.kaiming_normal_(self.w2.weight.data)
        self._handle = ctypes.c_void_p()
        N.check(N.lib().egn_lifter_create(input_size, output_size, n, num_blocks, ctypes.byref(self._handle)))
        self._dirty = True
        self._stats = None
        self._workspace = None

    def __del__(self):
        try:
            if self._handle:
                N.lib().egn_lifter_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def _apply(self, fn, *a, **k):
        self._dirty = True
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._dirty = True
        return super().load_state_dict(*a, **k)

    def refresh(self):
        self._dirty = True

    def set_stats(self, stats):
        """``LS`` dict of ``EgoNet`` (mean_in/std_in/mean_out/std_out); None = identity."""
        self._stats = stats
        self._dirty = True

    def _sync(self):
        import numpy as np
        L = N.lib()
        for key, t in host_state(self):
            shape = (ctypes.c_int64 * max(1, t.dim()))(*t.shape)
            N.check(L.egn_lifter_set_weight(self._handle, key.encode(), N.ptr(t), shape, t.dim()))
        if self._stats is None:
            st = [np.zeros(self.input_size), np.ones(self.input_size),
                  np.zeros(self.output_size), np.ones(self.output_size)]
        else:
            st = [np.ascontiguousarray(np.asarray(self._stats[k], dtype=np.float64).reshape(-1))
                  for k in ('mean_in', 'std_in', 'mean_out', 'std_out')]
            if st[0].size != self.input_size or st[2].size != self.output_size:
                raise ValueError('LS statistics do not match the lifter input/output size')
        N.check(L.egn_lifter_set_stats(self._handle, *[s.ctypes.data_as(ctypes.c_void_p) for s in st]))
        N.check(L.egn_lifter_finalize(self._handle))
        self._dirty = False

    def lift(self, kpts_2d, want_raw=False):
        """fp64 CUDA [n, input_size] screen key-points -> fp64 CUDA [n, output_size]
        with the LS (de)normalisation fused in (``EgoNet.lift_2d_to_3d`` egonet.py:473-485)."""
        if self.training:
            raise NotImplementedError('the native lifter is inference-only: call .eval() first')
        if not kpts_2d.is_cuda:
            raise RuntimeError('the native lifter has no CPU path: input must be a CUDA tensor')
        L = N.lib()
        with torch.cuda.device(kpts_2d.device):
            # the engine's weights live on the device that was current at the last sync: a handle is bound to
            # one device (and one stream at a time); an input on another GPU re-uploads them there
            if self._dirty or getattr(self, '_weights_device', None) != kpts_2d.device.index:
                self._sync()
                self._weights_device = kpts_2d.device.index
            x = kpts_2d.detach().to(torch.float64).contiguous()
            n = x.shape[0]
            out = torch.empty((n, self.output_size), device=x.device, dtype=torch.float64)
            raw = torch.empty((n, self.output_size), device=x.device, dtype=torch.float32) if want_raw else None
            need = L.egn_lifter_workspace_bytes(self._handle, max(n, 1))
            if self._workspace is None or self._workspace.numel() < need or self._workspace.device != x.device:
                self._workspace = torch.empty(need, device=x.device, dtype=torch.uint8)
            N.check(L.egn_lifter_forward(self._handle, N.ptr(x), n, N.ptr(out), N.ptr(raw),
                                         N.ptr(self._workspace), self._workspace.numel(), N.current_stream()))
        return (out, raw) if want_raw else out

    def forward(self, x):
        """[FCmodel.py:92-95] already-normalised fp32 [n, input_size] -> fp32 [n, output_size].

        The native chain normalises on load, so the input is mapped back through
        the fp64 statistics first; (x * std + mean - mean) / std evaluated in fp64
        rounds back to the original fp32 value."""
        x64 = x.detach().to(torch.float64)
        if self._stats is not None:
            std = torch.as_tensor(self._stats['std_in'], dtype=torch.float64, device=x.device).reshape(1, -1)
            mean = torch.as_tensor(self._stats['mean_in'], dtype=torch.float64, device=x.device).reshape(1, -1)
            x64 = x64 * std + mean
        _, raw = self.lift(x64, want_raw=True)
        return raw

    def get_representation(self, x):
        raise NotImplementedError('get_representation (FCmodel.py:97-105) is a training-time helper; '
                                  'the fused native lifter does not expose the hidden state')


def get_fc_model(stage_id, cfgs, input_size, output_size, architecture_type='FCModel'):
    """[FCmodel.py:107-121]"""
    c = cfgs[architecture_type]
    return FCModel(stage_id=stage_id, refine_3d=c['refine_3d'], norm_twoD=c['norm_twoD'],
                   num_blocks=c['num_blocks'], input_size=input_size, output_size=output_size,
                   num_neurons=c['num_neurons'], p_dropout=c['dropout'], leaky=c['leaky'])


def get_cascade():
    return nn.ModuleList([])
