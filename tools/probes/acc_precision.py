#!/usr/bin/env python
"""Accumulation error of the tcgen05 conv path on fp16-EXACT operands.

Inputs and weights are fp16-representable, so the only error of the fp32 result the kernel's epilogue sees
(TMEM accumulator + bias, read through the head1 fp32 side output) is how the tensor core accumulates the
K = taps x Cin products.  Compared against torch's fp64 conv on the same operands; the fp32 CUDA-core kernel
(impl 0, sequential fmaf) is printed next to it.  Decides whether an error-compensated hi/lo operand split can
reach the reference's fp32 accuracy on this hardware, or needs the K loop dealt over several accumulators.

    python tools/probes/acc_precision.py            (on the GPU box)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from egonet_b200 import _native as N  # noqa: E402

DEV = 'cuda'


def nhwc16(x):
    B, C, H, W = x.shape
    Cp = (C + 15) // 16 * 16
    out = torch.zeros((B, H, W, Cp), device=x.device, dtype=torch.float16)
    out[..., :C] = x.permute(0, 2, 3, 1).to(torch.float16)
    return out.contiguous()


def one(Cin, Cout, H, W, k, B, positive):
    g = torch.Generator().manual_seed(Cin + 7 * Cout + k)
    x = torch.randn((B, Cin, H, W), generator=g)
    if positive:
        x = x.clamp_min(0)               # post-ReLU activations
    x = x.to(torch.float16)
    w = (torch.randn((Cout, Cin, k, k), generator=g) / (Cin * k * k) ** 0.5).to(torch.float16)
    bias = torch.zeros((Cout,))
    w32 = w.float().contiguous()                 # keep the host buffer alive across the calls
    xin = nhwc16(x.float().to(DEV))
    ref = torch.nn.functional.conv2d(x.double().to(DEV), w.double().to(DEV), None, padding=1 if k == 3 else 0)
    row = {'shape': '%d->%d @%dx%d k%d B%d %s' % (Cin, Cout, H, W, k, B, 'relu-in' if positive else 'signed-in'),
           'K': Cin * k * k, 'ref_rms': float(ref.pow(2).mean().sqrt())}
    for impl, name in ((1, 'tc'), (0, 'simt')):
        out = torch.empty((B, H, W, (Cout + 15) // 16 * 16), device=DEV, dtype=torch.float16)
        acc = torch.full((B, Cout, H, W), float('nan'), device=DEV, dtype=torch.float32)
        N.check(N.lib().egn_debug_conv_acc(impl, N.ptr(xin), N.ptr(w32), N.ptr(bias), N.ptr(out),
                                           N.ptr(acc), B, H, W, Cin, Cout, k, 1, N.current_stream()))
        torch.cuda.synchronize()
        err = acc.double() - ref
        # signed error projected on the reference (a systematic shrink of magnitudes shows up as a negative slope)
        slope = float((err * ref).sum() / (ref * ref).sum())
        row[name] = {'max_abs': float(err.abs().max()), 'rms': float(err.pow(2).mean().sqrt()),
                     'rms_rel_to_ref_rms': float(err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()),
                     'slope': slope}
    return row


def main():
    rows = []
    for positive in (False, True):
        for (Cin, Cout, H, W, k, B) in ((48, 48, 64, 64, 3, 2), (96, 96, 32, 32, 3, 2), (192, 192, 16, 16, 3, 4),
                                        (384, 384, 8, 8, 3, 8), (256, 64, 64, 64, 1, 1)):
            rows.append(one(Cin, Cout, H, W, k, B, positive))
            print(json.dumps(rows[-1]))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    json.dump(rows, open(os.path.join(out, 'acc_precision.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
