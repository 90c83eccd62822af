"""Run on the GPU box: device time of the lifter chain (2 * blocks + 2 GEMM launches) on 256 instances."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from oracle import configs, lifter_ref
from egonet_b200.libs.model.FCmodel import get_fc_model

cfgs = configs.demo_cfgs()
fc = cfgs['FCModel']
L = get_fc_model(1, cfgs, fc['input_size'], fc['output_size']).eval()
L.load_state_dict(lifter_ref.make_weights(cfgs, 11))
L.set_stats(lifter_ref.make_stats(cfgs, 12))
L = L.cuda()
for n in (256, 64):
    x = torch.rand((n, fc['input_size']), dtype=torch.float64, device='cuda') * 300
    for _ in range(3):
        L.lift(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        L.lift(x)
    e1.record()
    torch.cuda.synchronize()
    print('lifter n=%d: %.1f us per call' % (n, e0.elapsed_time(e1) * 1e3 / 20))
