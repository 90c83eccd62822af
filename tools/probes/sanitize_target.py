#!/usr/bin/env python
"""Workload for compute-sanitizer (memcheck / racecheck): every tcgen05 conv kernel variant (persistent v3 with
and without CTA pairs, window-run v2, per-tap v1; fp16 and fp16x2 storage) on small layers, one tiny HC forward per
precision mode, the decode / affine / lifter / pose kernels, the crop front-end and one tiny training step.

    compute-sanitizer --tool memcheck  python tools/probes/sanitize_target.py
    compute-sanitizer --tool racecheck python tools/probes/sanitize_target.py convs
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from egonet_b200 import _native as N  # noqa: E402
import test_gpu_parity as T  # noqa: E402
from oracle import configs, egonet_ref, hrnet_ref  # noqa: E402

CASES = [(48, 48, 32, 32, 3, 1, 2), (96, 96, 32, 32, 3, 1, 3), (64, 64, 16, 16, 3, 2, 2), (64, 256, 32, 32, 1, 1, 1),
         (192, 192, 16, 16, 3, 1, 1), (35, 66, 32, 32, 3, 2, 1), (96, 96, 24, 20, 3, 1, 3)]
ENV_KEYS = ('EGN_TC_PAIR', 'EGN_TC_V3', 'EGN_TC_V2', 'EGN_TC_V2_SPLIT', 'EGN_TC_V4', 'EGN_TC_BLK', 'EGN_TC_V4_PERSIST',
            'EGN_TC_V1_STAGED', 'EGN_TC_V4_STAGED', 'EGN_TC_V4_PAIR', 'EGN_TC_V4_FOLD')
VARIANTS = {'auto': {}, 'no_pair': {'EGN_TC_PAIR': '0'}, 'no_blk': {'EGN_TC_BLK': '0'},
            'no_v3': {'EGN_TC_V3': '0', 'EGN_TC_V2_SPLIT': '1'},
            'v4': {'EGN_TC_V3': '0', 'EGN_TC_V2': '0', 'EGN_TC_V4': '2'},                      # tap-window kernel, persistent where it fits
            'v4_one_tile': {'EGN_TC_V3': '0', 'EGN_TC_V2': '0', 'EGN_TC_V4': '2', 'EGN_TC_V4_PERSIST': '0'},
            'v4_nopair': {'EGN_TC_V3': '0', 'EGN_TC_V2': '0', 'EGN_TC_V4': '2', 'EGN_TC_V4_PAIR': '0'},
            'v4_nofold': {'EGN_TC_V3': '0', 'EGN_TC_V2': '0', 'EGN_TC_V4': '2', 'EGN_TC_V4_FOLD': '0'},
            'v1_only': {'EGN_TC_V3': '0', 'EGN_TC_V2': '0', 'EGN_TC_V4': '0'},                # per-tap kernel, staged epilogue
            'v1_direct': {'EGN_TC_V3': '0', 'EGN_TC_V2': '0', 'EGN_TC_V4': '0', 'EGN_TC_V1_STAGED': '0'}}


def convs():
    for vname, env in VARIANTS.items():
        for k in ENV_KEYS:
            os.environ.pop(k, None)
        os.environ.update(env)
        for dtype, pack in ((1, T._nhwc16), (2, T._split16)):
            for (Cin, Cout, H, W, k, stride, B) in CASES:
                g = torch.Generator().manual_seed(Cin + Cout)
                x = torch.randn((B, Cin, H, W), generator=g).cuda()
                w = (torch.randn((Cout, Cin, k, k), generator=g) / (Cin * k * k) ** 0.5).contiguous()
                bias = torch.randn((Cout,), generator=g).contiguous()
                pad = 1 if k == 3 else 0
                OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
                xin = pack(x)
                res = pack(torch.randn((B, Cout, OH, OW), generator=g).cuda())
                out = torch.zeros_like(res)
                N.check(N.lib().egn_conv2d_fused(1, dtype, N.ptr(xin), N.ptr(w), N.ptr(bias), N.ptr(res), N.ptr(out),
                                                 B, H, W, Cin, Cout, k, stride, 1, N.current_stream()))
                torch.cuda.synchronize()
                assert torch.isfinite(out).all()
        print('convs ok:', vname, flush=True)
    for k in ENV_KEYS:
        os.environ.pop(k, None)


def network():
    cfgs = configs.tiny_cfgs()
    for prec in ('fp16x2', 'fp16', 'fp32'):
        ego = T._egonet(cfgs, prec)
        n = 3
        crops = egonet_ref.synth_crops(n, cfgs, 3).cuda()
        recs = egonet_ref.synth_boxes(n, cfgs, 4)
        out = ego.forward_crops(crops, np.array([r['center'] for r in recs]), np.array([r['scale'] for r in recs]),
                                K=egonet_ref.KITTI_K, alpha_mode='proj', return_all=True)
        maps, _ = ego.HC(crops)
        from egonet_b200.libs.common import img_proc
        img_proc.get_max_preds(maps)
        img_proc.soft_arg_max(maps)
        torch.cuda.synchronize()
        assert torch.isfinite(out['pose']).all()
        print('network ok:', prec, flush=True)


def train():
    from egonet_b200.libs.loss.function import JointsMSELoss
    from egonet_b200.libs.model.heatmapModel.hrnet import get_pose_net
    from egonet_b200.libs.optimizer.optimizer import prepare_optim
    from egonet_b200.libs.trainer.trainer import train_step
    cfgs = configs.tiny_cfgs('heatmap')
    m = get_pose_net(cfgs, is_train=False)
    m.load_state_dict(hrnet_ref.make_weights(cfgs, 2))
    m = m.cuda().train()
    cfgs['optimizer'] = dict(optim_type='adam', lr=1e-3, weight_decay=0.0, momentum=0.0, milestones=[9], gamma=0.1)
    optim, _ = prepare_optim(m, cfgs)
    x = egonet_ref.synth_crops(2, cfgs, 1).cuda()
    tgt = torch.rand((2, cfgs['heatmapModel']['num_joints'], 64, 64), device='cuda')
    w = torch.ones((2, cfgs['heatmapModel']['num_joints'], 1), device='cuda')
    loss = train_step(m, JointsMSELoss(True), optim, x, tgt, w)
    torch.cuda.synchronize()
    assert torch.isfinite(loss)
    print('train ok', flush=True)


if __name__ == '__main__':
    what = sys.argv[1:] or ['convs', 'network', 'train']
    for name in what:
        {'convs': convs, 'network': network, 'train': train}[name]()
    print('sanitize target done')
