#!/usr/bin/env python
"""Measured parity of every HC precision mode against the reference-generated goldens (tests/golden):
coordinate / heat-map error on the four HC configs and end-to-end pipeline error (screen key-points, 3D
key-points, Euler angles, alpha) on the tiny and demo configs.  Writes gpurun_out/parity_report.json.
    python tools/probes/parity_report.py            (on the GPU box)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import configs, egonet_ref, hrnet_ref, lifter_ref  # noqa: E402
from egonet_b200.libs.model.egonet import EgoNet  # noqa: E402
from egonet_b200.libs.model.heatmapModel.hrnet import get_pose_net  # noqa: E402
import test_gpu_parity as T  # noqa: E402


def golden(name):
    return np.load(os.path.join(ROOT, 'tests', 'golden', name), allow_pickle=False)


def main():
    rep = {'hc': {}, 'pipeline': {}}
    for tag, mk in T.HC_CASES:
        cfgs = mk()
        g = golden('hrnet_%s.npz' % tag)
        x = egonet_ref.synth_crops(int(g['batch']), cfgs, int(g['seed_x'])).cuda()
        for prec in ('fp32', 'fp16x2', 'fp16'):
            m = get_pose_net(cfgs, is_train=False, precision=prec).eval()
            m.load_state_dict(hrnet_ref.make_weights(cfgs, 1))
            out = m.cuda()(x)
            maps = (out[0] if isinstance(out, tuple) else out).cpu().numpy()
            rs = int(g['map_row_stride'])
            row = {'maps_err_rel_to_max': float(np.abs(maps[:, :, ::rs, :] - g['maps_sub']).max() / np.abs(g['maps_sub']).max()),
                   'argmax_flips': int((maps.reshape(maps.shape[0], maps.shape[1], -1).argmax(2) != g['maps_argmax']).sum())}
            if isinstance(out, tuple):
                row['coords_err'] = float(np.abs(out[1].cpu().numpy() - g['coords']).max())
            rep['hc']['%s/%s' % (tag, prec)] = row
            print(tag, prec, row, flush=True)
    for tag in ('tiny', 'demo'):
        cfgs = configs.tiny_cfgs() if tag == 'tiny' else configs.demo_cfgs()
        g = golden('pipeline_%s.npz' % tag)
        for prec in ('fp32', 'fp16x2', 'fp16'):
            got = T._run_pipeline(T._egonet(cfgs, prec), cfgs, g)
            row = {k: float(np.abs(got[k] - g[k]).max()) for k in got}
            rep['pipeline']['%s/%s' % (tag, prec)] = row
            print('pipeline', tag, prec, row, flush=True)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, 'gpurun_out', 'parity_report.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
