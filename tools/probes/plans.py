#!/usr/bin/env python
"""Print the tcgen05 plan chosen for every conv of the demo config (EGN_TC_VERBOSE) in a given precision mode.
    python tools/probes/plans.py [fp16|fp16x2]   (on the GPU box; plans need the driver's tensor-map encoder)"""
import os
import sys

os.environ['EGN_TC_VERBOSE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
from egonet_b200 import synth  # noqa: E402
from egonet_b200.libs.model.heatmapModel.hrnet import get_pose_net  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else 'fp16x2'
cfgs = synth.demo_cfgs()
m = get_pose_net(cfgs, is_train=False, precision=prec).eval()
m.load_state_dict(synth.hc_weights(m.state_dict(), 1))
m = m.cuda()
m(synth.crops(2, cfgs, 0).cuda())
torch.cuda.synchronize()
print('ok', m.stats())
