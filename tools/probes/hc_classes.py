"""Per-kernel-class device time of one HC forward (all classes).  (The fp16-vs-oracle error report lives in
tests/test_gpu_parity.py: only tests, smoke() and bench.py's CPU legs may import oracle/.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from egonet_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfgs = synth.demo_cfgs()
ego = bench.build_model(cfgs, torch.device('cuda'))
x = synth.crops(B, cfgs, 0).cuda()
with torch.no_grad():
    ego.HC(x)
    classes, total = bench.profile_hc(ego, x)
print('total ms', round(total, 3))
for k, v in sorted(classes.items(), key=lambda kv: -kv[1]['ms']):
    print('%-44s %8.3f ms  n=%3d  avg %7.1f us' % (k, v['ms'], v['launches'], v['ms'] / v['launches'] * 1e3))
