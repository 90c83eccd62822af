"""Per-kernel-class device time of one HC forward (all classes) + fp16-vs-model coordinate error."""
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
if B <= 8:
    from oracle import hrnet_ref
    sd = hrnet_ref.make_weights(cfgs, 1)
    xc = x.cpu()
    with torch.no_grad():
        maps, coords = ego.HC(x)
    for stem16 in (True, False):
        mq, cq = hrnet_ref.hrnet_forward(sd, cfgs, xc, ctx=hrnet_ref.Quantized(torch.float16, stem_fp16=stem16))
        print('vs model stem_fp16=%s: coords max %.2e mean %.2e maps max %.2e (scale %.2f)' % (
            stem16, (coords.cpu() - cq).abs().max(), (coords.cpu() - cq).abs().mean(), (maps.cpu() - mq).abs().max(), mq.abs().max()))
    me, ce = hrnet_ref.hrnet_forward(sd, cfgs, xc)
    print('vs fp32 reference: coords max %.2e mean %.2e' % ((coords.cpu() - ce).abs().max(), (coords.cpu() - ce).abs().mean()))
