"""Run on the GPU box: which descriptor base-offset encoding makes row-shifted swizzled operands work?"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from egonet_b200 import _native as N

res = {}
for sw in (128, 64, 32):
    kc = sw // 2
    rows = torch.arange(256).view(-1, 1)
    cols = torch.arange(kc).view(1, -1)
    a = (((rows * 7 + cols * 3) % 61) - 30).to(torch.float16).cuda().contiguous()
    b = torch.eye(kc, dtype=torch.float16).cuda().contiguous()
    for off in (0, 1, 2, 3, 4, 7, 8, 9, 66, 67, 128):
        for mode in (0, 1, 2):
            out = torch.full((128, kc), float('nan'), device='cuda')
            try:
                N.check(N.lib().egn_debug_umma_probe(sw, off, mode, N.ptr(a), N.ptr(b), N.ptr(out), None))
                torch.cuda.synchronize()
                ok = bool(torch.equal(out, a[off:off + 128].float()))
                nbad = int((out != a[off:off + 128].float()).sum())
            except Exception as e:  # noqa
                ok, nbad = False, str(e)[:80]
            res['sw%d_off%d_mode%d' % (sw, off, mode)] = [ok, nbad]
            print(sw, off, mode, ok, nbad, flush=True)
json.dump(res, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'gpurun_out', 'probe_row_offset.json'), 'w'), indent=0)
