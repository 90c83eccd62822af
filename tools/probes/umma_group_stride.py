"""Run on the GPU box: does a K-major swizzled UMMA descriptor accept a stride byte offset that is NOT a multiple of
the 8-row swizzle atom (8-row groups `gs` operand rows apart)?  Tile row i must read operand row off + (i//8)*gs + i%8."""
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
    for gs in (8, 9, 10, 12, 15, 16):
        for off in (0, 1, 3, 8, 11):
            if off + 15 * gs + 8 > 256:
                continue
            idx = torch.tensor([off + (i // 8) * gs + i % 8 for i in range(128)])
            want = a[idx.cuda()].float()
            out = torch.full((128, kc), float('nan'), device='cuda')
            try:
                N.check(N.lib().egn_debug_umma_probe(sw, off, gs << 4, N.ptr(a), N.ptr(b), N.ptr(out), None))
                torch.cuda.synchronize()
                ok = bool(torch.equal(out, want))
                nbad = int((out != want).sum())
            except Exception as e:  # noqa
                ok, nbad = False, str(e)[:80]
            res['sw%d_gs%d_off%d' % (sw, gs, off)] = [ok, nbad]
            print(sw, gs, off, ok, nbad, flush=True)
json.dump(res, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'gpurun_out', 'probe_group_stride.json'), 'w'), indent=0)
