"""Time single fused conv layers (tcgen05 v1 per-tap / v2 window-run / CUDA-core) at a given batch.

    python tools/layer_bench.py [--batch 64] [--iters 30]      (on the GPU box)
"""
import argparse, ctypes, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = [  # Cin, Cout, H, W, k, stride, residual
    (48, 48, 64, 64, 3, 1, 1), (96, 96, 32, 32, 3, 1, 1), (192, 192, 16, 16, 3, 1, 1), (384, 384, 8, 8, 3, 1, 1),
    (64, 64, 64, 64, 3, 1, 0), (256, 64, 64, 64, 1, 1, 0), (64, 256, 64, 64, 1, 1, 1), (64, 64, 128, 128, 3, 2, 0),
    (96, 48, 32, 32, 1, 1, 0), (384, 48, 8, 8, 1, 1, 0), (48, 96, 64, 64, 3, 2, 0),
]


def run_child(args):
    import torch
    from egonet_b200 import _native as N
    B = args.batch
    out_rows = []
    for (Cin, Cout, H, W, k, st, has_res) in SHAPES[args.first:args.shapes]:
        g = torch.Generator().manual_seed(1)
        Cip, Cop = (Cin + 15) // 16 * 16, (Cout + 15) // 16 * 16
        mul = 2 if args.dtype == 2 else 1          # fp16x2: [hi | lo] planes (random lo planes are fine for timing)
        x = torch.randn((B, H, W, mul * Cip), generator=g).to(torch.float16).cuda()
        pad = 1 if k == 3 else 0
        OH, OW = (H + 2 * pad - k) // st + 1, (W + 2 * pad - k) // st + 1
        res = torch.randn((B, OH, OW, mul * Cop), generator=g).to(torch.float16).cuda() if has_res else None
        out = torch.empty((B, OH, OW, mul * Cop), dtype=torch.float16, device='cuda')
        w = (torch.randn((Cout, Cin, k, k), generator=g) / (Cin * k * k) ** 0.5).contiguous()
        ms = ctypes.c_float(0)
        N.check(N.lib().egn_conv2d_bench(1, args.dtype, N.ptr(x), N.ptr(w), None, N.ptr(res), N.ptr(out), B, H, W, Cin, Cout,
                                         k, st, 1, None, args.iters, ctypes.byref(ms)))
        flops = 2.0 * B * OH * OW * Cout * Cin * k * k
        byts = 2.0 * (x.numel() + out.numel() + (res.numel() if has_res else 0)) + 2.0 * Cop * Cip * k * k
        out_rows.append({'shape': '%dx%d s%d %d->%d @%dx%d%s' % (k, k, st, Cin, Cout, OH, OW, '+res' if has_res else ''),
                         'us': round(ms.value * 1e3, 2), 'tflops': round(flops / ms.value / 1e9, 1),
                         'gbs': round(byts / ms.value / 1e6, 1)})
    print('RESULT ' + json.dumps(out_rows))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--iters', type=int, default=30)
    ap.add_argument('--child', action='store_true')
    ap.add_argument('--dtype', type=int, default=1, help='1 = fp16, 2 = fp16x2 split storage')
    ap.add_argument('--variants', default='', help='extra variants: name:ENV=V,ENV=V;name2:...')
    ap.add_argument('--shapes', type=int, default=len(SHAPES), help='only the first N shapes')
    ap.add_argument('--first', type=int, default=0, help='skip the first N shapes')
    args = ap.parse_args()
    if args.child:
        return run_child(args)
    table = {}
    variants = [('v3_pers', {}), ('v3_nopair', {'EGN_TC_PAIR': '0'}), ('v2_run', {'EGN_TC_V3': '0', 'EGN_TC_V2_SPLIT': '1'}),
                ('v1_tap', {'EGN_TC_V3': '0', 'EGN_TC_V2': '0'})]
    for item in filter(None, args.variants.split(';')):
        name, envs = item.split(':')
        variants.append((name, dict(kv.split('=') for kv in envs.split(','))))
    for label, env in variants:
        e = dict(os.environ, EGN_TC_VERBOSE='1', **env)
        r = subprocess.run([sys.executable, __file__, '--child', '--batch', str(args.batch), '--iters', str(args.iters),
                            '--dtype', str(args.dtype), '--shapes', str(args.shapes), '--first', str(args.first)],
                           capture_output=True, text=True, env=e)
        cfg = {}
        for l in r.stderr.splitlines():
            if l.startswith('[egn] conv '):
                key = l[11:].split(':')[0].replace(' fp16x2', '')
                cfg[key] = l.split(': ', 1)[1]
        rows = [json.loads(l[7:]) for l in r.stdout.splitlines() if l.startswith('RESULT ')]
        if not rows:
            print(label, 'FAILED', r.stdout[-1500:], r.stderr[-3000:])
            continue
        for row in rows[0]:
            table.setdefault(row['shape'], {})[label] = (row, cfg.get(row['shape'].replace('+res', ''), ''))
    for shape, d in table.items():
        print(shape)
        for label, (row, c) in d.items():
            print('   %-7s %8.2f us  %7.1f TFLOP/s  %7.1f GB/s   %s' % (label, row['us'], row['tflops'], row['gbs'], c))


if __name__ == '__main__':
    main()
