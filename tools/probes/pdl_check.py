"""Race probe for programmatic dependent launch: the HC forward must be bit-reproducible run to run and
identical with EGN_PDL=0 (fully serialised launches).  A missing griddepcontrol.wait shows up as a changing
digest.

    python tools/probes/pdl_check.py [--batch 256] [--runs 12]
"""
import argparse
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def child(args):
    import torch
    from egonet_b200 import synth
    from egonet_b200.libs.model.egonet import EgoNet
    cfgs = synth.demo_cfgs()
    ego = EgoNet(cfgs, pre_trained=False).eval()
    ego.HC.load_state_dict(synth.hc_weights(ego.HC.state_dict(), 1))
    ego = ego.cuda()
    xs = [synth.crops(args.batch, cfgs, s).cuda() for s in (0, 1)]
    digests = []
    with torch.no_grad():
        for r in range(args.runs):
            maps, coords = ego.HC(xs[r & 1])
            torch.cuda.synchronize()
            h = hashlib.sha1(maps.cpu().numpy().tobytes() + coords.cpu().numpy().tobytes()).hexdigest()[:16]
            digests.append(h)
    even, odd = set(digests[0::2]), set(digests[1::2])
    print('PDL=%s batch=%d even=%s odd=%s' % (os.environ.get('EGN_PDL', '1'), args.batch, sorted(even), sorted(odd)))
    print('DIGEST %s %s' % (digests[0], digests[1]))
    return 0 if len(even) == 1 and len(odd) == 1 else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--runs', type=int, default=12)
    ap.add_argument('--child', action='store_true')
    args = ap.parse_args()
    if args.child:
        sys.exit(child(args))
    out = {}
    rc = 0
    for pdl in ('1', '0'):
        env = dict(os.environ, EGN_PDL=pdl)
        r = subprocess.run([sys.executable, __file__, '--child', '--batch', str(args.batch), '--runs', str(args.runs)],
                           env=env, capture_output=True, text=True)
        print(r.stdout.strip())
        if r.returncode:
            print(r.stderr[-2000:])
            rc = 1
        out[pdl] = [l for l in r.stdout.splitlines() if l.startswith('DIGEST')]
    if out['1'] != out['0']:
        print('MISMATCH between PDL on and off')
        rc = 1
    else:
        print('PDL on == PDL off, bit-identical over %d runs' % args.runs)
    sys.exit(rc)


if __name__ == '__main__':
    main()
