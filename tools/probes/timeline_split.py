"""Per-window timeline (EGN_TC_TS) and phase ablation (EGN_TC_DBG) of the persistent kernel in fp16x2 storage.

    python tools/probes/timeline_split.py [--shapes 1] [--batch 256]
dbg bits: 8 = no window loads, 64 = epilogue skipped (accumulators released at once), 1 = no output stores, 2 = no residual
"""
import argparse, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument('--shapes', type=int, default=1)
ap.add_argument('--first', type=int, default=0)
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--dbg', default='0,64,8,72,1,3')
ap.add_argument('--env', default='')
a = ap.parse_args()
extra = dict(kv.split('=') for kv in a.env.split(',') if kv)
for dbg in [int(x) for x in a.dbg.split(',')]:
    for ts in (0, 1):
        e = dict(os.environ, EGN_TC_DBG=str(dbg), EGN_TC_VERBOSE='1', **extra)
        if ts:
            e.update(EGN_TC_TS='1', EGN_TC_TS_DUMP='1')
        r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'layer_bench.py'), '--child', '--batch', str(a.batch),
                            '--iters', '1' if ts else '20', '--dtype', '2', '--shapes', str(a.shapes), '--first', str(a.first)],
                           capture_output=True, text=True, env=e)
        print('=== dbg', dbg, 'timeline' if ts else 'timing')
        if ts:
            seen = 0
            for l in r.stderr.splitlines():
                if l.startswith('[egn-ts3]') and seen < 40:
                    print(l)
                    seen += 1
        else:
            for l in r.stdout.splitlines():
                if l.startswith('RESULT'):
                    print(l)
            for l in r.stderr.splitlines():
                if l.startswith('[egn] conv'):
                    print(l)
        if r.returncode:
            print('FAILED', r.stderr[-2000:])
