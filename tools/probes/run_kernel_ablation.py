"""Ablate phases of the v2 window-run kernel (EGN_TC_DBG bits) on the 48ch@64x64 and 96ch@32x32 layers."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
names = {0: 'full', 1: 'no stores', 2: 'no residual', 3: 'no stores+res', 4: 'no MMA', 8: 'no A load', 16: 'no B loads',
         24: 'no A/B loads', 28: 'no loads, no MMA', 31: 'nothing but TMEM reads', 7: 'loads only'}
for dbg, name in names.items():
    e = dict(os.environ, EGN_TC_DBG=str(dbg))
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'layer_bench.py'), '--child', '--batch', '64', '--iters', '20'],
                       capture_output=True, text=True, env=e)
    rows = [json.loads(l[7:]) for l in r.stdout.splitlines() if l.startswith('RESULT ')]
    if not rows:
        print(dbg, name, 'FAILED', r.stderr[-500:]); continue
    print('%2d %-24s %s' % (dbg, name, '  '.join('%s: %.1f us' % (x['shape'].split(' @')[0][7:] + '@' + x['shape'].split('@')[1], x['us']) for x in rows[0][:2] + rows[0][5:7])), flush=True)
