"""Per-CTA phase timeline (globaltimer stamps) of the v2 window-run kernel."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for dbg in (0, 31, 28, 4):
    e = dict(os.environ, EGN_TC_DBG=str(dbg), EGN_TC_TS='1', EGN_TC_TS_DUMP='1', EGN_TC_VERBOSE='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'layer_bench.py'), '--child', '--batch', '64', '--iters', '1'],
                       capture_output=True, text=True, env=e)
    print('=== dbg', dbg)
    for l in r.stderr.splitlines():
        if l.startswith('[egn'):
            print(l)
