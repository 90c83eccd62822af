"""Run on the GPU box: SM cycles per K16 slice of the fp16x2 MMA sequences, by A-operand layout (egn_debug_umma_seq)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from egonet_b200 import _native as N
L = N.lib()
names = {0: 'single N=n', 1: 'wide 2n + narrow n', 2: 'three N=n (H,L,L)', 3: 'grouped 2 wide, 2 narrow'}
for ctas in (1, 148):
    print('== %d CTAs' % ctas)
    print('%-26s %5s | %s' % ('pattern', 'n', '  '.join('sw%d/sbo%d' % (sw, sbo) for sw, sbo in ((128, 8), (128, 18), (64, 8), (64, 10)))))
    for n in (48, 96, 192):
        for pat in (0, 1, 2, 3):
            if pat in (1, 3) and 2 * n > 256:
                continue
            row = []
            for sw, sbo in ((128, 8), (128, 18), (64, 8), (64, 10)):
                v = ctypes.c_double(0)
                N.check(L.egn_debug_umma_seq(n, pat, 360, sw, sbo, ctas, ctypes.byref(v)))
                row.append('%9.1f' % v.value)
            print('%-26s %5d | %s' % (names[pat], n, '  '.join(row)), flush=True)
