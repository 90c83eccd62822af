"""SM cycles per tcgen05.mma (M=128,K=16,f16,SS) vs N / accumulator rotation / CTA count / operand data."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from egonet_b200 import _native as N
torch.zeros(1).cuda()
L = N.lib()
def rate(n, nacc, shift=0, ctas=1):
    v = ctypes.c_double(0)
    N.check(L.egn_debug_umma_rate(n, nacc, 200, shift, ctas, ctypes.byref(v)))
    return v.value
print('N   nacc=1  nacc=2  nacc=max | 148 CTAs nacc=max | shift 67 rows | RANDOM data nacc=max 1 CTA / 148 CTAs | random nacc=3')
for n in (16, 32, 48, 64, 96, 128, 192, 256):
    m = max(1, min(8, 512 // n))
    print('%3d %7.1f %7.1f %7.1f (x%d) | %7.1f          | %7.1f       | %7.1f / %7.1f                   | %7.1f' % (
        n, rate(n, 1), rate(n, min(2, m)), rate(n, m), m, rate(n, m, 0, 148), rate(n, m, 67), rate(n, m, 1000), rate(n, m, 1000, 148),
        rate(n, min(3, m), 1000, 148)))
# two issuing threads (warps 1 and 2, each its own accumulators) vs one, random data, 1 CTA and 148 CTAs
print('\nN   nacc | 1 issuer (1 CTA / 148) | 2 issuers (1 CTA / 148)   [cycles per MMA, aggregate]')
for n in (48, 96, 128, 192):
    for nacc in (2, 4):
        if nacc * n > 512:
            continue
        print('%3d %4d | %7.1f / %7.1f      | %7.1f / %7.1f' % (n, nacc, rate(n, nacc, 1000), rate(n, nacc, 1000, 148),
                                                                 rate(n, nacc, 2000), rate(n, nacc, 2000, 148)))
