"""SASS opcode histogram of the built library per kernel (cuobjdump -sass): proof of the Blackwell-native instructions on
the hot path (UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce-add, UBLKCP = bulk copy, MUFU.LG2, packed
FFMA2 / FADD2 / FMUL2, ATOMS = native shared-memory integer atomics, REDUX).  usage: sass_hist.py [lib.so] > profiles/..."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'coupe', 'dvsg_b200', 'libdvsg_warp.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
KEY = ['UTMALDG', 'UTMASTG', 'UTMAREDG', 'UBLKCP', 'SYNCS', 'MUFU.LG2', 'MUFU', 'FFMA2', 'FADD2', 'FMUL2', 'FFMA', 'LDS', 'STS', 'ATOMS', 'CREDUX', 'REDUX', 'RED', 'ATOMG', 'LDG', 'STG',
       'SHFL', 'DFMA', 'BAR']
kern, hist, arch = None, {}, set()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r'\(dvsg::\w+Params.*', '', kern)
        hist[kern] = collections.Counter()
        continue
    m = re.search(r'arch = (sm_\w+)', line)
    if m:
        arch.add(m.group(1))
    m = re.search(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', line)
    if m and kern:
        op = m.group(1)
        hist[kern]['_total'] += 1
        for k in KEY:
            if op == k or op.startswith(k + '.') or (k == 'MUFU.LG2' and op.startswith('MUFU.LG2')):
                hist[kern][k] += 1
print('# SASS opcode histogram, %s (cuobjdump -sass; static instruction counts per kernel); arch = %s' % (os.path.basename(lib), ', '.join(sorted(arch))))
print('# %-98s %6s %s' % ('kernel', 'instrs', ' '.join('%8s' % k for k in KEY)))
for k in sorted(hist, key=lambda k: -hist[k]['_total']):
    h = hist[k]
    if h['_total'] < 50:
        continue
    print('%-100s %6d %s' % (k[:100], h['_total'], ' '.join('%8d' % h[q] for q in KEY)))
