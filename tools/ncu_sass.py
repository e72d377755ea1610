"""Summarise an `ncu --page source --csv --print-source sass` dump: instructions executed per
output-pixel warp and stall samples, per SASS instruction.  usage: ncu_sass.py file.csv n_pixels"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
npx = float(sys.argv[2])
minper = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
data = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] in ('Kernel Name', 'Address'):
        if data:
            break
        continue
    data.append(r)
tot = sum(int(r[iex]) for r in data)
tots = sum(int(r[ismp]) for r in data)
warps = npx / 32
print('total warp-inst %d = %.1f per pixel-warp; samples %d; sass lines %d' % (tot, tot / warps, tots, len(data)))
for r in data:
    ex = int(r[iex])
    if ex / warps < minper:
        continue
    print('%6s %7.2f %5.1f%%  %s' % (r[ia][-5:], ex / warps, 100.0 * int(r[ismp]) / max(tots, 1), r[isrc][:100]))

# ---- opcode histogram outside the hottest loop -------------------------------------------
import collections, re
hist = collections.Counter()
smp = collections.Counter()
for r in data:
    ex = int(r[iex]) / warps
    m = re.search(r'(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', r[isrc])
    op = m.group(1).split('.')[0] if m else '?'
    hist[op] += ex
    smp[op] += int(r[ismp])
print('--- opcode histogram (warp-inst per 32 px, %% of stall samples)')
for op, v in hist.most_common(40):
    print('%-10s %7.2f  %5.1f%%' % (op, v, 100.0 * smp[op] / max(tots, 1)))
