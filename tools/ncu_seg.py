"""Segment an `ncu --page source --csv --print-source sass` dump into regions of equal execution count and
print instructions per 32 output pixels per region.  usage: ncu_seg.py file.csv n_pixels"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
npx = float(sys.argv[2])
hdr = rows[1]
isrc, iex, ismp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
data = [r for r in rows[2:] if len(r) >= len(hdr) and r[0] not in ('Kernel Name', 'Address')]
warps = npx / 32
tots = sum(int(r[ismp]) for r in data)
seg = []; cur = None
for i, r in enumerate(data):
    ex = int(r[iex]) / warps
    lvl = round(ex, 2)
    if cur is None or abs(cur[0] - lvl) > 0.015:
        cur = [lvl, i, i, 0.0, collections.Counter(), 0]
        seg.append(cur)
    cur[2] = i; cur[3] += ex; cur[5] += int(r[ismp])
    m = re.search(r'(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', r[isrc]); cur[4][m.group(1).split('.')[0]] += 1
print('total %.1f warp-inst per 32 px' % (sum(int(r[iex]) for r in data) / warps))
for s in seg:
    if s[3] > 0.8:
        print('lvl %.3f lines %4d-%4d (%4d instr) total %5.1f  samples %4.1f%%  %s' % (s[0], s[1], s[2], s[2] - s[1] + 1, s[3], 100.0 * s[5] / tots, dict(s[4].most_common(8))))
