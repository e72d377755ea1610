"""Shared-memory wavefronts per SASS instruction from an `ncu --page source --csv --print-source sass` dump, grouped by the
equal-execution-count regions of tools/ncu_seg.py.  usage: ncu_smem.py file.csv n_pixels"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
npx = float(sys.argv[2])
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) and r[0] not in ('Kernel Name', 'Address')]
tiles = npx / 256
def f(r, k):
    try:
        return float(r[col[k]])
    except ValueError:
        return 0.0
tot_w = sum(f(r, 'L1 Wavefronts Shared') for r in data)
tot_i = sum(f(r, 'L1 Wavefronts Shared Ideal') for r in data)
print('shared-memory wavefronts per 32x8 tile: %.1f (ideal %.1f, excessive %.1f)' % (tot_w / tiles, tot_i / tiles, (tot_w - tot_i) / tiles))
agg = collections.OrderedDict()
for i, r in enumerate(data):
    w = f(r, 'L1 Wavefronts Shared')
    if w <= 0:
        continue
    m = re.search(r'(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', r[col['Source']])
    op = m.group(1)
    key = (i // 40, op)
    a = agg.setdefault(key, [0.0, 0.0, 0, i])
    a[0] += w; a[1] += f(r, 'L1 Wavefronts Shared Ideal'); a[2] += 1
for (blk, op), (w, wi, n, i0) in agg.items():
    if w / tiles >= 1.0:
        print('sass lines ~%4d  %-14s x%3d  wavefronts/tile %6.1f  ideal %6.1f  executions/tile of one instr %.2f' % (i0, op, n, w / tiles, wi / tiles, f(data[i0], 'Instructions Executed') / tiles))
