"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: ncu_lines.py file.csv n_pixels [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
npx = float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
per_line = collections.Counter(); smp = collections.Counter(); src = {}
cur = None; hdr = None; fname = ''
for r in rows:
    if len(r) >= 2 and r[0] == 'File Name':
        fname = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == 'Line No':
        hdr = r; iex = hdr.index('Instructions Executed'); ismp = hdr.index('# Samples'); continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    if r[0] != '':
        cur = (fname, int(r[0])); src[cur] = r[1].strip()[:90]
    if r[2] not in ('', '...') and cur is not None:
        try:
            per_line[cur] += int(r[iex]); smp[cur] += int(r[ismp])
        except ValueError:
            pass
warps = npx / 32
tot = sum(per_line.values()); tots = sum(smp.values())
print('total %.1f warp-inst per 32 px, %d samples' % (tot / warps, tots))
for k, v in per_line.most_common(top):
    print('%-22s %6d  %7.2f  %5.1f%%  %s' % (k[0][:22], k[1], v / warps, 100.0 * smp[k] / max(tots, 1), src[k]))
