"""Turn the ncu dumps under gpurun_out/ into the committed summaries under profiles/ (text + csv)."""
import csv, collections, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r01d'
PX = {'cfg2': 64 * 720 * 1280, 'hd1080': 64 * 1080 * 1920, 'cfg4': 16 * 1080 * 1920, 'cfg3': 32 * 288 * 512, 'cfg5': 16 * 2160 * 3840, 'mesh5': 64 * 720 * 1280}
TITLE = {'cfg2': 'warp_fwd_tile_kernel<TMODE_TPS> at cfg2 (64 x 720p, 4x4 mesh)', 'hd1080': 'warp_fwd_tile_kernel<TMODE_TPS> at the north-star shape (64 x 1080p, 4x4 mesh)', 'cfg4': 'warp_fwd_tile_kernel<TMODE_FLOW> at cfg4 (16 x 1080p + flow)',
         'cfg3': 'warp_bwd_tile_kernel<TMODE_TPS> at cfg3 (32 x 288x512, 4x4 mesh, grads wrt image, grid and T)',
         'cfg5': 'warp_fwd_tile_kernel<TMODE_TPS, G = 0> at cfg5 (16 x 2160x3840, 16x16 mesh, generic tables)', 'mesh5': 'warp_fwd_tile_kernel<TMODE_TPS, G = 5> at mesh5 (64 x 720p, 5x5 mesh)'}
for wl in (sys.argv[2:] or ['cfg2', 'hd1080', 'cfg4', 'cfg3']):
    raw = os.path.join(ROOT, 'gpurun_out', 'final_%s_raw.csv' % wl)
    sass = os.path.join(ROOT, 'gpurun_out', 'final_%s_sass.csv' % wl)
    if not os.path.exists(raw):
        continue
    out = ['# ncu --set full --clock-control none, one launch of %s' % TITLE[wl],
           '# command: python bench.py --workload %s --steps 3 --warmup 3 --no-cpu --no-e2e (after a plain run of the same command exited 0)' % wl, '']
    out.append(subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_raw.py'), raw], capture_output=True, text=True).stdout)
    out.append('# instructions per 32 output pixels by kernel phase (tools/ncu_seg.py; "lvl" = executions per 32 pixels of each instruction in the region)')
    seg = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_seg.py'), sass, str(PX[wl])], capture_output=True, text=True).stdout
    out.append('\n'.join(l for l in seg.splitlines() if not l.startswith('lvl 0.0')))
    open(os.path.join(ROOT, 'profiles', '%s_%s_ncu_full.txt' % (tag, wl)), 'w').write('\n'.join(out) + '\n')
    # launch list
    ll = os.path.join(ROOT, 'gpurun_out', 'launches_%s.csv' % wl)
    rows = [r for r in csv.reader(open(ll)) if len(r) > 10]
    with open(os.path.join(ROOT, 'profiles', '%s_%s_launches.csv' % (tag, wl)), 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none -c 80; python bench.py --workload %s --steps 3 --warmup 3 --no-cpu --no-e2e\n' % wl)
        f.write('id,kernel,block,grid,ns\n')
        agg = collections.OrderedDict()
        for r in rows[1:]:
            k = r[4] if len(r[4]) < 90 else r[4][:70] + '...'
            f.write('%s,"%s","%s","%s",%s\n' % (r[0], k, r[7], r[8], r[-1]))
            agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += float(r[-1])
        f.write('# per-kernel averages (cold-cache, serialised by ncu)\n')
        ours = sum(v[1] for k, v in agg.items() if 'dvsg' in k)
        for k, v in agg.items():
            if 'dvsg' in k:
                f.write('# %-80s n=%d avg %.1f us  share of our kernels %.1f%%\n' % (k, v[0], v[1] / v[0] / 1e3, 100 * v[1] / ours))
print('ok')
