"""Turn the ncu dumps of tools/gpu_prof2.sh (gpurun_out/<tag>_<wl>_raw.csv, _sass.csv, <tag>_launches_<wl>.csv) into the
committed summaries under profiles/ (text + csv).  usage: mk_profiles2.py <src tag> <out tag> <kernel word> wl..."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src_tag, out_tag, kword = sys.argv[1], sys.argv[2], sys.argv[3]
PX = {'cfg2': 64 * 720 * 1280, 'hd1080': 64 * 1080 * 1920, 'cfg4': 16 * 1080 * 1920, 'cfg3': 32 * 288 * 512, 'cfg3m5': 32 * 288 * 512,
      'cfg5': 16 * 2160 * 3840, 'mesh5': 64 * 720 * 1280, 'cfg1': 288 * 512}
TITLE = {'cfg2': 'cfg2 (64 x 720p, 4x4 mesh)', 'hd1080': 'the north-star shape (64 x 1080p, 4x4 mesh)', 'cfg4': 'cfg4 (tf_warp, 16 x 1080p + flow)',
         'cfg3': 'cfg3 (32 x 288x512, 4x4 mesh, fwd + bwd)', 'cfg3m5': 'the training shape with the 5x5 mesh', 'cfg5': 'cfg5 (16 x 2160x3840, 16x16 mesh)',
         'mesh5': 'mesh5 (64 x 720p, 5x5 mesh)', 'cfg1': 'cfg1 (1 x 288x512, online loop)'}
traffic_path = os.path.join(ROOT, 'profiles', 'traffic.json')
try:
    traffic = json.load(open(traffic_path))
except Exception:
    traffic = {}
for wl in sys.argv[4:]:
    raw = os.path.join(ROOT, 'gpurun_out', '%s_%s_raw.csv' % (src_tag, wl))
    sass = os.path.join(ROOT, 'gpurun_out', '%s_%s_sass.csv' % (src_tag, wl))
    if not os.path.exists(raw):
        print('missing', raw)
        continue
    rows = list(csv.reader(open(raw)))
    hdr, data = rows[0], rows[2]
    kname = data[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else kword
    out = ['# ncu --set full --clock-control none --import-source on, one launch of %s at %s' % (kname[:90], TITLE[wl]),
           '# command: DVSG_BENCH_MIN_S=0.02 python bench.py --workload %s --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras' % wl,
           '# (captured after a plain run of the same command exited 0; the duration below is cold-cache and serialised by ncu --',
           '#  the bench line carries the live CUDA-event time)', '']
    out.append(subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_raw.py'), raw], capture_output=True, text=True).stdout)
    out.append('# shared-memory wavefronts (source page, per 32x8 output tile)')
    out.append(subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_smem.py'), sass, str(PX[wl])], capture_output=True, text=True).stdout.splitlines()[0])
    out.append('')
    out.append('# instructions per 32 output pixels by kernel phase (tools/ncu_seg.py; "lvl" = executions per 32 pixels of each instruction in the region)')
    seg = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_seg.py'), sass, str(PX[wl])], capture_output=True, text=True).stdout
    out.append('\n'.join(l for l in seg.splitlines() if not l.startswith('lvl 0.0')))
    open(os.path.join(ROOT, 'profiles', '%s_%s_ncu_full.txt' % (out_tag, wl)), 'w').write('\n'.join(out) + '\n')
    try:
        rd, wr = float(data[hdr.index('dram__bytes_read.sum')]), float(data[hdr.index('dram__bytes_write.sum')])
        unit_r, unit_w = rows[1][hdr.index('dram__bytes_read.sum')], rows[1][hdr.index('dram__bytes_write.sum')]
        mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        traffic[wl] = int(rd * mult[unit_r] + wr * mult[unit_w])
    except Exception as e:
        print('traffic', wl, e)
    ll = os.path.join(ROOT, 'gpurun_out', '%s_launches_%s.csv' % (src_tag, wl))
    if os.path.exists(ll):
        lrows = [r for r in csv.reader(open(ll)) if len(r) > 10]
        with open(os.path.join(ROOT, 'profiles', '%s_%s_launches.csv' % (out_tag, wl)), 'w') as f:
            f.write('# ncu --metrics gpu__time_duration.sum --clock-control none -c 60; DVSG_BENCH_MIN_S=0.02 python bench.py --workload %s --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras\n' % wl)
            f.write('id,kernel,block,grid,ns\n')
            agg = collections.OrderedDict()
            for r in lrows[1:]:
                k = r[4] if len(r[4]) < 90 else r[4][:70] + '...'
                f.write('%s,"%s","%s","%s",%s\n' % (r[0], k, r[7], r[8], r[-1]))
                agg.setdefault(k, [0, 0.0])
                agg[k][0] += 1
                agg[k][1] += float(r[-1])
            f.write('# per-kernel averages (cold-cache, serialised by ncu)\n')
            ours = sum(v[1] for k, v in agg.items() if 'dvsg' in k)
            for k, v in agg.items():
                if 'dvsg' in k:
                    f.write('# %-80s n=%d avg %.1f us  share of our kernels %.1f%%\n' % (k, v[0], v[1] / v[0] / 1e3, 100 * v[1] / ours))
traffic['_source'] = 'ncu --set full captures %s (dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel), profiles/%s_*_ncu_full.txt' % (out_tag, out_tag)
json.dump(traffic, open(traffic_path, 'w'), indent=1, sort_keys=True)
print('ok', sorted(traffic))
