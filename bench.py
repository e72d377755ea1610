#!/usr/bin/env python
"""bench.py -- throughput of the DVSG frame-warping hot path on B200 (one JSON line).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is `passes_per_step` passes of the hot path over one batch of synthetic frames (one pass = TPS coefficient solve
+ fused TPS-grid/bilinear warp, i.e. one ThinPlateSpline forward, for the TPS workloads; the dense flow warp for cfg4).
passes_per_step is chosen after the warm-up so that the K timed steps last >= 0.5 s (a sustained number with >= 20 clock
samples), and is reported in `timing`.  Default workload = BASELINE.json configs[1]: batch 64 synthetic 720p RGB frames,
4x4 control mesh, 1 B200; the same run also measures the north-star shape (64 x 1080p, 4x4) into `north_star`.
Frames shard across ranks with no collective (weak scaling: every rank processes its own full batch); the only
torch.distributed use is the barrier and the max-over-ranks of the elapsed time.

`value`      warped Mpix/s, inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`        same metric through the host-buffer entry point (pinned host frames in, warped frames back in host memory;
             H2D and D2H inside the timed region).
`roofline`   dominant kernel (fused warp) against the measured HBM copy bandwidth; per-launch CUDA events on >= 16 launches.
`north_star` value / roofline of ThinPlateSpline fwd at 64 x 1080p, 4x4 mesh (the target shape of BASELINE.json north_star).
`strong`     fixed-size jobs sharded over the ranks (BASELINE configs[3]: 16 x 1080p tf_warp; configs[4]: a 512-frame 4K clip
             with the 16x16 mesh), value = job pixels / max-over-ranks time.
`cpu_baseline` the reference's CPU implementation on a bounded sample (TensorFlow if importable, else the NumPy oracle).
`--impl reference` times that alone on all host threads (rank 0 only), same `config`.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: kind, frames per GPU, H, W, mesh rows/cols, algorithmic bytes per output px
    'cfg1': dict(kind='tps', B=1, H=288, W=512, mesh=4, bpp=24, desc='ThinPlateSpline fwd, batch 1 288x512x3, 4x4 mesh (BASELINE configs[0])'),
    'cfg2': dict(kind='tps', B=64, H=720, W=1280, mesh=4, bpp=24, desc='ThinPlateSpline fwd, batch 64 720p RGB, 4x4 mesh (BASELINE configs[1])'),
    'hd1080': dict(kind='tps', B=64, H=1080, W=1920, mesh=4, bpp=24, desc='ThinPlateSpline fwd, batch 64 1080p RGB, 4x4 mesh (north_star headline shape)'),
    'mesh5': dict(kind='tps', B=64, H=720, W=1280, mesh=5, bpp=24, desc='ThinPlateSpline fwd, batch 64 720p RGB, 5x5 mesh (the model\'s mesh)'),
    'cfg3': dict(kind='tps_train', B=32, H=288, W=512, mesh=4, bpp=32 + 56, desc='ThinPlateSpline fwd+bwd (grads wrt image and grid), batch 32 288x512, 4x4 mesh (BASELINE configs[2])'),
    'cfg3mask': dict(kind='tps_train', B=32, H=288, W=512, mesh=4, bpp=36 + 56, mask=True,
                     desc='ThinPlateSplineWithMask fwd+bwd (image + validity mask from one pass, model.py:81-85), batch 32 288x512, 4x4 mesh'),
    'cfg3m5': dict(kind='tps_train', B=32, H=288, W=512, mesh=5, bpp=32 + 56,
                   desc='ThinPlateSpline fwd+bwd, batch 32 288x512, 5x5 mesh (the training shape with the mesh model.py:18-19 really uses)'),
    'cfg4': dict(kind='flow', B=16, H=1080, W=1920, mesh=0, bpp=32, desc='tf_warp dense flow warp, batch 16 1080p + flow (BASELINE configs[3])'),
    'cfg5': dict(kind='tps', B=16, H=2160, W=3840, mesh=16, bpp=24, desc='ThinPlateSpline fwd, 4K frames, 16x16 mesh, ring of 16 resident frames (BASELINE configs[4])'),
}
MIN_TIMED_S = float(os.environ.get('DVSG_BENCH_MIN_S', '0.5'))


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def config_for(wl, world):
    """The workload description both arms print (identical keys and values: the driver compares them)."""
    bytes_per_pass = wl['B'] * wl['H'] * wl['W'] * (32 if wl['kind'] in ('flow', 'tps_train') else 24)
    return {'workload': wl['desc'], 'frames_per_gpu': wl['B'], 'height': wl['H'], 'width': wl['W'], 'channels': 3,
            'mesh': wl['mesh'], 'parallelism': 'frame-sharded x%d, no collective' % world,
            'passes_per_step': 'auto: the K timed steps last >= %.1f s (actual count in `timing`)' % MIN_TIMED_S,
            'l2': ('inputs+outputs per pass = %.0f MB > 126 MB L2 (no flush needed)' % (bytes_per_pass / 1e6)) if bytes_per_pass > 2.5e8
            else ('working set %.0f MB fits L2: the HBM fraction is an upper bound' % (bytes_per_pass / 1e6))}


class ClockSampler(object):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, device_index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.period = float(os.environ.get("DVSG_CLOCK_PERIOD", "0.02"))
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(visible.split(',')[device_index]) if visible and visible.replace(',', '').isdigit() else device_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {}
        for n in dir(nv):
            if n.startswith('nvmlClocksThrottleReason') or n.startswith('nvmlClocksEventReason'):
                v = getattr(nv, n)
                if isinstance(v, int) and v not in (0,):
                    names.setdefault(v, n.replace('nvmlClocksThrottleReason', '').replace('nvmlClocksEventReason', ''))
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, n in names.items():
                    if mask & bit and bit & (bit - 1) == 0:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self):
        if not self.samples:
            return None
        s = sorted(self.samples)
        reasons = sorted(r for r in self.reasons if r not in ('GpuIdle', 'None', 'All'))
        return {'sm_mhz': s[len(s) // 2], 'sm_max_mhz': self.max_mhz, 'reasons': reasons, 'samples': len(s)}


# ---------------------------------------------------------------------------------------------
# CPU side: the reference's own CPU implementation.  TensorFlow (the unmodified reference files under tf.compat.v1) when it
# is importable and the reference tree is reachable; otherwise the NumPy oracle, an op-for-op port of the same TF graph.
# ---------------------------------------------------------------------------------------------
def cpu_run(wl, n_frames, threads, seed=0):
    """Time the NumPy oracle on `n_frames` frames of workload `wl`. Returns (seconds, pixels)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import dvsg_oracle as O
    rng = np.random.default_rng(seed)
    H, W = wl['H'], wl['W']
    if wl['kind'] == 'flow':
        ims = rng.random((n_frames, H, W, 3), dtype=np.float32)
        flows = rng.uniform(-8, 8, (n_frames, H, W, 2)).astype(np.float32)

        def one(i):
            return O.tf_warp(ims[i:i + 1], flows[i:i + 1], H, W)
    else:
        m = wl['mesh']
        coord = O.regular_mesh(m, m)[None]
        ims = rng.random((n_frames, H, W, 3), dtype=np.float32)
        vecs = rng.uniform(-0.1, 0.1, (n_frames, m * m, 2)).astype(np.float32)

        def one(i):
            return O.thin_plate_spline(ims[i:i + 1], coord, vecs[i:i + 1], (H, W))
    t0 = time.perf_counter()
    if threads <= 1:
        for i in range(n_frames):
            one(i)
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(one, range(n_frames)))
    return time.perf_counter() - t0, n_frames * H * W


def tf_run(v1, wl, n_frames, threads, seed=0, _cache={}):
    """Same sample through the UNMODIFIED reference under tensorflow.compat.v1 (graph built once per shape)."""
    import numpy as np
    from oracle import dvsg_oracle as O
    from oracle import tf_reference
    rng = np.random.default_rng(seed)
    H, W = wl['H'], wl['W']
    key = (wl['kind'], n_frames, H, W, wl['mesh'])
    if wl['kind'] == 'flow':
        run = _cache.get(key) or _cache.setdefault(key, tf_reference.FlowRunner(v1, (n_frames, H, W, 3), threads))
        args = (rng.random((n_frames, H, W, 3), dtype=np.float32), rng.uniform(-8, 8, (n_frames, H, W, 2)).astype(np.float32))
    else:
        m = wl['mesh']
        run = _cache.get(key) or _cache.setdefault(key, tf_reference.TpsRunner(v1, (n_frames, H, W, 3), m * m, (H, W), threads))
        coord = np.tile(O.regular_mesh(m, m)[None], (n_frames, 1, 1))
        args = (rng.random((n_frames, H, W, 3), dtype=np.float32), coord, rng.uniform(-0.1, 0.1, coord.shape).astype(np.float32))
    t0 = time.perf_counter()
    run(*args)
    return time.perf_counter() - t0, n_frames * H * W


def cpu_sample_plan(wl, threads, want_frames=None):
    """(threads, frames) of one CPU sample: the whole batch of the workload when memory allows, else as many frames as fit."""
    px = wl['H'] * wl['W']
    basis = (wl['mesh'] ** 2 + 3) if wl['kind'] != 'flow' else 4
    # memory guard: the oracle materialises ~6 x [N, H*W] fp32 per frame in flight
    mem_per_frame = 6 * basis * px * 4
    try:
        import psutil
        budget = psutil.virtual_memory().available * 0.5
    except Exception:
        budget = 16e9
    threads = max(1, min(threads, int(budget // max(mem_per_frame, 1))))
    frames = wl['B'] if want_frames is None else want_frames
    return threads, max(1, frames)


def reference_impl():
    """('reference', tf.compat.v1 module, note) when TensorFlow can run the unmodified files here, else ('port', None, why)."""
    try:
        from oracle import tf_reference
        v1, why = tf_reference.probe()
    except Exception as e:      # never let the probe break the arm
        v1, why = None, 'probe failed: %s' % e
    if v1 is not None:
        return 'reference', v1, 'unmodified reference files under tensorflow.compat.v1 (CPU)'
    return 'port', None, why


def run_reference(args, wl):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    world = int(os.environ.get('WORLD_SIZE', '1'))
    kind, v1, note = reference_impl()
    threads, frames = cpu_sample_plan(wl, min(os.cpu_count() or 1, 64))
    run = (lambda n, seed: tf_run(v1, wl, n, threads, seed)) if v1 is not None else (lambda n, seed: cpu_run(wl, n, threads, seed))
    # bounded run: a step is the workload's own batch (config.frames_per_gpu frames) unless warmup + steps of it would
    # exceed DVSG_REF_BUDGET_S seconds on this host; then the largest sample that fits (one calibration frame per thread
    # measures the host first), and `cpu_baseline.sample` says so
    budget = float(os.environ.get('DVSG_REF_BUDGET_S', '170'))
    t_cal, _ = run(threads, 1)
    est_step = t_cal * frames / threads      # `threads` frames took t_cal with every host thread busy
    total_steps = args.steps + args.warmup
    if est_step * total_steps > budget:
        frames = max(1, int(frames * budget / (est_step * total_steps)))
    for _ in range(args.warmup):
        run(frames, 1)
    total_t, total_px = 0.0, 0
    for s in range(args.steps):
        t, px = run(frames, 2 + s)
        total_t += t
        total_px += px
    value = total_px / total_t / 1e6
    sample = ('%d frames per step (%s), %d host threads; %s' %
              (frames, 'the whole batch of the workload' if frames == wl['B'] else 'bounded sample of the %d-frame batch' % wl['B'], threads,
               note if v1 is not None else 'NumPy fp32 oracle, an op-for-op port of the reference TF graph, one frame per thread; probe: ' + str(note)))
    line = {
        'impl': 'reference', 'metric': 'warped Mpix/s', 'value': value, 'unit': 'Mpix/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total_t / max(args.steps, 1),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': config_for(wl, world),
        'timing': {'clock': 'host wall clock (CPU arm)', 'frames_per_step': frames},
        'cpu_baseline': {'value': value, 'unit': 'Mpix/s', 'cores': threads, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': 'Mpix/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------
class Pass(object):
    """One pass of the hot path over a resident batch: builds the inputs once, `run(sample)` enqueues the kernels."""

    def __init__(self, wl, B, dev, seed, lib):
        import torch
        from coupe.dvsg_b200 import _lib, ops
        self.torch, self._lib, self.ops, self.lib, self.wl, self.dev = torch, _lib, ops, lib, wl, dev
        self.B, self.H, self.W, self.kind = B, wl['H'], wl['W'], wl['kind']
        H, W = self.H, self.W
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        self.U = torch.rand((B, H, W, 3), device=dev, generator=g)
        self.stream = torch.cuda.current_stream(dev)
        self.fwd_pairs, self.bwd_pairs = [], []
        self.pix = B * H * W
        if self.kind == 'flow':
            lat = (torch.rand((B, 2, 9, 16), device=dev, generator=g) - 0.5) * 16.0
            flow = torch.nn.functional.interpolate(lat, size=(H, W), mode='bilinear', align_corners=True)
            self.flow = (flow + (torch.rand((B, 2, H, W), device=dev, generator=g) - 0.5)).permute(0, 2, 3, 1).contiguous()
            self.out = torch.empty_like(self.U)
            self.kernel_name = 'warp_fwd_tile_kernel<TMODE_FLOW> (warp_fwd_tile.cu)'
            self.fwd_bpp = 32
        else:
            m = wl['mesh']
            lin = torch.arange(m, device=dev, dtype=torch.float32) * (2.0 / (m - 1)) - 1.0
            self.mesh = torch.stack(torch.meshgrid(lin, lin, indexing='xy'), dim=-1).reshape(m * m, 2).contiguous()
            self.coord = self.mesh.unsqueeze(0).expand(B, -1, -1)     # one mesh shared by the batch (model.py:68)
            self.vec = (torch.rand((B, m * m, 2), device=dev, generator=g) - 0.5) * 0.2
            self.train = self.kind == 'tps_train'
            self.mask = bool(wl.get('mask'))
            self.g_out = torch.rand((B, H, W, 3), device=dev, generator=g) if self.train else None
            # training step: grad_image buffers are preallocated and zero-filled on a side stream while the solve and the
            # forward kernel run (the fill is part of the step and of the 56 B/px; it only leaves the critical path)
            self.gU_bufs = [torch.empty_like(self.U) for _ in range(2)] if self.train else None
            self.zero_stream = torch.cuda.Stream(dev) if self.train else None
            self.step_no = 0
            self.online = ops.OnlineWarper(self.mesh, B, H, W) if (B == 1 and not self.train) else None     # cfg1: eval.py's loop
            self.kernel_name = 'warp_fwd_tile_kernel<TMODE_TPS> (warp_fwd_tile.cu)'
            self.fwd_bpp = (36 if self.mask else 32) if self.train else 24

    def _events(self):
        return self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)

    def run(self, sample=False):
        torch, ops, lib, stream = self.torch, self.ops, self.lib, self.stream
        B, H, W = self.B, self.H, self.W
        if self.kind == 'flow':
            if sample:
                e0, e1 = self._events()
                e0.record(stream)
            rc = lib.dvsg_flow_warp_fwd(self.U.data_ptr(), self.flow.data_ptr(), self.out.data_ptr(), B, H, W, 3, 0, stream.cuda_stream)
            if sample:
                e1.record(stream)
                self.fwd_pairs.append((e0, e1))
            self._lib.check(rc, 'dvsg_flow_warp_fwd')
            return
        if self.online is not None:
            if sample:
                e0, e1 = self._events()
                e0.record(stream)
            self.online.warp(self.U, self.vec)
            if sample:
                e1.record(stream)
                self.fwd_pairs.append((e0, e1))
            return
        if self.train:
            gU = self.gU_bufs[self.step_no & 1]
            self.step_no += 1
            self.zero_stream.wait_stream(stream)      # the backward that last accumulated into this buffer has been issued
            with torch.cuda.stream(self.zero_stream):
                gU.zero_()
        T = ops.tps_solve(self.coord, self.vec, offsets=True)      # target = coord + vec is formed inside the prepared solve
        if sample:
            e0, e1 = self._events()
            e0.record(stream)
        ops.tps_warp_fwd(self.U, self.coord, T, (H, W), want_grid=self.train, want_mask=self.train and self.mask)
        if sample:
            e1.record(stream)
            self.fwd_pairs.append((e0, e1))
        if self.train:
            stream.wait_stream(self.zero_stream)
            if sample:
                b0, b1 = self._events()
                b0.record(stream)
            _, gT, _, _ = ops.tps_warp_bwd(self.U, self.coord, T, (H, W), self.g_out, None, None, need_grad_U=True, want_grid_grad=True, grad_U_out=gU)
            if sample:
                b1.record(stream)
                self.bwd_pairs.append((b0, b1))
            ops.tps_solve_bwd(self.coord, gT)


def median_ms(pairs):          # median of the per-launch samples
    ts = sorted(a.elapsed_time(b) for a, b in pairs)
    return ts[len(ts) // 2] if ts else float('nan')


def timed_loop(torch, dist, world, dev, p, n_passes, n_samples=32, clock_index=None):
    """n_passes back-to-back passes between two events (barrier + synchronize on both sides); per-launch event pairs on
    ~n_samples of them (an event pair between back-to-back kernels costs ~15 us of bubble, charged to the total).
    Returns (elapsed ms, max over ranks; clock summary or None)."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
    every = max(1, n_passes // n_samples) if n_samples > 0 else 0
    p.fwd_pairs.clear()
    p.bwd_pairs.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(clock_index) if clock_index is not None else None
    if clk is not None:
        clk.__enter__()
    barrier()
    t0.record()
    for i in range(n_passes):
        p.run(sample=(every > 0 and i % every == every // 2))
    t1.record()
    barrier()
    if clk is not None:
        clk.__exit__()
    ms = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, (clk.summary() if clk is not None else None)


def calibrate_passes(torch, dist, world, dev, p, steps, min_s):
    """passes per step so that `steps` steps last >= min_s (same count on every rank: max over ranks of the pass time)."""
    n = 4
    while True:
        ms, _ = timed_loop(torch, dist, world, dev, p, n, n_samples=0)
        if ms >= 20.0 or n >= 4096:
            break
        n *= 4
    per_pass_s = ms * 1e-3 / n
    return max(1, int(-(-min_s // (per_pass_s * max(steps, 1)))))


def roofline_of(p, peak, peak_src, clocks, traffic=None, traffic_src=None):
    """Roofline record of pass `p` from its per-launch event samples."""
    wl = p.wl
    kern_ms = median_ms(p.fwd_pairs)
    name, bpp = p.kernel_name, p.fwd_bpp
    achieved = p.pix * bpp / (kern_ms * 1e-3) / 1e9
    extra = {}
    if p.kind == 'tps_train':
        # the dominant kernel of the training shape is the backward: 56 B/px (grad_out 12 + source 12 + grad_image zero
        # fill 12 + grad_image accumulate 12 + grad_x,y 8)
        extra['forward_kernel'] = {'kernel': name, 'kernel_ms': kern_ms, 'achieved': achieved, 'algorithmic_bytes_per_px': bpp,
                                   'frac': achieved / peak}
        name, kern_ms, bpp = 'warp_bwd_tile_kernel<TMODE_TPS> (warp_bwd_tile.cu)', median_ms(p.bwd_pairs), 56
        achieved = p.pix * bpp / (kern_ms * 1e-3) / 1e9
    if p.kind in ('tps', 'tps_train'):
        # second roofline of the EXACT fused TPS kernel: one MUFU.LG2 per pixel and control point on the XU pipe (16 lanes
        # per SM and clock).  The tile-node evaluation needs 1/8 of them, so this is no longer a bound of the
        # default path; it is reported to show by how much the direct evaluation would be limited.
        mhz = (clocks or {}).get('sm_mhz') or 1965
        n_logs = wl['mesh'] ** 2 * (2 if p.kind == 'tps_train' else 1)
        xu_gpix = 148 * 16 * mhz * 1e6 / n_logs / 1e9
        extra['xu_bound_of_direct_evaluation'] = {'unit': 'Gpix/s', 'peak': xu_gpix, 'achieved': p.pix / (kern_ms * 1e-3) / 1e9,
                                                  'frac': p.pix / (kern_ms * 1e-3) / 1e9 / xu_gpix, 'logs_per_px': n_logs, 'sm_mhz': mhz}
    rec = {'bound': 'hbm', 'kernel': name, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
           'traffic': traffic, 'peak_source': peak_src, 'algorithmic_bytes_per_px': bpp, 'kernel_ms': kern_ms,
           'launches_sampled': len(p.bwd_pairs if p.kind == 'tps_train' else p.fwd_pairs)}
    if traffic_src:
        rec['traffic_source'] = traffic_src
    rec.update(extra)
    return rec


def sustained_copy(torch, dev, achieved_gbs, seconds=0.5):
    """GB/s (read + write) of b.copy_(a) over 1 GiB held for `seconds` -- the copy MEASURED_PEAKS.json times in burst."""
    n = 1 << 29
    a = torch.empty(n, dtype=torch.bfloat16, device=dev)
    b = torch.empty_like(a)
    for _ in range(3):
        b.copy_(a)
    torch.cuda.synchronize(dev)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 16
    t0.record()
    for _ in range(reps):
        b.copy_(a)
    t1.record()
    torch.cuda.synchronize(dev)
    reps = max(reps, int(seconds / (t0.elapsed_time(t1) * 1e-3 / reps)))
    t0.record()
    for _ in range(reps):
        b.copy_(a)
    t1.record()
    torch.cuda.synchronize(dev)
    gbs = 2.0 * n * 2 * reps / (t0.elapsed_time(t1) * 1e-3) / 1e9
    del a, b
    torch.cuda.empty_cache()
    return {'sustained_copy_gbs': gbs, 'frac_of_sustained_copy': achieved_gbs / gbs,
            'sustained_copy': 'torch b.copy_(a), 1 Gi bf16 elements, %d back-to-back launches (%.2f s), CUDA events' % (reps, t0.elapsed_time(t1) * 1e-3)}


def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from coupe.dvsg_b200 import _lib, ops
    from coupe.dvsg_b200.sharding import frame_shard

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- this path has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()
    warm = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- main line: weak scaling, every rank its own batch -------------------------------------------------------
    B, H, W = wl['B'], wl['H'], wl['W']
    p = Pass(wl, B, dev, args.seed + rank, lib)
    for _ in range(warm):
        p.run()
    barrier()
    passes = calibrate_passes(torch, dist, world, dev, p, args.steps, MIN_TIMED_S)
    l0 = _lib.launch_count()
    elapsed_ms, clocks = timed_loop(torch, dist, world, dev, p, passes * args.steps, clock_index=local)
    launches = _lib.launch_count() - l0
    peak, peak_src = load_peaks()
    traffic, traffic_src, tj = None, None, {}
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            tj = json.load(fh)
            traffic, traffic_src = tj.get(args.workload), tj.get('_source')
    except Exception:
        pass
    main_roof = roofline_of(p, peak, peak_src, clocks, traffic, traffic_src) if rank == 0 else None
    if rank == 0 and main_roof is not None and not args.no_extras:
        # context for a SUSTAINED number: the timed region lasts >= 0.5 s, long enough for the board's power cap to pull the SM
        # clock down (clocks.reasons), while `peak` is a burst copy (best of 10).  The same plain copy, sustained for 0.5 s:
        main_roof.update(sustained_copy(torch, dev, main_roof['achieved']))
    value = world * p.pix * passes * args.steps / (elapsed_ms * 1e-3) / 1e6

    # ---- e2e: host buffers through the C-ABI host pipeline --------------------------------------------------------
    e2e = None
    e2e_u8 = None
    kind = wl['kind']
    if kind == 'tps' and not args.no_e2e:
        Be = min(B, args.e2e_frames)
        pipe = ops.HostPipeline(H, W, 3, wl['mesh'] ** 2, frames_per_chunk=max(1, min(args.e2e_chunk, Be)), n_slots=args.e2e_slots, device=local)
        U_h = torch.empty((Be, H, W, 3), dtype=torch.float32).pin_memory()
        U_h.copy_(p.U[:Be])
        out_h = torch.empty_like(U_h).pin_memory()
        mesh_h, vec_h = p.mesh.cpu().contiguous(), p.vec[:Be].cpu().contiguous()

        def time_pipe(fn):
            """(streaming Mpix/s, blocking Mpix/s): e2e_steps batches submitted back to back with one wait at the end (a clip is a
            stream of batches: the upload of batch k+1 overlaps the download of batch k), and one blocking call per batch."""
            res = []
            for streaming in (True, False):
                pipe.set_async(streaming)
                for _ in range(2):
                    fn()
                pipe.sync()
                barrier()
                ts = time.perf_counter()
                for _ in range(args.e2e_steps):
                    fn()   # every call uploads its batch from pinned host memory and downloads the warped frames
                pipe.sync()
                te = time.perf_counter() - ts
                if world > 1:
                    t = torch.tensor([te], device=dev, dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    te = float(t.item())
                res.append(world * Be * H * W * args.e2e_steps / te / 1e6)
            pipe.set_async(False)
            return res
        v, vb = time_pipe(lambda: pipe.thin_plate_spline(U_h, mesh_h, vec_h, out_h))
        e2e = {'value': v, 'unit': 'Mpix/s', 'blocking_value': vb,
               'h2d_bytes_per_step': int(U_h.numel() * 4 + vec_h.numel() * 4 + mesh_h.numel() * 4),
               'd2h_bytes_per_step': int(out_h.numel() * 4), 'frames_per_step': Be, 'pcie_gbs': v * 24e-3,
               'api': 'dvsg_host_tps_warp (coupe.dvsg_b200.ops.HostPipeline.thin_plate_spline), pinned host buffers; value = batches streamed '
                      'through the pipeline (dvsg_host_pipeline_set_async + one sync per %d batches), blocking_value = one blocking call per batch' % args.e2e_steps}
        # same call with uint8 frames on the host side (N4: ingest u/255 and egress uint8(x*255) on the device)
        U8_h = (U_h * 255.0).to(torch.uint8).pin_memory()
        out8_h = torch.empty_like(U8_h).pin_memory()
        v8, v8b = time_pipe(lambda: pipe.thin_plate_spline_u8(U8_h, mesh_h, vec_h, out8_h))
        e2e_u8 = {'value': v8, 'unit': 'Mpix/s', 'blocking_value': v8b,
                  'h2d_bytes_per_step': int(U8_h.numel() + vec_h.numel() * 4 + mesh_h.numel() * 4),
                  'd2h_bytes_per_step': int(out8_h.numel()), 'frames_per_step': Be, 'pcie_gbs': v8 * 6e-3,
                  'api': 'dvsg_host_tps_warp_u8 (HostPipeline.thin_plate_spline_u8): uint8 BGR frames in and out, eval.py:76-81,112-113 on the device'}
        pipe.close()
        del U_h, out_h, U8_h, out8_h
    if kind == 'flow' and not args.no_e2e:
        Be = min(B, args.e2e_frames)
        pipe = ops.HostPipeline(H, W, 3, 16, frames_per_chunk=max(1, min(args.e2e_chunk, Be)), n_slots=args.e2e_slots, device=local)
        im_h = torch.empty((Be, H, W, 3), dtype=torch.float32).pin_memory()
        im_h.copy_(p.U[:Be])
        fl_h = torch.empty((Be, H, W, 2), dtype=torch.float32).pin_memory()
        fl_h.copy_(p.flow[:Be])
        out_h = torch.empty_like(im_h).pin_memory()
        res = []
        for streaming in (True, False):
            pipe.set_async(streaming)
            for _ in range(2):
                pipe.tf_warp(im_h, fl_h, out_h)
            pipe.sync()
            barrier()
            ts = time.perf_counter()
            for _ in range(args.e2e_steps):
                pipe.tf_warp(im_h, fl_h, out_h)
            pipe.sync()
            te = time.perf_counter() - ts
            if world > 1:
                t = torch.tensor([te], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                te = float(t.item())
            res.append(world * Be * H * W * args.e2e_steps / te / 1e6)
        e2e = {'value': res[0], 'unit': 'Mpix/s', 'blocking_value': res[1], 'h2d_bytes_per_step': int(im_h.numel() * 4 + fl_h.numel() * 4),
               'd2h_bytes_per_step': int(out_h.numel() * 4), 'frames_per_step': Be, 'pcie_gbs': res[0] * 32e-3,
               'api': 'dvsg_host_flow_warp (HostPipeline.tf_warp), pinned host buffers; value = batches streamed, blocking_value = one blocking call per batch'}
        pipe.close()
        del im_h, fl_h, out_h
    del p
    torch.cuda.empty_cache()

    # ---- the north-star shape in the same run (driver-verifiable): 64 x 1080p, 4x4 mesh ---------------------------
    north = None
    if args.workload == 'cfg2' and not args.no_extras:
        nwl = WORKLOADS['hd1080']
        q = Pass(nwl, nwl['B'], dev, args.seed + rank, lib)
        for _ in range(warm):
            q.run()
        barrier()
        n_q = calibrate_passes(torch, dist, world, dev, q, 1, MIN_TIMED_S)
        ms_q, clk_q = timed_loop(torch, dist, world, dev, q, n_q, clock_index=local)
        if rank == 0:
            north = {'workload': nwl['desc'], 'value': world * q.pix * n_q / (ms_q * 1e-3) / 1e6, 'unit': 'Mpix/s', 'passes': n_q,
                     'timed_s': ms_q * 1e-3, 'ms_per_pass': ms_q / n_q, 'scaling': 'weak',
                     'roofline': roofline_of(q, peak, peak_src, clk_q, tj.get('hd1080') if traffic_src else None, traffic_src), 'clocks': clk_q,
                     'target': '>= 0.70 of HBM on 1 GPU (BASELINE.json north_star)'}
        del q
        torch.cuda.empty_cache()

    # ---- strong scaling: fixed-size jobs cut over the ranks with sharding.frame_shard (no collective) -------------
    strong = []
    if args.workload == 'cfg2' and not args.no_extras:
        # BASELINE configs[3]: 16 x 1080p tf_warp "sharded across 8 B200"
        fwl = WORKLOADS['cfg4']
        a, b = frame_shard(16, rank, world)
        s4 = Pass(fwl, b - a, dev, args.seed + 100 + rank, lib) if b > a else None
        if s4 is not None:
            for _ in range(warm):
                s4.run()
        barrier()
        n4 = 256      # same count on every rank; a 2-frame shard lasts ~25 us
        if s4 is not None:
            ms4, _ = timed_loop(torch, dist, world, dev, s4, n4)
        else:
            ms4, _ = timed_loop(torch, dist, world, dev, _Idle(), n4)
        strong.append({'workload': fwl['desc'], 'job_frames': 16, 'frames_this_rank': b - a, 'passes': n4, 'ms_per_pass': ms4 / n4,
                       'value': 16 * fwl['H'] * fwl['W'] * n4 / (ms4 * 1e-3) / 1e6, 'unit': 'Mpix/s', 'scaling': 'strong'})
        del s4
        torch.cuda.empty_cache()
        # BASELINE configs[4]: one long 4K clip with the 16x16 mesh, frame-sharded; a fixed 512-frame clip streams through a
        # ring of 16 resident frames per rank (the 8192-frame clip is 16 such jobs back to back)
        cwl = WORKLOADS['cfg5']
        clip = int(os.environ.get('DVSG_BENCH_CLIP', '512'))
        a, b = frame_shard(clip, rank, world)
        ring = Pass(cwl, 16, dev, args.seed + 200 + rank, lib)
        tail = (b - a) % 16
        ring_tail = Pass(cwl, tail, dev, args.seed + 300 + rank, lib) if tail else None
        for _ in range(warm):
            ring.run()
            if ring_tail is not None:
                ring_tail.run()
        barrier()

        class _Clip(object):
            fwd_pairs, bwd_pairs = ring.fwd_pairs, ring.bwd_pairs

            @staticmethod
            def run(sample=False):
                for _ in range((b - a) // 16):
                    ring.run(sample)
                    sample = False
                if ring_tail is not None:
                    ring_tail.run()
        ms5, clk5 = timed_loop(torch, dist, world, dev, _Clip(), 2, n_samples=2, clock_index=local)
        rec = {'workload': cwl['desc'], 'job_frames': clip, 'frames_this_rank': b - a, 'passes': 2, 'ms_per_pass': ms5 / 2,
               'value': clip * cwl['H'] * cwl['W'] * 2 / (ms5 * 1e-3) / 1e6, 'unit': 'Mpix/s', 'scaling': 'strong', 'clocks': clk5}
        if rank == 0 and ring.fwd_pairs:
            rec['roofline'] = roofline_of(ring, peak, peak_src, clk5, tj.get('cfg5') if traffic_src else None, traffic_src)
        strong.append(rec)
        del ring, ring_tail
        torch.cuda.empty_cache()

    if rank == 0:
        line = {
            'metric': 'warped Mpix/s', 'value': value, 'unit': 'Mpix/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': warm, 'ms_per_step': elapsed_ms / max(args.steps, 1), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': config_for(wl, world),
            'timing': {'passes_per_step': passes, 'ms_per_pass': elapsed_ms / (passes * args.steps), 'timed_s': elapsed_ms * 1e-3,
                       'clock': 'CUDA events on the launching stream, max over ranks'},
            'roofline': main_roof,
            'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks,
        }
        if e2e_u8 is not None:
            line['e2e_u8'] = e2e_u8
        if north is not None:
            line['north_star'] = north
        if strong:
            line['strong'] = strong
        if not args.no_cpu and world == 1:
            kind_ref, v1, note = reference_impl()
            threads = min(os.cpu_count() or 1, 16)
            threads, frames = cpu_sample_plan(wl, threads, want_frames=min(wl['B'], 2 * threads))
            tcpu, px = tf_run(v1, wl, frames, threads) if v1 is not None else cpu_run(wl, frames, threads)
            line['cpu_baseline'] = {'value': px / tcpu / 1e6, 'unit': 'Mpix/s', 'cores': threads, 'kind': kind_ref,
                                    'sample': '%d frames of the same workload; %s' % (frames, note if v1 is not None else 'NumPy fp32 oracle, one frame per thread; probe: ' + str(note))}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


class _Idle(object):
    """A rank that owns no frame of a strong-scaling job still takes part in the barriers."""
    fwd_pairs, bwd_pairs = [], []

    @staticmethod
    def run(sample=False):
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=sorted(WORKLOADS))
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--e2e-steps', type=int, default=5)
    ap.add_argument('--e2e-frames', type=int, default=64)
    ap.add_argument('--e2e-chunk', type=int, default=8, help='frames per staging chunk of the host pipeline')
    ap.add_argument('--e2e-slots', type=int, default=4, help='device staging slots (streams) of the host pipeline')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the north_star and strong sub-records of the default workload')
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == '__main__':
    main()
