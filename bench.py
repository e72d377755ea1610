#!/usr/bin/env python
"""bench.py -- throughput of the DVSG frame-warping hot path on B200 (one JSON line).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic frames: TPS coefficient
solve + fused TPS-grid/bilinear warp (ThinPlateSpline forward) for the TPS workloads, the
dense flow warp for cfg4.  Default workload = BASELINE.json configs[1]:
batch 64 synthetic 720p RGB frames, 4x4 control mesh, 1 B200.  Frames shard across ranks
with no collective (weak scaling: every rank processes its own full batch); the only
torch.distributed use is the barrier and the max-over-ranks of the elapsed time.

`value`      warped Mpix/s, inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`        same metric through the host-buffer entry point (pinned host frames in, warped
             frames back in host memory; H2D and D2H inside the timed region).
`roofline`   dominant kernel (fused warp) against the measured HBM copy bandwidth.
`cpu_baseline` the NumPy oracle (port of the reference's TF graph) on a bounded sample.
`--impl reference` times that oracle alone on all host threads (rank 0 only).
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
    'cfg4': dict(kind='flow', B=16, H=1080, W=1920, mesh=0, bpp=32, desc='tf_warp dense flow warp, batch 16 1080p + flow (BASELINE configs[3])'),
    'cfg5': dict(kind='tps', B=16, H=2160, W=3840, mesh=16, bpp=24, desc='ThinPlateSpline fwd, 4K frames, 16x16 mesh, ring of 16 resident frames (BASELINE configs[4])'),
}


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


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
# CPU side: the oracle (a port of the reference's TF graph), frames spread over host threads
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


def cpu_sample_plan(wl, threads):
    """Frames per CPU sample so that one sample is a few seconds of work per thread."""
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
    # ~1 s per 720p frame and thread: several frames per thread so that a sample is 10-30 s of CPU work
    per_thread = max(1, min(8, int(round(8 * 921600 / px))))
    if wl['mesh'] > 8:
        per_thread = 1
    return threads, threads * per_thread


def run_reference(args, wl):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    threads, frames = cpu_sample_plan(wl, min(threads, 64))
    # bounded run: size the per-step sample so that warmup + steps finish in about DVSG_REF_BUDGET_S seconds whatever
    # --steps / --warmup the caller chose (one calibration step of one frame per thread measures the host first)
    budget = float(os.environ.get('DVSG_REF_BUDGET_S', '150'))
    t_cal, _ = cpu_run(wl, threads, threads, seed=1)
    per_thread = int(budget / max((args.steps + args.warmup) * t_cal, 1e-9))
    frames = threads * max(1, min(frames // threads, per_thread))
    for _ in range(args.warmup):
        cpu_run(wl, frames, threads, seed=1)
    total_t, total_px = 0.0, 0
    for s in range(args.steps):
        t, px = cpu_run(wl, frames, threads, seed=2 + s)
        total_t += t
        total_px += px
    value = total_px / total_t / 1e6
    line = {
        'impl': 'reference', 'metric': 'warped Mpix/s', 'value': value, 'unit': 'Mpix/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total_t / max(args.steps, 1),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': wl['desc'], 'frames_per_step': frames, 'timing': 'host wall clock (CPU arm)'},
        'cpu_baseline': {'value': value, 'unit': 'Mpix/s', 'cores': threads, 'kind': 'port',
                         'sample': '%d frames per step, one frame per thread, NumPy fp32 oracle (op-for-op port of the '
                                   'reference TF graph; TensorFlow is not installable here)' % frames},
        'e2e': {'value': value, 'unit': 'Mpix/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from coupe.dvsg_b200 import _lib, ops

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

    B, H, W = wl['B'], wl['H'], wl['W']
    g = torch.Generator(device=dev)
    g.manual_seed(args.seed + rank)
    U = torch.rand((B, H, W, 3), device=dev, generator=g)
    pix_per_step = B * H * W
    kind = wl['kind']
    ev_pairs = []

    if kind == 'flow':
        lat = (torch.rand((B, 2, 9, 16), device=dev, generator=g) - 0.5) * 16.0
        flow = torch.nn.functional.interpolate(lat, size=(H, W), mode='bilinear', align_corners=True)
        flow = (flow + (torch.rand((B, 2, H, W), device=dev, generator=g) - 0.5)).permute(0, 2, 3, 1).contiguous()
        out = torch.empty_like(U)
        stream = torch.cuda.current_stream(dev)

        def step(sample=False):
            if sample:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            rc = lib.dvsg_flow_warp_fwd(U.data_ptr(), flow.data_ptr(), out.data_ptr(), B, H, W, 3, 0, stream.cuda_stream)
            if sample:
                e1.record(stream)
                ev_pairs.append((e0, e1))
            _lib.check(rc, 'dvsg_flow_warp_fwd')
        kernel_name = 'warp_fwd_tile_kernel<TMODE_FLOW> (warp_fwd_tile.cu)'
        launches_per_step = 1
    else:
        m = wl['mesh']
        lin = torch.arange(m, device=dev, dtype=torch.float32) * (2.0 / (m - 1)) - 1.0
        mesh = torch.stack(torch.meshgrid(lin, lin, indexing='xy'), dim=-1).reshape(m * m, 2).contiguous()
        coord = mesh.unsqueeze(0).expand(B, -1, -1)     # one mesh shared by the batch (model.py:68)
        vec = (torch.rand((B, m * m, 2), device=dev, generator=g) - 0.5) * 0.2
        stream = torch.cuda.current_stream(dev)
        train = kind == 'tps_train'
        g_out = torch.rand((B, H, W, 3), device=dev, generator=g) if train else None

        bwd_pairs = []
        # training step: grad_image buffers are preallocated and zero-filled on a side stream while the solve and the
        # forward kernel run (the fill is part of the step and of the 56 B/px; it only leaves the critical path)
        gU_bufs = [torch.empty_like(U) for _ in range(2)] if train else None
        zero_stream = torch.cuda.Stream(dev) if train else None
        step_no = [0]

        online = ops.OnlineWarper(mesh, B, H, W) if (B == 1 and not train) else None     # cfg1: the per-frame loop of eval.py

        def step(sample=False):
            if online is not None:
                if sample:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                res = online.warp(U, vec)
                if sample:
                    e1.record(stream)
                    ev_pairs.append((e0, e1))
                return res
            if train:
                gU = gU_bufs[step_no[0] & 1]
                step_no[0] += 1
                zero_stream.wait_stream(stream)      # the backward that last accumulated into this buffer has been issued
                with torch.cuda.stream(zero_stream):
                    gU.zero_()
            T = ops.tps_solve(coord, vec, offsets=True)      # target = coord + vec is formed inside the prepared solve
            if sample:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            res = ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=train)
            if sample:
                e1.record(stream)
                ev_pairs.append((e0, e1))
            if train:
                stream.wait_stream(zero_stream)
                if sample:
                    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    b0.record(stream)
                _, gT, _, _ = ops.tps_warp_bwd(U, coord, T, (H, W), g_out, None, None, need_grad_U=True, want_grid_grad=True, grad_U_out=gU)
                if sample:
                    b1.record(stream)
                    bwd_pairs.append((b0, b1))
                ops.tps_solve_bwd(coord, gT)
            return res
        kernel_name = 'warp_fwd_tile_kernel<TMODE_TPS> (warp_fwd_tile.cu)'
        launches_per_step = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ev_pairs.clear()
    if kind == 'tps_train':
        bwd_pairs.clear()
    l0 = _lib.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        t0.record()
        # per-launch CUDA events on every 8th step only: an event pair between back-to-back kernels costs
        # ~15 us of pipeline bubble, which would otherwise be charged to `value`
        for i in range(args.steps):
            step(sample=(i % 8 == 0))
        t1.record()
        barrier()
    elapsed_ms = t0.elapsed_time(t1)
    launches = _lib.launch_count() - l0
    def median_ms(pairs):          # median of the per-launch samples: the first sampled step runs right after the barrier
        ts = sorted(a.elapsed_time(b) for a, b in pairs)
        return ts[len(ts) // 2] if ts else float('nan')
    kern_ms = median_ms(ev_pairs)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())

    # ---- e2e: host buffers through the C-ABI host pipeline (TPS workloads with small meshes) ----
    e2e = None
    e2e_u8 = None
    if kind == 'tps' and wl['mesh'] ** 2 + 3 <= 32 and not args.no_e2e:
        Be = min(B, args.e2e_frames)
        pipe = ops.HostPipeline(H, W, 3, wl['mesh'] ** 2, frames_per_chunk=max(1, min(args.e2e_chunk, Be)), n_slots=args.e2e_slots, device=local)
        U_h = torch.empty((Be, H, W, 3), dtype=torch.float32).pin_memory()
        U_h.copy_(U[:Be])
        out_h = torch.empty_like(U_h).pin_memory()
        mesh_h, vec_h = mesh.cpu().contiguous(), vec[:Be].cpu().contiguous()
        for _ in range(2):
            pipe.thin_plate_spline(U_h, mesh_h, vec_h, out_h)
        barrier()
        ts = time.perf_counter()
        for _ in range(args.e2e_steps):
            pipe.thin_plate_spline(U_h, mesh_h, vec_h, out_h)   # blocking: returns when out_h is complete
        te = time.perf_counter() - ts
        if world > 1:
            t = torch.tensor([te], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        e2e = {'value': world * Be * H * W * args.e2e_steps / te / 1e6, 'unit': 'Mpix/s',
               'h2d_bytes_per_step': int(U_h.numel() * 4 + vec_h.numel() * 4 + mesh_h.numel() * 4),
               'd2h_bytes_per_step': int(out_h.numel() * 4), 'frames_per_step': Be,
               'api': 'dvsg_host_tps_warp (coupe.dvsg_b200.ops.HostPipeline.thin_plate_spline), pinned host buffers'}
        # same call with uint8 frames on the host side (N4: ingest u/255 and egress uint8(x*255) on the device)
        U8_h = (U_h * 255.0).to(torch.uint8).pin_memory()
        out8_h = torch.empty_like(U8_h).pin_memory()
        for _ in range(2):
            pipe.thin_plate_spline_u8(U8_h, mesh_h, vec_h, out8_h)
        barrier()
        ts = time.perf_counter()
        for _ in range(args.e2e_steps):
            pipe.thin_plate_spline_u8(U8_h, mesh_h, vec_h, out8_h)
        te8 = time.perf_counter() - ts
        if world > 1:
            t = torch.tensor([te8], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te8 = float(t.item())
        e2e_u8 = {'value': world * Be * H * W * args.e2e_steps / te8 / 1e6, 'unit': 'Mpix/s',
                  'h2d_bytes_per_step': int(U8_h.numel() + vec_h.numel() * 4 + mesh_h.numel() * 4),
                  'd2h_bytes_per_step': int(out8_h.numel()), 'frames_per_step': Be,
                  'api': 'dvsg_host_tps_warp_u8 (HostPipeline.thin_plate_spline_u8): uint8 BGR frames in and out, eval.py:76-81,112-113 on the device'}
        pipe.close()

    if rank == 0:
        peak, peak_src = load_peaks()
        value = world * pix_per_step * args.steps / (elapsed_ms * 1e-3) / 1e6
        fwd_bpp = 32 if kind in ('flow', 'tps_train') else 24
        achieved = pix_per_step * fwd_bpp / (kern_ms * 1e-3) / 1e9
        extra = {}
        if kind == 'tps_train':
            # the dominant kernel of the training shape is the backward: 56 B/px (grad_out 12 + source 12 +
            # grad_image zero fill 12 + grad_image accumulate 12 + grad_x,y 8)
            bwd_ms = median_ms(bwd_pairs)
            extra = {'forward_kernel': {'kernel': kernel_name, 'kernel_ms': kern_ms, 'achieved': achieved, 'algorithmic_bytes_per_px': 32}}
            kernel_name, kern_ms, fwd_bpp = 'warp_bwd_tile_kernel<TMODE_TPS> (warp_bwd_tile.cu)', bwd_ms, 56
            achieved = pix_per_step * fwd_bpp / (kern_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
                traffic = json.load(fh).get(args.workload)
        except Exception:
            pass
        clocks = clk.summary()
        if kind in ('tps', 'tps_train'):
            # second roofline of the fused TPS kernel: one MUFU.LG2 per pixel and control point on the XU pipe
            # (16 lanes per SM and clock); it binds for large meshes (cfg5), see DESIGN.md
            mhz = (clocks or {}).get('sm_mhz') or 1965
            n_logs = wl['mesh'] ** 2 * (2 if kind == 'tps_train' else 1)
            xu_gpix = 148 * 16 * mhz * 1e6 / n_logs / 1e9
            extra['xu_bound'] = {'unit': 'Gpix/s', 'peak': xu_gpix, 'achieved': pix_per_step / (kern_ms * 1e-3) / 1e9,
                                 'frac': pix_per_step / (kern_ms * 1e-3) / 1e9 / xu_gpix, 'logs_per_px': n_logs, 'sm_mhz': mhz}
        line = {
            'metric': 'warped Mpix/s', 'value': value, 'unit': 'Mpix/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': elapsed_ms / max(args.steps, 1), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': wl['desc'], 'frames_per_gpu': B, 'height': H, 'width': W, 'channels': 3,
                       'mesh': wl['mesh'], 'parallelism': 'frame-sharded x%d, no collective' % world,
                       'l2': 'inputs+outputs per step = %.0f MB > 126 MB L2 (no flush needed)' % (pix_per_step * fwd_bpp / 1e6)
                       if pix_per_step * fwd_bpp > 2.5e8 else 'working set %.0f MB fits L2: HBM fraction is an upper bound' % (pix_per_step * fwd_bpp / 1e6)},
            'roofline': {'bound': 'hbm', 'kernel': kernel_name, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                         'algorithmic_bytes_per_px': fwd_bpp, 'kernel_ms': kern_ms, **extra},
            'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks,
        }
        if e2e_u8 is not None:
            line['e2e_u8'] = e2e_u8
        if not args.no_cpu and world == 1:
            threads = min(os.cpu_count() or 1, 16)
            threads, frames = cpu_sample_plan(wl, threads)
            tcpu, px = cpu_run(wl, frames, threads)
            line['cpu_baseline'] = {'value': px / tcpu / 1e6, 'unit': 'Mpix/s', 'cores': threads, 'kind': 'port',
                                    'sample': '%d frames of the same workload, one per thread, NumPy fp32 oracle' % frames}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=sorted(WORKLOADS))
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--e2e-steps', type=int, default=5)
    ap.add_argument('--e2e-frames', type=int, default=64)
    ap.add_argument('--e2e-chunk', type=int, default=4, help='frames per staging chunk of the host pipeline')
    ap.add_argument('--e2e-slots', type=int, default=4, help='device staging slots (streams) of the host pipeline')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == '__main__':
    main()
