"""Times kernel variants with CUDA events (tuning knobs of the library) -- run on the GPU box."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coupe.dvsg_b200 import _lib, ops

lib = _lib.load()
dev = torch.device('cuda', 0)


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) for x, y in evs)
    return ts[len(ts) // 2]


def tps_case(B, H, W, m, amp=0.2):
    torch.manual_seed(0)
    U = torch.rand((B, H, W, 3), device=dev)
    lin = torch.arange(m, device=dev, dtype=torch.float32) * (2.0 / (m - 1)) - 1.0
    mesh = torch.stack(torch.meshgrid(lin, lin, indexing='xy'), dim=-1).reshape(m * m, 2).contiguous()
    coord = mesh.unsqueeze(0).expand(B, -1, -1)
    vec = (torch.rand((B, m * m, 2), device=dev) - 0.5) * amp
    T = ops.tps_solve(coord, coord + vec)
    return U, coord, T


def smooth_flow(B, H, W, amp=8.0, jitter=0.5):
    lat = (torch.rand((B, 2, 9, 16), device=dev) - 0.5) * 2 * amp
    f = torch.nn.functional.interpolate(lat, size=(H, W), mode='bilinear', align_corners=True)
    f = f + (torch.rand((B, 2, H, W), device=dev) - 0.5) * 2 * jitter
    return f.permute(0, 2, 3, 1).contiguous()


def main():
    res = []
    B, H, W = 64, 720, 1280
    px = B * H * W
    U, coord, T = tps_case(B, H, W, 4)

    def flow_call(im, flow, out, flags):
        rc = lib.dvsg_flow_warp_fwd(im.data_ptr(), flow.data_ptr(), out.data_ptr(), im.shape[0], im.shape[1], im.shape[2], 3, flags, 0)
        assert rc == 0, lib.dvsg_last_error()

    for pipe in (1, 0):
        lib.dvsg_set_strip_tuning(148 * 12, pipe)
        for smem in (20, 28, 36):
            lib.dvsg_set_tuning(smem * 1024, 1)
            ms = timeit(lambda: ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=False))
            res.append(('tps720 4x4 strip pipe=%d smem=%dK' % (pipe, smem), ms, px * 24 / ms / 1e6))
        lib.dvsg_set_tuning(28 * 1024, 0)
        ms = timeit(lambda: ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=False))
        res.append(('tps720 4x4 strip pipe=%d nopack' % pipe, ms, px * 24 / ms / 1e6))
        lib.dvsg_set_tuning(28 * 1024, 1)
        for tc in (148 * 4, 148 * 24, 148 * 48):
            lib.dvsg_set_strip_tuning(tc, pipe)
            ms = timeit(lambda: ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=False))
            res.append(('tps720 4x4 strip pipe=%d target_ctas=%d' % (pipe, tc), ms, px * 24 / ms / 1e6))
        lib.dvsg_set_strip_tuning(148 * 12, pipe)
        ms = timeit(lambda: ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=True))
        res.append(('tps720 strip pipe=%d +xy' % pipe, ms, px * 32 / ms / 1e6))
    for name, flags in (('legacy-staged', 2), ('direct', 1)):
        ms = timeit(lambda: ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=False, flags=flags))
        res.append(('tps720 4x4 %s' % name, ms, px * 24 / ms / 1e6))
    U5, c5, T5 = tps_case(B, H, W, 5)
    for pipe in (1, 0):
        lib.dvsg_set_strip_tuning(148 * 12, pipe)
        ms = timeit(lambda: ops.tps_warp_fwd(U5, c5, T5, (H, W), want_grid=False))
        res.append(('tps720 5x5 strip pipe=%d' % pipe, ms, px * 24 / ms / 1e6))
    del U, U5
    Uc, cc, Tc = tps_case(4, 2160, 3840, 16)
    for pipe in (1, 0):
        lib.dvsg_set_strip_tuning(148 * 12, pipe)
        ms = timeit(lambda: ops.tps_warp_fwd(Uc, cc, Tc, (2160, 3840), want_grid=False), n=5, warm=1)
        res.append(('tps4k 16x16 strip B=4 pipe=%d' % pipe, ms, 4 * 2160 * 3840 * 24 / ms / 1e6))
    del Uc
    Us, cs, Ts = tps_case(32, 288, 512, 4)
    for pipe in (1, 0):
        lib.dvsg_set_strip_tuning(148 * 12, pipe)
        ms = timeit(lambda: ops.tps_warp_fwd(Us, cs, Ts, (288, 512), want_grid=True))
        res.append(('tps 288x512 B=32 +xy (L2-resident) pipe=%d' % pipe, ms, 32 * 288 * 512 * 32 / ms / 1e6))
    # flow
    im = torch.rand((16, 1080, 1920, 3), device=dev)
    out = torch.empty_like(im)
    pxf = 16 * 1080 * 1920
    for fname, flow in (('smooth', smooth_flow(16, 1080, 1920)), ('random+-8', (torch.rand((16, 1080, 1920, 2), device=dev) - 0.5) * 16)):
        for pipe in (1, 0):
            lib.dvsg_set_strip_tuning(148 * 12, pipe)
            for smem in (20, 28, 40):
                lib.dvsg_set_tuning(smem * 1024, 1)
                ms = timeit(lambda: flow_call(im, flow, out, 0))
                res.append(('flow1080 %s strip pipe=%d smem=%dK' % (fname, pipe, smem), ms, pxf * 32 / ms / 1e6))
        lib.dvsg_set_tuning(28 * 1024, 1)
        for name, flags in (('legacy-staged', 2), ('direct', 1)):
            ms = timeit(lambda: flow_call(im, flow, out, flags))
            res.append(('flow1080 %s %s' % (fname, name), ms, pxf * 32 / ms / 1e6))
    ms = timeit(lambda: out.copy_(im))
    res.append(('torch copy 398MB', ms, pxf * 24 / ms / 1e6))
    for name, ms, gbs in res:
        print('%-52s %8.3f ms  %8.1f GB/s  %5.1f%%' % (name, ms, gbs, 100 * gbs / 6548.2))


if __name__ == "__main__":
    main()
