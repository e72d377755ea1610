"""A/B of the per-pixel fallback of the forward tile kernel: cp.async tile gather (default) against the
one-pixel-at-a-time path (debug bit 8), on inputs where many tiles fit no staging box."""
import sys
import torch
sys.path.insert(0, '.')
from coupe.dvsg_b200 import ops
from tools.sweep import timeit, tps_case, flow_call, rec, dev, lib


def ab(name, fn, px, bpp, check=None):
    outs = []
    for dbg in (8, 0):
        lib.dvsg_set_tile_tuning(dbg, -1, -1)
        ms = timeit(fn)
        rec('%s [%s]' % (name, 'per-pixel' if dbg else 'pair path'), ms, px, bpp)
        outs.append(check().clone() if check else None)
    lib.dvsg_set_tile_tuning(0, -1, -1)
    if check:
        print('    bit-identical:', torch.equal(outs[0], outs[1]))


B, H, W = 64, 720, 1280
res = {}
for amp in (0.2, 0.4, 0.6):
    U, coord, T = tps_case(B, H, W, 4, amp)
    def f():
        res['o'] = ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=False)[0]
    ab('tps720 4x4 amp=%.2f' % amp, f, B * H * W, 24, lambda: res['o'])
    def fm():
        res['o'] = ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=False, want_mask=True)[3]
    if amp == 0.6:
        ab('tps720 4x4 amp=%.2f +mask' % amp, fm, B * H * W, 28, lambda: res['o'])
del U
Uc, cc, Tc = tps_case(4, 2160, 3840, 16)
def f4():
    res['o'] = ops.tps_warp_fwd(Uc, cc, Tc, (2160, 3840), want_grid=False)[0]
ab('tps4k 16x16 B=4', f4, 4 * 2160 * 3840, 24, lambda: res['o'])
del Uc
im = torch.rand((16, 1080, 1920, 3), device=dev)
out = torch.empty_like(im)
pxf = 16 * 1080 * 1920
for fname, flow in (('random+-8', (torch.rand((16, 1080, 1920, 2), device=dev) - 0.5) * 16), ('random+-2', (torch.rand((16, 1080, 1920, 2), device=dev) - 0.5) * 4)):
    ab('flow1080 %s' % fname, lambda: flow_call(im, flow, out, 0), pxf, 32, lambda: out)
x = torch.rand(pxf, device=dev) * 2.2 - 1.1
y = torch.rand(pxf, device=dev) * 2.2 - 1.1
def fb():
    res['o'] = ops.bilinear_interp(im, x, y, (1080, 1920))
ab('bilinear1080 random xy', fb, pxf, 32, lambda: res['o'])
