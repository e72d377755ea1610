"""experiment: tiles per CTA of the forward tile kernel, old rule (DVSG_FWD_OLDSEG) vs tile_pick_seg_len vs fixed values"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coupe.dvsg_b200 import _lib, ops
exec(open(os.path.join(os.path.dirname(__file__), 'sweep.py')).read().split("def main():")[0])
def variants(run, tag, fixed):
    for name, env in [('old', {'DVSG_FWD_OLDSEG': '1'}), ('new', {})] + [('len %s' % f, {'DVSG_FWD_SEGLEN': str(f)}) for f in fixed]:
        for k in ('DVSG_FWD_OLDSEG', 'DVSG_FWD_SEGLEN'): os.environ.pop(k, None)
        os.environ.update(env)
        ms = min(timeit(run) for _ in range(3))
        print('%-34s %-8s %.4f ms' % (tag, name, ms), flush=True)
    for k in ('DVSG_FWD_OLDSEG', 'DVSG_FWD_SEGLEN'): os.environ.pop(k, None)
for B, H, W, grid in ((32, 288, 512, True), (1, 288, 512, False), (4, 288, 512, False), (16, 720, 1280, False), (16, 1080, 1920, False), (64, 720, 1280, False)):
    U, coord, T = tps_case(B, H, W, 4, 0.2)
    variants(lambda: ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=grid), 'tps %dx%dx%d grid=%s' % (B, H, W, grid), [4, 8, 16, 20])
    del U
im = torch.rand((16, 1080, 1920, 3), device=dev); flow = smooth_flow(16, 1080, 1920); out = torch.empty_like(im)
s = torch.cuda.current_stream().cuda_stream
def fl():
    assert lib.dvsg_flow_warp_fwd(im.data_ptr(), flow.data_ptr(), out.data_ptr(), 16, 1080, 1920, 3, 0, 0) == 0
variants(fl, 'flow 16x1080p', [15, 20, 30, 60])
