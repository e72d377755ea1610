"""experiment: tiles per CTA (seg_len) of the backward tile kernel at the training shape and at 720p / 1080p"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coupe.dvsg_b200 import _lib, ops
exec(open(os.path.join(os.path.dirname(__file__), 'sweep.py')).read().split("def main():")[0])
s = torch.cuda.current_stream().cuda_stream
for B, H, W in ((32, 288, 512), (16, 720, 1280), (16, 1080, 1920)):
    U, coord, T = tps_case(B, H, W, 4, 0.2)
    g = torch.rand((B, H, W, 3), device=dev)
    gU = torch.zeros_like(U); gT = torch.empty((B, 2, 19), device=dev); gx = torch.empty(B * H * W, device=dev); gy = torch.empty_like(gx)
    cb = coord[0].contiguous()
    def run():
        rc = lib.dvsg_tps_warp_bwd(U.data_ptr(), cb.data_ptr(), 0, T.data_ptr(), g.data_ptr(), None, None, gU.data_ptr(), gT.data_ptr(), gx.data_ptr(), gy.data_ptr(), B, H, W, 3, H, W, 16, s)
        assert rc == 0
    for sl in sys.argv[1:] or ['', '4', '8', '12', '16', '20', '40', '60']:
        if sl: os.environ['DVSG_BWD_SEGLEN'] = sl
        else: os.environ.pop('DVSG_BWD_SEGLEN', None)
        ms = min(timeit(run) for _ in range(3))
        print('%dx%dx%d seg_len %-8s %.4f ms' % (B, H, W, sl or 'default', ms), flush=True)
    os.environ.pop('DVSG_BWD_SEGLEN', None)
