import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
exec(open(os.path.join(os.path.dirname(__file__), 'sweep.py')).read().split("def main():")[0])
for (B, H, W) in ((32, 288, 512), (1, 288, 512), (8, 720, 1280), (1, 720, 1280), (1, 1080, 1920)):
    U, coord, T = tps_case(B, H, W, 4, 0.2)
    for tc in (370, 740, 1480, 2960, 5920):
        lib.dvsg_set_tile_tuning(-1, tc, -1)
        ms = timeit(lambda: ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=True))
        print('%dx%dx%d +xy target_ctas=%d  %.4f ms  %.1f Gpix/s' % (B, H, W, tc, ms, B * H * W / ms / 1e6))
    lib.dvsg_set_tile_tuning(-1, 2960, -1)
