"""experiment: host-side cost per call of the C ABI and of the Python wrappers (single 288x512 frame)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
exec(open(os.path.join(os.path.dirname(__file__), 'sweep.py')).read().split("def main():")[0])
B, H, W = 1, 288, 512
U, coord, T = tps_case(B, H, W, 4, 0.2)
out = torch.empty_like(U)
cb = coord[0].contiguous()
s = torch.cuda.current_stream().cuda_stream
target = (coord + 0.01).contiguous()
def raw():
    lib.dvsg_tps_warp_fwd(U.data_ptr(), cb.data_ptr(), 0, T.data_ptr(), out.data_ptr(), None, None, None, B, H, W, 3, H, W, 16, 0, s)
def wrapped():
    ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=False)
def solve():
    ops.tps_solve(coord, target)
def dropin():
    from coupe.dvsg_b200.ThinPlateSpline import ThinPlateSpline
    ThinPlateSpline(U, coord, target - coord, [H, W], return_grid=False)
for name, fn in (('C ABI dvsg_tps_warp_fwd', raw), ('ops.tps_warp_fwd', wrapped), ('ops.tps_solve (prepared)', solve), ('ThinPlateSpline drop-in (solve + warp)', dropin)):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    n = 2000
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print('%-42s host %.1f us/call   incl. GPU drain %.1f us/call' % (name, (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))
