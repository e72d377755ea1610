import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coupe.dvsg_b200 import ops
from tools.sweep import tps_case
from coupe.dvsg_b200 import _lib
_lib.load().dvsg_set_tile_tuning(int(os.environ.get('DBG', '0')), -1, -1)
B, H, W = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
U, coord, T = tps_case(B, H, W, 4, 0.0 if len(sys.argv) < 5 else float(sys.argv[4]))
out = ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=False)
torch.cuda.synchronize()
print('ok', float(out[0].sum()))
