import torch, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coupe.dvsg_b200 import ops, _lib
dev = torch.device('cuda', 0)
B, H, W, m = 64, 720, 1280, 4
g = torch.Generator(device=dev); g.manual_seed(0)
U = torch.rand((B, H, W, 3), device=dev, generator=g)
lin = torch.arange(m, device=dev, dtype=torch.float32) * (2.0 / (m - 1)) - 1.0
mesh = torch.stack(torch.meshgrid(lin, lin, indexing='xy'), dim=-1).reshape(m * m, 2).contiguous()
coord = mesh.unsqueeze(0).expand(B, -1, -1)
vec = (torch.rand((B, m * m, 2), device=dev, generator=g) - 0.5) * 0.2
def run(name, fn, n=100):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n): fn()
    t1.record(); torch.cuda.synchronize()
    print('%-40s %.1f us/step' % (name, t0.elapsed_time(t1) / n * 1e3))
T0 = ops.tps_solve(coord, coord + vec)
run('warp only', lambda: ops.tps_warp_fwd(U, coord, T0, (H, W), want_grid=False))
run('solve+warp', lambda: ops.tps_warp_fwd(U, coord, ops.tps_solve(coord, coord + vec), (H, W), want_grid=False))
tgt = coord + vec
run('solve(no add)+warp', lambda: ops.tps_warp_fwd(U, coord, ops.tps_solve(coord, tgt), (H, W), want_grid=False))
run('solve only', lambda: ops.tps_solve(coord, tgt))
run('add only', lambda: coord + vec)
out = torch.empty_like(U)
lib = _lib.load()
cbuf, cstride, pn = ops._mesh_args(coord, B, 16)
s = torch.cuda.current_stream().cuda_stream
run('raw warp', lambda: lib.dvsg_tps_warp_fwd(U.data_ptr(), cbuf.data_ptr(), cstride, T0.data_ptr(), out.data_ptr(), 0, 0, 0, B, H, W, 3, H, W, 16, 0, s))
Tb = torch.empty_like(T0)
def raw2():
    lib.dvsg_tps_solve(cbuf.data_ptr(), cstride, tgt.data_ptr(), Tb.data_ptr(), B, 16, 0, 0, s)
    lib.dvsg_tps_warp_fwd(U.data_ptr(), cbuf.data_ptr(), cstride, Tb.data_ptr(), out.data_ptr(), 0, 0, 0, B, H, W, 3, H, W, 16, 0, s)
run('raw solve+warp', raw2)
