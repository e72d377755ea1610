"""experiment: PCIe ceiling of the box (pinned host memory, simultaneous H2D + D2H) next to the host pipeline"""
import time, torch
n = 64 * 720 * 1280 * 3
h_in = torch.empty(n, dtype=torch.float32).pin_memory(); h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, dtype=torch.float32, device='cuda'); d_out = torch.empty(n, dtype=torch.float32, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both(chunks):
    c = n // chunks
    for i in range(chunks):
        with torch.cuda.stream(s1): d_in[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)
        with torch.cuda.stream(s2): h_out[i * c:(i + 1) * c].copy_(d_out[i * c:(i + 1) * c], non_blocking=True)
    torch.cuda.synchronize()
def one(dirn):
    if dirn == 'h2d': d_in.copy_(h_in, non_blocking=True)
    else: h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
for name, fn in (('H2D alone', lambda: one('h2d')), ('D2H alone', lambda: one('d2h')), ('both, 1 chunk', lambda: both(1)), ('both, 32 chunks', lambda: both(32))):
    fn(); t0 = time.perf_counter()
    for _ in range(5): fn()
    dt = (time.perf_counter() - t0) / 5
    tot = n * 4 * (2 if 'both' in name else 1)
    print('%-16s %.2f ms  %.1f GB/s total' % (name, dt * 1e3, tot / dt / 1e9))
