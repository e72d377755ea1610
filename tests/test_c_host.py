"""The C ABI from a C host: examples/c_host/warp_demo.c (no Python, no torch in the process) compiles against
include/dvsg_warp.h as C, links libdvsg_warp.so, and -- on the GPU -- produces the frames and gradients of the Python
drop-in on the same inputs."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, 'examples', 'c_host', 'warp_demo.c')
CUDA = os.environ.get('CUDA_HOME', '/usr/local/cuda')


def _compile(tmp_path):
    from coupe.dvsg_b200 import _build
    lib = _build.build()
    libdir = os.path.dirname(lib)
    exe = str(tmp_path / 'warp_demo')
    cmd = ['gcc', '-O2', '-Wall', '-Werror', '-std=c99', '-I', os.path.join(ROOT, 'include'), '-I', os.path.join(CUDA, 'include'), SRC, '-o', exe,
           '-L', libdir, '-ldvsg_warp', '-L', os.path.join(CUDA, 'lib64'), '-lcudart', '-lm', '-Wl,-rpath,' + libdir,
           '-Wl,-rpath,' + os.path.join(CUDA, 'lib64')]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


@pytest.mark.skipif(shutil.which('gcc') is None, reason='no C compiler')
def test_c_host_compiles_and_links_against_the_header_as_c(tmp_path):
    exe = _compile(tmp_path)
    assert os.path.getsize(exe) > 0
    undefined = subprocess.run(['nm', '-u', exe], capture_output=True, text=True).stdout
    for sym in ('dvsg_tps_solve', 'dvsg_tps_warp_fwd', 'dvsg_tps_warp_bwd', 'dvsg_last_error'):
        assert sym in undefined                      # resolved from libdvsg_warp.so at run time


def _lcg_stream(n, state):
    out = np.empty(n, np.float32)
    for i in range(n):
        state = (state * 1664525 + 1013904223) & 0xFFFFFFFF
        out[i] = np.float32(state >> 8) * np.float32(1.0 / 16777216.0)
    return out, state


def _checksum(v):
    v = np.asarray(v, np.float64).reshape(-1)
    return float((v * (1 + (np.arange(v.size) % 7))).sum())


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which('gcc') is None, reason='no C compiler')
def test_c_host_reproduces_the_python_dropin(tmp_path):
    import torch
    from coupe.dvsg_b200 import ops
    exe = _compile(tmp_path)
    B, H, W, C, m = 2, 288, 512, 3, 4
    pn = m * m
    res = subprocess.run([exe, str(B)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    got = dict(line.split() for line in res.stdout.strip().splitlines())
    im, state = _lcg_stream(B * H * W * C, 12345)
    k = np.arange(pn)
    mesh = np.stack([np.float32(-1.0) + np.float32(2.0) * (k % m).astype(np.float32) / np.float32(m - 1),
                     np.float32(-1.0) + np.float32(2.0) * (k // m).astype(np.float32) / np.float32(m - 1)], 1).astype(np.float32)
    r, _ = _lcg_stream(B * pn * 2, state)
    tgt = (mesh.reshape(1, -1) + (r.reshape(B, -1) - np.float32(0.5)) * np.float32(0.2)).astype(np.float32).reshape(B, pn, 2)
    U = torch.from_numpy(im.reshape(B, H, W, C)).cuda()
    C_ = torch.from_numpy(mesh).cuda()
    T = ops.tps_solve(C_, torch.from_numpy(tgt).cuda())
    out, _, _, _ = ops.tps_warp_fwd(U, C_, T, (H, W))
    _, gT, _, _ = ops.tps_warp_bwd(U, C_, T, (H, W), U)
    want_out, want_gT = _checksum(out.cpu().numpy()), _checksum(gT.cpu().numpy())
    assert abs(float(got['out_checksum']) - want_out) <= 1e-7 * abs(want_out)       # the same kernels on the same bits
    assert abs(float(got['gradT_checksum']) - want_gT) <= 1e-4 * abs(want_gT)       # atomics: another order of the same sums
    assert int(got['launches']) >= 3
