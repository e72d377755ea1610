"""Build recipe for libdvsg_warp.so (hand-written CUDA for sm_100a, in-tree, no JIT cache)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libdvsg_warp.so')
SOURCES = ['api.cu', 'tps_solve.cu', 'warp_fwd.cu', 'warp_fwd_tile.cu', 'warp_bwd.cu', 'warp_bwd_tile.cu', 'frames_u8.cu', 'losses.cu', 'elastic.cu', 'homography.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '--fmad=true', '-Xcompiler', '-fPIC', '-Xcompiler', '-O2', '--threads', '4']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', '..', 'include', 'dvsg_warp.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None, defines=()):
    """Compile every .cu under csrc/ into one shared library next to this file.  `out` / `defines`: an alternative build
    of the same sources for A/B experiments (loaded with DVSG_LIB=...), never a different implementation."""
    if out is None and not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ['-D' + d for d in defines] + (['-Xptxas', '-v'] if verbose else []) + ['-shared', '-o', out or LIB] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building libdvsg_warp.so')
    if verbose:
        sys.stderr.write(res.stderr)
    return out or LIB


if __name__ == '__main__':
    defs = [a[2:] for a in sys.argv[1:] if a.startswith('-D')]
    outs = [a[2:] for a in sys.argv[1:] if a.startswith('-o')]
    print(build(force=True, verbose='-v' in sys.argv, out=outs[0] if outs else None, defines=defs))
