"""CPU: the C-ABI library loads and exports every symbol include/dvsg_warp.h declares, the
drop-in modules keep the reference's signatures and fail loudly without CUDA, and the
multi-rank host logic (frame sharding + max-over-ranks) works under gloo with world_size 2."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built_lib():
    from coupe.dvsg_b200 import _build, _lib
    _build.build()          # no-op when the in-tree .so is newer than its sources
    return _lib.load()


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'dvsg_warp.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dvsg_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported_and_bound(built_lib):
    from coupe.dvsg_b200 import _lib
    names = header_functions()
    assert len(names) >= 17
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), 'symbol %s declared in dvsg_warp.h is not exported' % n
        assert n in _lib.PROTOTYPES, 'symbol %s has no ctypes prototype' % n
    assert sorted(_lib.PROTOTYPES) == names
    assert built_lib.dvsg_version() >= 100
    assert built_lib.dvsg_last_error() is not None


def test_argument_validation_needs_no_gpu(built_lib):
    # invalid shapes are rejected before anything touches a device
    rc = built_lib.dvsg_tps_warp_fwd(0, 0, 0, 0, 0, 0, 0, 0, 1, -1, 4, 3, 4, 4, 16, 0, 0)
    assert rc == -1 and b'bad shape' in built_lib.dvsg_last_error()
    rc = built_lib.dvsg_tps_solve(0, 0, 0, 0, 2, 2, 0, 0, 0)
    assert rc == -1
    assert built_lib.dvsg_tps_solve_workspace_bytes(4, 16, 32) == 0
    assert built_lib.dvsg_tps_solve_workspace_bytes(4, 256, 0) == 259 * 518 * 8 + 256
    assert built_lib.dvsg_tps_solve_workspace_bytes(4, 256, 512) == 4 * 259 * 518 * 8 + 256


def test_dropin_signatures_match_the_reference():
    from coupe.dvsg_b200.ThinPlateSpline import ThinPlateSpline
    from coupe.dvsg_b200.ThinPlateSpline2 import ThinPlateSpline2
    from coupe.dvsg_b200 import spatial_transformer as st
    from coupe.dvsg_b200.warp_with_optical_flow import tf_warp
    # /root/reference/ThinPlateSpline.py:4, ThinPlateSpline2.py:4, spatial_transformer.py:460,489,496,
    # warp_with_optical_flow.py:96 -- positional names and order
    assert list(inspect.signature(ThinPlateSpline).parameters)[:4] == ['U', 'coord', 'vector', 'out_size']
    assert list(inspect.signature(ThinPlateSpline2).parameters)[:4] == ['U', 'source', 'target', 'out_size']
    assert list(inspect.signature(st._meshgrid).parameters)[:1] == ['out_size']
    assert list(inspect.signature(st._interpolate).parameters) == ['im', 'x', 'y', 'out_size', 'method']
    assert list(inspect.signature(st.bilinear_interp).parameters) == ['im', 'x', 'y', 'out_size']
    assert list(inspect.signature(tf_warp).parameters) == ['im', 'flow', 'out_height', 'out_width']
    assert list(inspect.signature(st.ProjectiveTransformer.transform).parameters) == ['self', 'inp', 'theta']
    assert st.ProjectiveTransformer([4, 4]).param_dim == 8 and st.AffineTransformer([4, 4]).param_dim == 6
    assert st._interpolate(None, None, None, None, 'nearest') is None
    with pytest.raises(NotImplementedError):
        st._interpolate(None, None, None, None, 'bicubic')


def test_cpu_tensors_are_rejected_not_silently_computed(built_lib):
    from coupe.dvsg_b200.ThinPlateSpline import ThinPlateSpline
    from coupe.dvsg_b200.spatial_transformer import bilinear_interp
    from coupe.dvsg_b200.warp_with_optical_flow import tf_warp
    with pytest.raises(ValueError, match='CUDA only'):
        ThinPlateSpline(torch.zeros(1, 4, 4, 3), torch.zeros(1, 16, 2), torch.zeros(1, 16, 2), [4, 4])
    with pytest.raises(ValueError, match='CUDA only'):
        bilinear_interp(torch.zeros(1, 4, 4, 3), torch.zeros(16), torch.zeros(16), [4, 4])
    with pytest.raises(ValueError, match='CUDA only'):
        tf_warp(torch.zeros(1, 4, 4, 3), torch.zeros(1, 4, 4, 2), 4, 4)


def test_missing_library_fails_loudly(monkeypatch):
    from coupe.dvsg_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', os.path.join(ROOT, 'does', 'not', 'exist.so'))
    with pytest.raises(ImportError, match='no CPU or PyTorch fallback'):
        _lib.load()


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'coupe')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text and 'tf1_shim' not in text, f


def test_frame_shard_partitions():
    from coupe.dvsg_b200.sharding import frame_shard
    for n in (0, 1, 7, 16, 8192):
        for world in (1, 2, 4, 8):
            blocks = [frame_shard(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert frame_shard(16, 3, 8) == (6, 8)          # cfg4: 2 frames per GPU
    assert frame_shard(8192, 7, 8) == (7168, 8192)  # cfg5: 1024-frame sub-clips
    with pytest.raises(ValueError):
        frame_shard(4, 2, 2)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from coupe.dvsg_b200.sharding import frame_shard, max_over_ranks
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    a, b = frame_shard(11, rank, world)
    # every rank "processes" its own frames: checksum of frame ids, no data-path collective
    local = float(sum(range(a, b)))
    t = max_over_ranks(10.0 + rank)
    dist.barrier()
    q.put((rank, a, b, local, t))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_timing_reduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 6), (6, 11)]
    assert sum(r[3] for r in res) == sum(range(11))     # host-side gather of per-rank checksums
    assert all(r[4] == 11.0 for r in res)               # max over ranks seen by every rank


def test_tiles_per_cta_cut_of_the_tile_kernels(built_lib):
    """Host logic of the tile kernels (tile_common.cuh: tile_pick_seg_len): a strip of n_tx tiles is cut into CTAs of
    seg_len tiles.  Pure host code, so it is checked here without a device: bounds, and the cuts the measurements in
    profiles/README.md were taken with."""
    f = built_lib.dvsg_debug_seg_len
    fwd, bwd = 148 * 5, 148 * 4
    for strips in (1, 36, 1152, 5760, 100000):
        for n_tx in (1, 3, 4, 5, 16, 40, 60, 120, 1000):
            for slots, cost in ((fwd, 0.75), (fwd, 0.2), (bwd, 0.5), (bwd, 0.05)):
                n = f(strips, n_tx, slots, cost)
                assert 1 <= n <= n_tx
                assert n == n_tx or n >= 4, 'every warp of a CTA keeps at least one tile'
    assert f(32 * 36, 16, fwd, 0.75) == 16      # training shape: one CTA per strip (forward and TPS backward)
    assert f(32 * 36, 16, bwd, 0.5) == 16
    assert f(32 * 36, 16, bwd, 0.05) == 4       # tf_warp backward at the training shape: short CTAs
    assert f(36, 16, fwd, 0.75) == 4            # one 288x512 frame: as many CTAs as there are warps' worth of tiles
    assert f(64 * 90, 40, fwd, 0.75) == 40      # cfg2: enough strips, one CTA each
    assert f(16 * 135, 60, fwd, 0.2) == 20      # cfg4 flow warp: three CTAs per strip
