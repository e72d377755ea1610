"""Runs only where TensorFlow AND the reference tree exist (neither is true of the build image or of the GPU boxes cut from
it -- the test is then SKIPPED, visibly): the unmodified reference under tensorflow.compat.v1 against the NumPy oracle
and against the committed golden vectors, i.e. the pin of TensorFlow's own kernels (matrix_inverse, LinSpace, matmul
order) that the TF1-shim goldens cannot give."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import dvsg_oracle as O

tensorflow = pytest.importorskip('tensorflow', reason='TensorFlow is not part of this image (no network to install it)')


@pytest.fixture(scope='module')
def v1():
    from oracle import tf_reference
    mod, why = tf_reference.probe()
    if mod is None:
        pytest.skip(why)
    return mod


@pytest.mark.parametrize('name', ['tps_4x4', 'tps_5x5', 'tps_4x4_big'])
def test_reference_under_tensorflow_matches_oracle_and_goldens(v1, name):
    from oracle import tf_reference
    g = load_golden(name)
    oh, ow = (int(v) for v in g['out_size'])
    run = tf_reference.TpsRunner(v1, g['u'].shape, g['coord'].shape[1], (oh, ow))
    out, x, y = run(g['u'], g['coord'], g['second'])
    o_out, o_x, o_y = O.thin_plate_spline(g['u'], g['coord'], g['second'], (oh, ow))
    # TF >= 2 pins the last LinSpace element to `stop`: allow that one row/column of ulps
    assert max(np.abs(x - o_x).max(), np.abs(y - o_y).max()) <= 2e-5
    assert max(np.abs(x - g['x']).max(), np.abs(y - g['y']).max()) <= 2e-5
    assert np.abs(out - o_out).max() <= 1e-4
    np.testing.assert_array_equal(out.reshape(-1, out.shape[-1]), O.tps_interpolate(g['u'], x, y, oh, ow))


@pytest.mark.parametrize('name', ['flow_small', 'flow_large'])
def test_tf_warp_under_tensorflow_is_bit_identical_to_the_oracle(v1, name):
    from oracle import tf_reference
    g = load_golden(name)
    run = tf_reference.FlowRunner(v1, g['im'].shape)
    out = run(g['im'], g['flow'])
    np.testing.assert_array_equal(out, O.tf_warp(g['im'], g['flow'], g['im'].shape[1], g['im'].shape[2]))
    np.testing.assert_array_equal(out, g['out'])
