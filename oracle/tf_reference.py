"""TEST / BASELINE INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference files under real TensorFlow when it exists.

BASELINE.md section 4, steps 1-2 / SURVEY.md 8(c)(3): at run time, probe `import tensorflow`; when it is importable AND
the reference tree is reachable (DVSG_REFERENCE, default /root/reference -- it is never copied into this repo), execute
ThinPlateSpline.py / spatial_transformer.py / warp_with_optical_flow.py as they are under `tensorflow.compat.v1` in graph
mode on the CPU, all host cores.  TensorFlow is NOT part of the build image (no network: it cannot be installed), so in
practice this module reports why the probe failed and callers fall back to the NumPy oracle (`kind: "port"`); the code
path exists so that a box that does have TensorFlow pins the oracle against the real kernels (tests/test_tf_reference.py)
and times the real reference (bench.py --impl reference).

Only tests/, bench.py's CPU legs and __graft_entry__.smoke() may import anything under oracle/.
"""
import os
import sys
import types

REF_DIR = os.environ.get('DVSG_REFERENCE', '/root/reference')
FILES = ('ThinPlateSpline.py', 'ThinPlateSpline2.py', 'spatial_transformer.py', 'warp_with_optical_flow.py')


def probe():
    """(tf.compat.v1 module, None) when the unmodified reference can run here, else (None, one-line reason)."""
    try:
        import tensorflow as tf   # noqa: F401
    except Exception as e:       # ModuleNotFoundError in the build image
        return None, 'tensorflow not importable (%s: %s)' % (type(e).__name__, e)
    missing = [f for f in FILES if not os.path.exists(os.path.join(REF_DIR, f))]
    if missing:
        return None, 'tensorflow %s present, but the reference tree is not reachable at %s' % (tf.__version__, REF_DIR)
    v1 = tf.compat.v1
    v1.disable_eager_execution()
    return v1, None


_modules = {}


def load(v1, name):
    """Execute one unmodified reference file with `import tensorflow as tf` resolving to tf.compat.v1."""
    if name in _modules:
        return _modules[name]
    path = os.path.join(REF_DIR, name + '.py')
    saved = sys.modules.get('tensorflow')
    sys.modules['tensorflow'] = v1
    try:
        with open(path, 'r') as fh:
            src = fh.read()
        mod = types.ModuleType('ref_' + name)
        mod.__file__ = path
        exec(compile(src, path, 'exec'), mod.__dict__)
    finally:
        if saved is not None:
            sys.modules['tensorflow'] = saved
    _modules[name] = mod
    return mod


def _session(v1, threads):
    cfg = v1.ConfigProto(intra_op_parallelism_threads=threads, inter_op_parallelism_threads=threads, device_count={'GPU': 0})
    return v1.Session(config=cfg)


class TpsRunner(object):
    """ThinPlateSpline(U, coord, vector, out_size) of the reference as a reusable graph (built once, run many times)."""

    def __init__(self, v1, shape, pn, out_size, threads=None):
        self.v1 = v1
        mod = load(v1, 'ThinPlateSpline')
        self.graph = v1.Graph()
        with self.graph.as_default():
            self.u = v1.placeholder(v1.float32, shape)
            self.c = v1.placeholder(v1.float32, [shape[0], pn, 2])
            self.v = v1.placeholder(v1.float32, [shape[0], pn, 2])
            self.outs = mod.ThinPlateSpline(self.u, self.c, self.v, list(out_size))
            self.sess = _session(v1, threads or (os.cpu_count() or 1))

    def __call__(self, u, coord, vector):
        return self.sess.run(self.outs, {self.u: u, self.c: coord, self.v: vector})


class FlowRunner(object):
    """tf_warp(im, flow, h, w) of the reference as a reusable graph."""

    def __init__(self, v1, shape, threads=None):
        mod = load(v1, 'warp_with_optical_flow')
        self.graph = v1.Graph()
        with self.graph.as_default():
            self.im = v1.placeholder(v1.float32, shape)
            self.fl = v1.placeholder(v1.float32, [shape[0], shape[1], shape[2], 2])
            self.out = mod.tf_warp(self.im, self.fl, shape[1], shape[2])
            self.sess = _session(v1, threads or (os.cpu_count() or 1))

    def __call__(self, im, flow):
        return self.sess.run(self.out, {self.im: im, self.fl: flow})
