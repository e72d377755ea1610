"""TEST INFRASTRUCTURE ONLY -- a minimal TensorFlow-1.x API emulation over torch CPU.

Purpose: TensorFlow is not installable in the build container (no network), so the
reference's hot-path files cannot be imported as they are.  This module provides just
enough of the `tensorflow` 1.x module surface for the UNMODIFIED reference files

    /root/reference/ThinPlateSpline.py, ThinPlateSpline2.py,
    /root/reference/spatial_transformer.py (incl. ElasticTransformer), /root/reference/warp_with_optical_flow.py
    (+ the methods masked_MSE, temporal_loss, get_surf_loss of /root/reference/trainer.py)

to execute eagerly (each tf.* call computes immediately on torch CPU fp32/int32 tensors).
`tests/golden/make_golden.py` installs it as `sys.modules['tensorflow']`, imports the
reference files from where they lie, runs them on seeded inputs and stores the results
(and, through torch autograd over the very same op sequence, the gradients TF autodiff
would build) as golden vectors under tests/golden/.

What this pins and what it does not:
  * pinned: the reference's own op sequence, argument wiring, index arithmetic and
    border semantics, exactly as written in its source files;
  * NOT pinned: TensorFlow 1.11's kernels themselves.  Each tf op below is implemented
    from the documented TF 1.x semantics, e.g. LinSpace is
    `start + step*i` in fp32, float->int32 cast truncates, clip_by_value is
    min(max(x, lo), hi), add_n sums in list order, matrix_inverse is a partial-pivot LU
    inverse (LAPACK getrf/getri through torch.linalg.inv).

Nothing in the product path (coupe/dvsg_b200) may import this module.
"""
import builtins
import contextlib
import sys
import types

import numpy as np
import torch

float32 = torch.float32
int32 = torch.int32

_DTYPES = {'float32': torch.float32, 'int32': torch.int32, 'float64': torch.float64,
           'int64': torch.int64, torch.float32: torch.float32, torch.int32: torch.int32,
           torch.float64: torch.float64, torch.int64: torch.int64}

# The default float type of the emulated session.  fp32 is what the reference uses; the
# golden generator flips this to float64 to obtain a high-precision run of the same graph.
_FLOAT = [torch.float32]


def set_float(dtype):
    _FLOAT[0] = dtype


def _dt(d):
    d = _DTYPES[d]
    if d is torch.float32:
        return _FLOAT[0]
    return d


class _Shape(object):
    def __init__(self, dims):
        self._dims = [int(d) for d in dims]

    def as_list(self):
        return list(self._dims)


def _get_shape(self):
    return _Shape(self.shape)


# reference calls `tensor.get_shape().as_list()` (ThinPlateSpline.py:118)
torch.Tensor.get_shape = _get_shape


def _t(x, dtype=None):
    """anything -> torch tensor (python floats become the session float type)."""
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    if isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
        if t.dtype.is_floating_point and dtype is None:
            t = t.to(_FLOAT[0])
        return t if dtype is None else t.to(dtype)
    if isinstance(x, (list, tuple)) and any(isinstance(e, torch.Tensor) for e in x):
        return torch.stack([_t(e) for e in x])
    if dtype is None:
        if isinstance(x, float) or (isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], float)):
            dtype = _FLOAT[0]
        elif isinstance(x, (int, np.integer)):
            dtype = torch.int32
    return torch.tensor(x, dtype=dtype)


def _ints(x):
    """shape-like (list of ints / 0-d tensors, or 1-d int tensor) -> list of python ints."""
    if isinstance(x, torch.Tensor):
        return [int(v) for v in x.reshape(-1).tolist()]
    if isinstance(x, (int, np.integer)):
        return [int(x)]
    return [int(v) for v in x]


def shape(x):
    return [int(d) for d in _t(x).shape]


def cast(x, dtype):
    d = _dt(dtype)
    if isinstance(x, torch.Tensor):
        return x.to(d)          # float -> int truncates toward zero, as tf.cast does
    return torch.tensor(x).to(d)


def ones(shape, dtype='float32'):
    return torch.ones(_ints(shape), dtype=_dt(dtype))


def zeros(shape, dtype='float32'):
    return torch.zeros(_ints(shape), dtype=_dt(dtype))


def ones_like(x):
    return torch.ones_like(_t(x))


def zeros_like(x):
    return torch.zeros_like(_t(x))


def constant(x, dtype=None):
    return _t(x, None if dtype is None else _dt(dtype))


def stack(values, axis=0):
    return torch.stack([_t(v) for v in values], dim=axis)


def concat(values, axis):
    return torch.cat([_t(v) for v in values], dim=axis)


def reshape(x, shp):
    return _t(x).reshape(_ints(shp))


def transpose(x, perm=None):
    x = _t(x)
    return x.permute(*reversed(builtins.range(x.dim()))) if perm is None else x.permute(*perm)


def where(cond, a, b):
    return torch.where(cond, _t(a), _t(b))


def is_inf(x):
    return torch.isinf(_t(x))


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def tile(x, multiples):
    return _t(x).repeat(*_ints(multiples))


def slice(x, begin, size):  # noqa: A001 - mirrors tf.slice
    x = _t(x)
    idx = []
    for d, (b, s) in enumerate(zip(begin, size)):
        idx.append(builtins.slice(b, x.shape[d] if s == -1 else b + s))
    return x[tuple(idx)]


def matmul(a, b):
    return torch.matmul(_t(a), _t(b))


def matrix_inverse(a):
    return torch.linalg.inv(_t(a))


def square(x):
    x = _t(x)
    return x * x


def log(x):
    return torch.log(_t(x))


def floor(x):
    return torch.floor(_t(x))


def minimum(a, b):
    a = _t(a)
    return torch.minimum(a, _t(b, a.dtype) if not isinstance(b, torch.Tensor) else b.to(a.dtype))


def maximum(a, b):
    a = _t(a)
    return torch.maximum(a, _t(b, a.dtype) if not isinstance(b, torch.Tensor) else b.to(a.dtype))


def clip_by_value(x, lo, hi):
    # TF: minimum(maximum(x, lo), hi); gradient passes where lo <= x <= hi (inclusive).
    x = _t(x)
    lo = lo.to(x.dtype) if isinstance(lo, torch.Tensor) else torch.tensor(lo, dtype=x.dtype)
    hi = hi.to(x.dtype) if isinstance(hi, torch.Tensor) else torch.tensor(hi, dtype=x.dtype)
    return torch.minimum(torch.maximum(x, lo), hi)


def reduce_sum(x, axis=None):
    x = _t(x)
    return x.sum() if axis is None else x.sum(dim=axis)


def reduce_mean(x, axis=None, name=None):
    x = _t(x)
    return x.mean() if axis is None else x.mean(dim=axis)


def squared_difference(a, b):
    d = _t(a) - _t(b)
    return d * d


def batch_gather(params, indices):
    # params [B, N], indices [B, P] -> params[b, indices[b, p]]   (trainer.py:379-380)
    return torch.gather(_t(params), 1, _t(indices).to(torch.int64))


def add_n(values):
    out = values[0]
    for v in values[1:]:
        out = out + v           # list order, left to right
    return out


def gather(params, indices):
    return _t(params)[_t(indices).to(torch.int64)]


def range(*args):  # noqa: A001 - mirrors tf.range
    vals = [int(a) for a in args]
    return torch.arange(*vals, dtype=torch.int32)


def linspace(start, stop, num):
    """TF 1.x LinSpace kernel: out[i] = start + step * i, step = (stop-start)/(num-1), all in T."""
    num = int(num)
    f = _FLOAT[0]
    start_t = torch.tensor(start, dtype=f)
    if num == 1:
        return start_t.reshape(1)
    step = (torch.tensor(stop, dtype=f) - start_t) / torch.tensor(num - 1, dtype=f)
    i = torch.arange(num, dtype=f)
    return start_t + step * i    # separate mul and add: stock TF 1.11 wheels carry no FMA


def meshgrid(x, y):
    # default indexing='xy'
    x = _t(x)
    y = _t(y)
    return x.reshape(1, -1).repeat(y.numel(), 1), y.reshape(-1, 1).repeat(1, x.numel())


def pad(x, paddings, mode="CONSTANT"):
    assert mode.upper() == "CONSTANT"
    x = _t(x)
    flat = []
    for lo, hi in reversed(paddings):
        flat += [int(lo), int(hi)]
    return torch.nn.functional.pad(x, flat, mode='constant', value=0)


def div_no_nan(x, y):
    x = _t(x)
    y = _t(y)
    safe = torch.where(y == 0, torch.ones_like(y), y)
    return torch.where(y == 0, torch.zeros_like(x * y), x / safe)


@contextlib.contextmanager
def variable_scope(*args, **kwargs):
    yield None


def install():
    """Register this module as `tensorflow` so `import tensorflow as tf` resolves to it."""
    mod = sys.modules[__name__]
    sys.modules['tensorflow'] = mod
    return mod


def load_reference_methods(path, class_name, method_names, namespace):
    """Compile the UNMODIFIED source of selected methods of a reference class (its file imports packages that do not
    exist here, so the file cannot be executed whole): the method bodies are cut out of the file's AST with their
    original line numbers and executed against `namespace` (+ this shim as `tf`).  Returns {name: function}."""
    import ast
    install()
    with open(path, 'r') as fh:
        tree = ast.parse(fh.read(), path)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name][0]
    funcs = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in method_names]
    assert len(funcs) == len(method_names), 'reference methods not found'
    mod = ast.Module(body=funcs, type_ignores=[])
    ns = dict(namespace)
    ns['tf'] = sys.modules[__name__]
    exec(compile(mod, path, 'exec'), ns)
    return {n: ns[n] for n in method_names}


def load_reference(path, name):
    """Execute an unmodified reference source file against the shim and return its module."""
    install()
    with open(path, 'r') as fh:
        src = fh.read()
    mod = types.ModuleType(name)
    mod.__file__ = path
    exec(compile(src, path, 'exec'), mod.__dict__)
    return mod
