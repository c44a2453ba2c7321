"""Torch-backed stand-in for the handful of `tf.*` symbols the reference's loss path touches.

TEST INFRASTRUCTURE ONLY.  TensorFlow is not installed in this image, so the only way to execute
the reference's own lines (`/root/reference/gan_utils.py`, `data_utils.py:478-586`) is to give them
a module called `tensorflow` that maps each symbol they use onto the equivalent torch CPU op.
Nothing in the product path (`kccotgan_b200/`) may import this package.

Symbols provided = the complete list the two reference files use:
gan_utils.py: float32, expand_dims, reduce_sum, math.{add,subtract,log,abs,greater,reduce_std,
reduce_logsumexp}, shape, cast, ones, range, squeeze, transpose, reduce_logsumexp, exp, reshape.
KernelSmoothing: range(a,b,dtype), newaxis, constant, meshgrid, pad(REFLECT), nn.conv{1,2,3}d,
reduce_max.

`set_default_dtype(torch.float64)` before calling turns the same reference source into the fp64
arbiter: every place the reference writes `tf.float32` resolves to the current default dtype.
"""
import types as _types

import torch as _torch

newaxis = None


class _DefaultFloat:
    """`tf.float32` placeholder that resolves to torch's current default dtype at use time."""

    def resolve(self):
        return _torch.get_default_dtype()


float32 = _DefaultFloat()


def _dt(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, _DefaultFloat):
        return dtype.resolve()
    return dtype


def _t(x, dtype=None):
    if isinstance(x, _torch.Tensor):
        return x if dtype is None else x.to(dtype)
    return _torch.as_tensor(x, dtype=dtype)


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def reduce_sum(x, axis=None, keepdims=False):
    x = _t(x)
    if axis is None:
        return x.sum()
    return x.sum(dim=axis, keepdim=keepdims)


def reduce_max(x, axis=None, keepdims=False):
    x = _t(x)
    if axis is None:
        return x.max()
    return x.amax(dim=axis, keepdim=keepdims)


def reduce_logsumexp(x, axis=None, keepdims=False):
    x = _t(x)
    if axis is None:
        return _torch.logsumexp(x.reshape(-1), dim=0)
    return _torch.logsumexp(x, dim=axis, keepdim=keepdims)


class _Shape(tuple):
    """tf.shape(x): indexable, entries int-convertible."""


def shape(x):
    return _Shape(int(s) for s in _t(x).shape)


def cast(x, dtype=None):
    dtype = _dt(dtype)
    if isinstance(x, _torch.Tensor):
        return x.to(dtype)
    return _torch.tensor(float(x), dtype=dtype)


def ones(n, dtype=None):
    if isinstance(n, _torch.Tensor):
        n = int(n)
    return _torch.ones(n, dtype=_dt(dtype))


def constant(v, dtype=None):
    return _torch.as_tensor(v, dtype=_dt(dtype))


def range(*args, dtype=None):  # noqa: A001 - mirrors tf.range
    import builtins
    args = [int(a) for a in args]
    if dtype is None:
        # `for i in tf.range(L)`: yields scalars comparable with >=
        return builtins.range(*args)
    return _torch.arange(*args, dtype=_dt(dtype))


def squeeze(x, axis=None):
    return _t(x).squeeze() if axis is None else _t(x).squeeze(axis)


def transpose(x, perm=None):
    x = _t(x)
    if perm is None:
        import builtins
        perm = tuple(reversed(builtins.range(x.dim())))
    return x.permute(*perm)


def exp(x):
    return _torch.exp(_t(x))


def reshape(x, shp):
    return _t(x).reshape([int(s) for s in shp])


def meshgrid(*xs, indexing="xy"):
    return _torch.meshgrid(*xs, indexing=indexing)


def pad(x, paddings, mode="CONSTANT"):
    """tf.pad with mode REFLECT (edge sample not repeated) over arbitrary axes, one axis at a time."""
    x = _t(x)
    pads = [(int(a), int(b)) for a, b in _t(paddings).tolist()]
    if mode.upper() != "REFLECT":
        raise NotImplementedError(mode)
    for ax, (lo, hi) in enumerate(pads):
        if lo == 0 and hi == 0:
            continue
        n = x.shape[ax]
        if lo >= n or hi >= n:
            raise ValueError("REFLECT padding must be smaller than the dimension")
        idx = list(_builtin_range(lo, 0, -1)) + list(_builtin_range(n)) + \
            list(_builtin_range(n - 2, n - 2 - hi, -1))
        x = x.index_select(ax, _torch.tensor(idx, dtype=_torch.long))
    return x


import builtins as _b  # noqa: E402

_builtin_range = _b.range


def _conv_nd(nd, x, w, padding):
    """TF layout: input [N, spatial..., Cin], filter [k..., Cin, Cout], VALID, stride 1,
    cross-correlation (same as torch.nn.functional.convNd)."""
    if padding != "VALID":
        raise NotImplementedError(padding)
    x = _t(x)
    w = _t(w, x.dtype)
    sp = list(_builtin_range(1, 1 + nd))
    xin = x.permute(0, 1 + nd, *sp)                       # -> [N, Cin, spatial...]
    win = w.permute(nd + 1, nd, *_builtin_range(nd))      # -> [Cout, Cin, k...]
    f = {1: _torch.nn.functional.conv1d, 2: _torch.nn.functional.conv2d,
         3: _torch.nn.functional.conv3d}[nd]
    y = f(xin, win)
    return y.permute(0, *_builtin_range(2, 2 + nd), 1)    # -> [N, spatial..., Cout]


nn = _types.SimpleNamespace(
    conv1d=lambda x, w, stride=1, padding="VALID": _conv_nd(1, x, w, padding),
    conv2d=lambda x, w, strides=None, padding="VALID": _conv_nd(2, x, w, padding),
    conv3d=lambda x, w, strides=None, padding="VALID": _conv_nd(3, x, w, padding),
)


def _reduce_std(x, axis=None, keepdims=False):
    x = _t(x)
    if axis is None:
        return x.std(unbiased=False)
    return x.std(dim=axis, unbiased=False, keepdim=keepdims)


math = _types.SimpleNamespace(
    add=lambda a, b: _t(a) + _t(b),
    subtract=lambda a, b: _t(a) - _t(b),
    log=lambda x: _torch.log(_t(x)),
    abs=lambda x: _torch.abs(_t(x)),
    greater=lambda a, b: bool(_t(a) > _t(b)),
    reduce_std=_reduce_std,
    reduce_logsumexp=reduce_logsumexp,
)
