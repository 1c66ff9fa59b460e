"""mlx.core stand-in over torch (CPU). See oracle/mlx_shim/mlx/__init__.py."""
from __future__ import annotations

import math as _math
from types import SimpleNamespace as _NS

import numpy as _np
import torch as _torch

Dtype = _torch.dtype
float32, float16, bfloat16 = _torch.float32, _torch.float16, _torch.bfloat16
int32, int64, uint8, uint32, bool_ = _torch.int32, _torch.int64, _torch.uint8, _torch.int64, _torch.bool  # uint32 lives in int64
uint16 = _torch.uint16


def _unwrap(x):
    if isinstance(x, array):
        return x._t
    if isinstance(x, (list, tuple)):
        return type(x)(_unwrap(v) for v in x)
    return x


class array:
    """Thin wrapper giving a torch.Tensor the mlx.core.array surface used by the reference."""

    __array_priority__ = 1000

    def __init__(self, value, dtype=None):
        if isinstance(value, array):
            t = value._t
        elif isinstance(value, _torch.Tensor):
            t = value
        elif isinstance(value, _np.ndarray):
            t = _torch.from_numpy(_np.ascontiguousarray(value))
        else:
            t = _torch.tensor(value)
            if t.dtype == _torch.float64:
                t = t.to(_torch.float32)  # mlx default float is float32
        if dtype is not None:
            t = t.to(dtype)
        self._t = t

    # -- properties
    shape = property(lambda s: tuple(s._t.shape))
    ndim = property(lambda s: s._t.dim())
    dtype = property(lambda s: s._t.dtype)
    size = property(lambda s: s._t.numel())

    def astype(self, dt):
        return array(self._t.to(dt))

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return array(self._t.reshape(*shape))

    def tolist(self):
        return self._t.tolist()

    def item(self):
        return self._t.item()

    def __len__(self):
        return self._t.shape[0]

    def __setitem__(self, idx, value):
        # mlx arrays support slice assignment (``a[..., lo:hi] = b``; used by video_vae/tiling.py:429-434)
        if isinstance(idx, tuple):
            idx = tuple(_unwrap(i) for i in idx)
        else:
            idx = _unwrap(idx)
        self._t[idx] = _unwrap(value)

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            idx = tuple(_unwrap(i) for i in idx)
        else:
            idx = _unwrap(idx)
        # mlx (like numpy) allows negative slice steps; torch does not: reverse those axes with flip
        items = idx if isinstance(idx, tuple) else (idx,)
        if any(isinstance(i, slice) and i.step is not None and i.step < 0 for i in items):
            t, fixed, axis = self._t, [], 0
            for i in items:
                if i is None:
                    fixed.append(i)
                    continue
                if isinstance(i, slice) and i.step is not None and i.step < 0:
                    sel = list(range(t.shape[axis]))[i]
                    t = t.index_select(axis, _torch.tensor(sel, dtype=_torch.long))
                    fixed.append(slice(None))
                else:
                    fixed.append(i)
                if not isinstance(i, int):
                    axis += 1
                else:
                    axis += 1
            return array(t[tuple(fixed)])
        return array(self._t[idx])

    def flatten(self):
        return array(self._t.flatten())

    def __array__(self, dtype=None):
        a = self._t.detach().float().numpy() if self._t.dtype == _torch.bfloat16 else self._t.detach().numpy()
        return a.astype(dtype) if dtype is not None else a

    # -- arithmetic (python scalars keep the array dtype, as in mlx)
    def _bin(self, other, fn, rev=False):
        o = _unwrap(other)
        if isinstance(o, _np.ndarray):
            o = _torch.from_numpy(o)
        a, b = (o, self._t) if rev else (self._t, o)
        return array(fn(a, b))

    __add__ = lambda s, o: s._bin(o, lambda a, b: a + b)
    __radd__ = lambda s, o: s._bin(o, lambda a, b: a + b, True)
    __sub__ = lambda s, o: s._bin(o, lambda a, b: a - b)
    __rsub__ = lambda s, o: s._bin(o, lambda a, b: a - b, True)
    __mul__ = lambda s, o: s._bin(o, lambda a, b: a * b)
    __rmul__ = lambda s, o: s._bin(o, lambda a, b: a * b, True)
    __truediv__ = lambda s, o: s._bin(o, lambda a, b: a / b)
    __rtruediv__ = lambda s, o: s._bin(o, lambda a, b: a / b, True)
    __pow__ = lambda s, o: s._bin(o, lambda a, b: a ** b)
    __matmul__ = lambda s, o: s._bin(o, lambda a, b: a @ b)
    __neg__ = lambda s: array(-s._t)
    __gt__ = lambda s, o: s._bin(o, lambda a, b: a > b)
    __lt__ = lambda s, o: s._bin(o, lambda a, b: a < b)
    __ge__ = lambda s, o: s._bin(o, lambda a, b: a >= b)
    __le__ = lambda s, o: s._bin(o, lambda a, b: a <= b)

    @property
    def T(self):
        return array(self._t.T)

    def __repr__(self):
        return f"array({self._t})"


def _w(t):
    return array(t)


def matmul(a, b):
    return _w(_unwrap(a) @ _unwrap(b))


def reshape(a, shape):
    return _w(_unwrap(a).reshape(tuple(shape)))


def swapaxes(a, a1, a2):
    return _w(_unwrap(a).transpose(a1, a2))


def transpose(a, axes=None):
    t = _unwrap(a)
    return _w(t.permute(*axes) if axes is not None else t.permute(*reversed(range(t.dim()))))


def expand_dims(a, axis):
    return _w(_unwrap(a).unsqueeze(axis))


def squeeze(a, axis=None):
    t = _unwrap(a)
    return _w(t.squeeze() if axis is None else t.squeeze(axis))


def zeros(shape, dtype=float32):
    return _w(_torch.zeros(tuple(shape) if not isinstance(shape, int) else (shape,), dtype=dtype))


def ones(shape, dtype=float32):
    return _w(_torch.ones(tuple(shape) if not isinstance(shape, int) else (shape,), dtype=dtype))


def zeros_like(a):
    return _w(_torch.zeros_like(_unwrap(a)))


def ones_like(a):
    return _w(_torch.ones_like(_unwrap(a)))


def concatenate(arrs, axis=0):
    return _w(_torch.cat([_unwrap(a) for a in arrs], dim=axis))


def split(a, indices_or_sections, axis=0):
    """mx.split: an int = that many equal sections; a list = split points."""
    t = _unwrap(a)
    if isinstance(indices_or_sections, int):
        return [_w(x) for x in _torch.chunk(t, indices_or_sections, dim=axis)]
    return [_w(x) for x in _torch.tensor_split(t, list(indices_or_sections), dim=axis)]


def stack(arrs, axis=0):
    return _w(_torch.stack([_unwrap(a) for a in arrs], dim=axis))


def repeat(a, repeats, axis=None):
    return _w(_torch.repeat_interleave(_unwrap(a), repeats, dim=axis))


def tile(a, reps):
    return _w(_unwrap(a).repeat(*reps))


def full(shape, vals, dtype=float32):
    return _w(_torch.full(tuple(shape), float(_unwrap(vals)) if not isinstance(vals, (int, float)) else vals, dtype=dtype))


def broadcast_to(a, shape):
    return _w(_unwrap(a).broadcast_to(tuple(shape)))


def sin(a):
    return _w(_torch.sin(_unwrap(a)))


def cos(a):
    return _w(_torch.cos(_unwrap(a)))


def exp(a):
    return _w(_torch.exp(_unwrap(a)))


def sqrt(a):
    return _w(_torch.sqrt(_unwrap(a)))


def mean(a, axis=None, keepdims=False):
    t = _unwrap(a)
    return _w(t.mean() if axis is None else t.mean(dim=axis, keepdim=keepdims))


def var(a, axis=None, keepdims=False, ddof=0):
    """mx.var: population variance by default (ddof=0)."""
    t = _unwrap(a)
    if axis is None:
        return _w(t.var(correction=ddof))
    return _w(t.var(dim=axis, keepdim=keepdims, correction=ddof))


def _conv_nd(x, w, stride, padding, dilation, groups, nd):
    """mx.conv2d / mx.conv3d: channels-last input (N, *spatial, C_in), weight (C_out, *kernel, C_in / groups);
    cross-correlation like torch's conv, computed in the promoted dtype."""
    t, k = _unwrap(x), _unwrap(w)
    dt = _torch.promote_types(t.dtype, k.dtype)
    perm_in = (0, nd + 1) + tuple(range(1, nd + 1))          # -> channels first
    perm_w = (0, nd + 1) + tuple(range(1, nd + 1))
    perm_out = (0,) + tuple(range(2, nd + 2)) + (1,)          # -> channels last
    fn = _torch.nn.functional.conv3d if nd == 3 else _torch.nn.functional.conv2d
    y = fn(t.to(dt).permute(*perm_in), k.to(dt).permute(*perm_w), None, stride, padding, dilation, groups)
    return _w(y.permute(*perm_out).contiguous())


def conv3d(x, w, stride=1, padding=0, dilation=1, groups=1):
    return _conv_nd(x, w, stride, padding, dilation, groups, 3)


def conv2d(x, w, stride=1, padding=0, dilation=1, groups=1):
    return _conv_nd(x, w, stride, padding, dilation, groups, 2)


def power(a, b):
    a, b = _unwrap(a), _unwrap(b)
    if not isinstance(a, _torch.Tensor):
        a = _torch.tensor(a, dtype=_torch.float32)
    return _w(_torch.pow(a, b))


def linspace(start, stop, num=50, dtype=float32):
    return _w(_torch.linspace(float(start), float(stop), int(num), dtype=dtype))


def arange(start, stop=None, step=1, dtype=None):
    if stop is None:
        start, stop = 0, start
    t = _torch.arange(start, stop, step)
    if dtype is not None:
        t = t.to(dtype)
    elif t.dtype == _torch.int64:
        t = t.to(_torch.int32)
    return _w(t)


def pad(a, pad_width):
    t = _unwrap(a)
    flat = []
    for lo, hi in reversed(list(pad_width)):
        flat += [lo, hi]
    return _w(_torch.nn.functional.pad(t, flat))


def where(c, a, b):
    return _w(_torch.where(_unwrap(c), _unwrap(a), _unwrap(b)))


def maximum(a, b):
    return _w(_torch.maximum(_unwrap(a), _torch.as_tensor(_unwrap(b))))


def clip(a, a_min=None, a_max=None):
    return _w(_torch.clamp(_unwrap(a), a_min, a_max))


def eval(*args, **kwargs):  # lazy-graph barrier in mlx; a no-op here
    return None


def clear_cache():
    return None


def compile(fn=None, **kwargs):  # graph compilation in mlx; identity here
    if fn is None:
        return lambda f: f
    return fn


def _rms_norm(x, weight, eps):
    """mx.fast.rms_norm: statistics in fp32, result in the input dtype."""
    t = _unwrap(x)
    w = _unwrap(weight)
    tf = t.float()
    y = tf * _torch.rsqrt(tf.pow(2).mean(-1, keepdim=True) + eps)
    if w is not None:
        y = y * w.float()
    return _w(y.to(t.dtype))


def _sdpa(q, k, v, *, scale, mask=None):
    """mx.fast.scaled_dot_product_attention: softmax in fp32, additive float mask / boolean keep-mask."""
    q, k, v = _unwrap(q), _unwrap(k), _unwrap(v)
    s = (q.float() @ k.float().transpose(-1, -2)) * scale
    if mask is not None:
        m = _unwrap(mask)
        if m.dtype == _torch.bool:
            s = s.masked_fill(~m, float("-inf"))
        else:
            s = s + m.float()
    p = _torch.softmax(s, dim=-1)
    return _w((p @ v.float()).to(q.dtype))


fast = _NS(rms_norm=_rms_norm, scaled_dot_product_attention=_sdpa)
random = _NS(
    normal=lambda shape, dtype=float32, key=None: _w(_torch.randn(tuple(shape)).to(dtype)),
    uniform=lambda low=0.0, high=1.0, shape=(), dtype=float32, key=None: _w((_torch.rand(tuple(shape)) * (high - low) + low).to(dtype)),
    seed=lambda s: _torch.manual_seed(s),
)
metal = _NS(start_capture=lambda *a, **k: None, stop_capture=lambda *a, **k: None)


def view(a, dtype):
    """mx.view: reinterpret the bytes.  uint32 arrays are held as int64 values in this shim."""
    t = _unwrap(a)
    if t.dtype == _torch.int64:  # a uint32 array
        if dtype == uint32:
            return _w(t)
        raw = _torch.from_numpy(t.numpy().astype(_np.uint32).view(_np.int32).copy())
        return _w(raw.view(dtype))
    return _w(t.view(dtype))


# ---- affine group quantisation (mx.quantize / mx.dequantize / mx.quantized_matmul, mode="affine"), from the
# published definition: w ~= scales * q + biases per group of `group_size` along the last axis, 32/bits levels per
# uint32 word, lowest bits first; scales / biases in the input dtype.
def quantize(w, group_size=64, bits=4, mode="affine"):
    t = _unwrap(w)
    rows, cols = t.shape
    n_bins = float((1 << bits) - 1)
    g = t.reshape(rows, cols // group_size, group_size)
    hi, lo = g.max(-1, keepdim=True).values.float(), g.min(-1, keepdim=True).values.float()
    mask = lo.abs() > hi.abs()
    scales = _torch.maximum((hi - lo) / n_bins, _torch.tensor(1e-7))
    scales = _torch.where(mask, scales, -scales)
    edge = _torch.where(mask, lo, hi)
    q0 = _torch.round(edge / scales)
    scales = _torch.where(q0 != 0, edge / _torch.where(q0 != 0, q0, _torch.ones_like(q0)), scales)
    biases = _torch.where(q0 == 0, _torch.zeros_like(edge), edge)
    scales, biases = scales.to(t.dtype), biases.to(t.dtype)
    q = _torch.clamp(_torch.round((g.float() - biases.float()) / scales.float()), 0, n_bins).to(_torch.int64)
    per = 32 // bits
    q = q.reshape(rows, cols // per, per)
    packed = (q << (_torch.arange(per) * bits)).sum(-1)
    return _w(packed), _w(scales.squeeze(-1)), _w(biases.squeeze(-1))


def dequantize(w, scales, biases, group_size=64, bits=4, mode="affine"):
    p, s, b = _unwrap(w), _unwrap(scales), _unwrap(biases)
    per = 32 // bits
    q = (p.unsqueeze(-1) >> (_torch.arange(per) * bits)) & ((1 << bits) - 1)
    q = q.reshape(p.shape[0], -1, group_size).float()
    out = q * s.float().unsqueeze(-1) + b.float().unsqueeze(-1)
    return _w(out.reshape(p.shape[0], -1).to(s.dtype))


def quantized_matmul(x, w, scales, biases, transpose=True, group_size=64, bits=4, mode="affine"):
    """x @ dequantize(w)^T (transpose=True), products accumulated in fp32, result in x's dtype."""
    t = _unwrap(x)
    dt = _torch.promote_types(t.dtype, _unwrap(scales).dtype)  # mlx: result_type(x, scales, biases); operands cast to it
    sc, bi = _w(_unwrap(scales).to(dt)), _w(_unwrap(biases).to(dt))
    wd = _unwrap(dequantize(w, sc, bi, group_size, bits)).float()  # the matrix kernels dequantise tiles into that dtype
    y = t.float() @ (wd.t() if transpose else wd)
    return _w(y.to(dt))


def load(path, *a, **k):
    """mx.load of a safetensors file -> dict of arrays (the only format the reference's loaders hand it here)."""
    if not str(path).endswith(".safetensors"):
        raise NotImplementedError("mlx shim: only safetensors files")
    from safetensors.torch import load_file

    return {key: _w(t) for key, t in load_file(str(path)).items()}
