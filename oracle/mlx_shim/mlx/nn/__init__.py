"""mlx.nn stand-in over torch (CPU). See oracle/mlx_shim/mlx/__init__.py."""
from __future__ import annotations

import math as _math

import torch as _torch

from .. import core as mx


class Module:
    """mlx.nn.Module is a dict-like container; the reference needs attribute storage, __call__ and — for
    ``LTXModel.from_pretrained`` (ltx.py:535-885) — the parameter tree: ``parameters()`` (nested dict of arrays,
    attributes starting with "_" excluded), ``named_modules()``, ``load_weights(list of (name, array), strict)``."""

    def __init__(self):
        pass

    def _members(self):
        return {k: v for k, v in self.__dict__.items() if not k.startswith("_")}

    def parameters(self):
        def walk(v):
            if isinstance(v, Module):
                return {k: w for k, w in ((k, walk(x)) for k, x in v._members().items()) if w is not None}
            if isinstance(v, dict):
                return {k: w for k, w in ((k, walk(x)) for k, x in v.items()) if w is not None} or None
            if isinstance(v, (list, tuple)):
                items = [walk(x) for x in v]
                return items if any(i is not None for i in items) else None
            return v if isinstance(v, mx.array) else None

        return walk(self)

    def named_modules(self):
        out = []

        def walk(prefix, v):
            if isinstance(v, Module):
                out.append((prefix, v))
                for k, x in v._members().items():
                    walk(f"{prefix}.{k}" if prefix else k, x)
            elif isinstance(v, dict):
                for k, x in v.items():
                    walk(f"{prefix}.{k}" if prefix else str(k), x)
            elif isinstance(v, (list, tuple)):
                for i, x in enumerate(v):
                    walk(f"{prefix}.{i}" if prefix else str(i), x)

        walk("", self)
        return out

    def _resolve(self, name):
        """(container, last key) of a dotted parameter / module path."""
        obj = self
        parts = name.split(".")
        for part in parts[:-1]:
            if isinstance(obj, dict):
                obj = obj[int(part)] if part.isdigit() and int(part) in obj else obj[part]
            elif isinstance(obj, (list, tuple)):
                obj = obj[int(part)]
            else:
                obj = getattr(obj, part)
        return obj, parts[-1]

    def load_weights(self, weights, strict=True):
        """mlx: with strict every parameter must be supplied with its own shape, and nothing else."""
        from ..utils import tree_flatten

        items = list(weights.items()) if isinstance(weights, dict) else list(weights)
        current = dict(tree_flatten(self.parameters()))
        if strict:
            given = {k for k, _ in items}
            if given - set(current):
                raise ValueError(f"Received parameters not in model: {sorted(given - set(current))[:8]}")
            if set(current) - given:
                raise ValueError(f"Missing parameters: {sorted(set(current) - given)[:8]}")
        for k, v in items:
            if k not in current:
                if strict:
                    raise ValueError(f"unknown parameter {k}")
                continue
            if strict and tuple(v.shape) != tuple(current[k].shape):
                raise ValueError(f"Expected shape {current[k].shape} but received shape {v.shape} for parameter {k}")
            obj, last = self._resolve(k)
            if isinstance(obj, dict):
                obj[last] = v
            else:
                setattr(obj, last, v)
        return self

    def eval(self):
        return self


class Linear(Module):
    """y = x W^T + b with W stored (out, in); default init U(-1/sqrt(in), 1/sqrt(in))."""

    def __init__(self, input_dims, output_dims, bias=True):
        super().__init__()
        k = 1.0 / _math.sqrt(input_dims)
        self.weight = mx.array((_torch.rand(output_dims, input_dims) * 2 - 1) * k)
        if bias:
            self.bias = mx.array((_torch.rand(output_dims) * 2 - 1) * k)

    def to_quantized(self, group_size=64, bits=4, mode="affine"):
        return QuantizedLinear.from_linear(self, group_size, bits, mode)

    def __call__(self, x):
        w = self.weight._t
        b = self.bias._t if "bias" in self.__dict__ else None
        dt = _torch.promote_types(x._t.dtype, w.dtype)  # mlx promotes mixed operands (fp32 activations x bf16 weights -> fp32)
        return mx.array(_torch.nn.functional.linear(x._t.to(dt), w.to(dt), None if b is None else b.to(dt)))


class Conv2d(Module):
    """mlx nn.Conv2d: NHWC input, weight (out, kh, kw, in), default init U(+-1/sqrt(in*kh*kw)), zero bias."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True):
        super().__init__()
        ks = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        k = 1.0 / _math.sqrt(in_channels * ks[0] * ks[1])
        self.weight = mx.array((_torch.rand(out_channels, ks[0], ks[1], in_channels // groups) * 2 - 1) * k)
        if bias:
            self.bias = mx.zeros((out_channels,))
        self.stride, self.padding, self.dilation, self.groups = stride, padding, dilation, groups

    def __call__(self, x):
        y = mx.conv2d(x, self.weight, self.stride, self.padding, self.dilation, self.groups)
        return y + self.bias if "bias" in self.__dict__ else y


class Conv3d(Module):
    """mlx nn.Conv3d: NDHWC input, weight (out, kd, kh, kw, in), default init U(+-1/sqrt(in*kd*kh*kw)), zero bias."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, bias=True):
        super().__init__()
        ks = (kernel_size,) * 3 if isinstance(kernel_size, int) else tuple(kernel_size)
        k = 1.0 / _math.sqrt(in_channels * ks[0] * ks[1] * ks[2])
        self.weight = mx.array((_torch.rand(out_channels, *ks, in_channels) * 2 - 1) * k)
        if bias:
            self.bias = mx.zeros((out_channels,))
        self.stride, self.padding, self.dilation = stride, padding, dilation

    def __call__(self, x):
        y = mx.conv3d(x, self.weight, self.stride, self.padding, self.dilation)
        return y + self.bias if "bias" in self.__dict__ else y


class RMSNorm(Module):
    def __init__(self, dims, eps=1e-5):
        super().__init__()
        self.weight = mx.ones((dims,))
        self.eps = eps

    def __call__(self, x):
        return mx.fast.rms_norm(x, self.weight, self.eps)


class LayerNorm(Module):
    def __init__(self, dims, eps=1e-5, affine=True, bias=True):
        super().__init__()
        self.dims, self.eps, self.affine = dims, eps, affine
        if affine:
            self.weight = mx.ones((dims,))
            if bias:
                self.bias = mx.zeros((dims,))

    def __call__(self, x):
        t = x._t
        y = _torch.nn.functional.layer_norm(t.float(), (self.dims,), eps=self.eps)
        if self.affine:
            y = y * self.weight._t.float()
            if "bias" in self.__dict__:
                y = y + self.bias._t.float()
        return mx.array(y.to(t.dtype))


def silu(x):
    return mx.array(_torch.nn.functional.silu(x._t))


def gelu(x):
    return mx.array(_torch.nn.functional.gelu(x._t))


def gelu_approx(x):
    """0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))"""
    return mx.array(_torch.nn.functional.gelu(x._t, approximate="tanh"))


class SiLU(Module):
    def __call__(self, x):
        return silu(x)


class GELU(Module):
    def __init__(self, approx="none"):
        super().__init__()
        self.approx = approx

    def __call__(self, x):
        return gelu_approx(x) if self.approx in ("tanh", "precise") else gelu(x)


class QuantizedLinear(Module):
    """mlx nn.QuantizedLinear: parameters ``weight`` (uint32, packed), ``scales``, ``biases`` (+ the linear's ``bias``);
    forward mx.quantized_matmul(x, weight, scales, biases, transpose=True) + bias."""

    def __init__(self, input_dims, output_dims, bias=True, group_size=64, bits=4, mode="affine"):
        super().__init__()
        if mode != "affine":
            raise NotImplementedError("mlx shim: only affine quantisation")
        self.group_size, self.bits, self.mode = group_size, bits, mode
        k = 1.0 / _math.sqrt(input_dims)
        self.weight, self.scales, self.biases = mx.quantize(mx.array((_torch.rand(output_dims, input_dims) * 2 - 1) * k),
                                                            group_size, bits)
        if bias:
            self.bias = mx.zeros((output_dims,))

    @classmethod
    def from_linear(cls, lin, group_size=64, bits=4, mode="affine"):
        out_f, in_f = lin.weight.shape
        q = cls(in_f, out_f, False, group_size, bits, mode)
        q.weight, q.scales, q.biases = mx.quantize(lin.weight, group_size, bits)
        if "bias" in lin.__dict__:
            q.bias = lin.bias
        return q

    def __call__(self, x):
        y = mx.quantized_matmul(x, self.weight, self.scales, self.biases, transpose=True, group_size=self.group_size,
                                bits=self.bits)
        return y + self.bias if "bias" in self.__dict__ else y


def quantize(model, group_size=64, bits=4, *, mode="affine", class_predicate=None):
    """mlx nn.quantize: every module with ``to_quantized`` for which class_predicate(path, module) holds is replaced,
    in place, by its quantised twin."""
    class_predicate = class_predicate or (lambda _, m: hasattr(m, "to_quantized"))
    for path, m in model.named_modules():
        if not path or isinstance(m, QuantizedLinear) or not hasattr(m, "to_quantized"):
            continue
        if class_predicate(path, m):
            obj, last = model._resolve(path)
            new = m.to_quantized(group_size=group_size, bits=bits, mode=mode)
            if isinstance(obj, dict):
                obj[int(last) if last.isdigit() and int(last) in obj else last] = new
            elif isinstance(obj, list):
                obj[int(last)] = new
            else:
                setattr(obj, last, new)
    return model


def value_and_grad(*a, **k):
    raise NotImplementedError("mlx shim: training is out of scope")
