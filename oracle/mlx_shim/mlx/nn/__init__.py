"""mlx.nn stand-in over torch (CPU). See oracle/mlx_shim/mlx/__init__.py."""
from __future__ import annotations

import math as _math

import torch as _torch

from .. import core as mx


class Module:
    """mlx.nn.Module is a dict-like container; the reference only needs attribute storage + __call__."""

    def __init__(self):
        pass

    def parameters(self):
        return {k: v for k, v in self.__dict__.items()}

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

    def __call__(self, x):
        w = self.weight._t
        b = self.bias._t if "bias" in self.__dict__ else None
        return mx.array(_torch.nn.functional.linear(x._t, w, b))


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


def quantize(*a, **k):
    raise NotImplementedError("mlx shim: quantisation is out of scope for golden generation")


def value_and_grad(*a, **k):
    raise NotImplementedError("mlx shim: training is out of scope")
