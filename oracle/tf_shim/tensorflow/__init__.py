"""Fake `tensorflow` over torch (CPU).  Only the API surface /root/reference/calamity touches."""
import contextlib

import numpy as np
import torch

from . import optimizers  # noqa: F401
from . import linalg  # noqa: F401
from . import config  # noqa: F401
from . import profiler  # noqa: F401
from .tensor import Tensor, Variable, _t, _unwrap, _torch_dtype  # noqa: F401

float32 = np.float32
float64 = np.float64


def convert_to_tensor(value, dtype=None):
    return Tensor(_t(value, dtype))


def constant(value, dtype=None):
    return Tensor(_t(value, dtype))


def reduce_sum(x, axis=None):
    x = _unwrap(x)
    return Tensor(x.sum() if axis is None else x.sum(dim=axis))


def stack(values, axis=0):
    return Tensor(torch.stack([_unwrap(v) for v in values], dim=axis))


def square(x):
    return Tensor(_unwrap(x) ** 2)


def abs(x):  # noqa: A001
    return Tensor(torch.abs(_unwrap(x)))


def maximum(a, b):
    return Tensor(torch.maximum(_unwrap(a), torch.as_tensor(_unwrap(b), dtype=_unwrap(a).dtype)))


def gather(params, indices):
    idx = torch.as_tensor(np.asarray(indices), dtype=torch.long)
    if isinstance(params, Variable):
        params._gathered = True  # Keras sees IndexedSlices gradients for gathered variables
    return Tensor(_unwrap(params)[idx])


def gather_nd(params, indices):
    idx = torch.as_tensor(np.asarray(indices), dtype=torch.long)
    p = _unwrap(params)
    return Tensor(p[tuple(idx[..., k] for k in range(idx.shape[-1]))])


def reshape(x, shape):
    return Tensor(_unwrap(x).reshape(tuple(int(s) for s in shape)))


def transpose(x):
    x = _unwrap(x)
    return Tensor(x.permute(*reversed(range(x.dim()))))


def pad(x, paddings):
    flat = []
    for lo, hi in reversed(list(paddings)):
        flat += [int(lo), int(hi)]
    return Tensor(torch.nn.functional.pad(_unwrap(x), flat))


def ones(shape, dtype=None):
    return Tensor(torch.ones(shape, dtype=_torch_dtype(dtype)))


def meshgrid(a, b, indexing="xy"):
    ga, gb = torch.meshgrid(_unwrap(a), _unwrap(b), indexing=indexing)
    return Tensor(ga), Tensor(gb)


class GradientTape:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def gradient(self, loss, variables):
        grads = torch.autograd.grad(_unwrap(loss), [v._t for v in variables], allow_unused=True)
        return [None if g is None else Tensor(g) for g in grads]


def function(**kwargs):
    def deco(fn):
        return fn

    return deco


@contextlib.contextmanager
def device(name):
    yield


class _Math:
    square = staticmethod(square)
    abs = staticmethod(abs)

    @staticmethod
    def sqrt(x):
        return Tensor(torch.sqrt(_unwrap(x)))


math = _Math()


class _ExperimentalNumpy:
    @staticmethod
    def outer(a, b):
        return Tensor(torch.outer(_t(a), _t(b)))

    @staticmethod
    def sinc(x):
        return Tensor(torch.sinc(_unwrap(x)))


class _Experimental:
    numpy = _ExperimentalNumpy()


experimental = _Experimental()
