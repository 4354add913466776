import numpy as np
import torch


def _torch_dtype(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, torch.dtype):
        return dtype
    return {np.dtype("float32"): torch.float32, np.dtype("float64"): torch.float64, np.dtype("bool"): torch.bool,
            np.dtype("int64"): torch.int64, np.dtype("int32"): torch.int32}[np.dtype(dtype)]


def _unwrap(x):
    return x._t if isinstance(x, Tensor) else x


def _t(value, dtype=None):
    if isinstance(value, Tensor):
        t = value._t
    elif isinstance(value, torch.Tensor):
        t = value
    else:
        t = torch.as_tensor(np.ascontiguousarray(value))
    td = _torch_dtype(dtype)
    return t if td is None or t.dtype == td else t.to(td)


class Tensor:
    """tf.Tensor look-alike: .numpy(), .shape, .dtype (a numpy dtype), arithmetic, slicing."""

    __array_priority__ = 1000

    def __init__(self, t):
        self._t = t

    def numpy(self):
        out = self._t.detach().numpy()
        return out if out.ndim else out[()]

    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def dtype(self):
        return np.dtype(str(self._t.dtype).replace("torch.", ""))

    def __getitem__(self, idx):
        return Tensor(self._t[idx])

    def __len__(self):
        return self._t.shape[0]

    def _bin(self, other, op):
        o = _unwrap(other)
        if not isinstance(o, torch.Tensor):
            o = torch.as_tensor(np.asarray(o), dtype=self._t.dtype)
        return Tensor(op(self._t, o))

    def __add__(self, o):
        return self._bin(o, torch.add)

    __radd__ = __add__

    def __sub__(self, o):
        return self._bin(o, torch.sub)

    def __rsub__(self, o):
        return self._bin(o, lambda a, b: b - a)

    def __mul__(self, o):
        return self._bin(o, torch.mul)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._bin(o, torch.div)

    def __rtruediv__(self, o):
        return self._bin(o, lambda a, b: b / a)

    def __neg__(self):
        return Tensor(-self._t)

    def __pow__(self, p):
        return Tensor(self._t ** p)

    def __matmul__(self, o):
        return self._bin(o, torch.matmul)

    def __rmatmul__(self, o):
        return Tensor(torch.as_tensor(np.asarray(o), dtype=self._t.dtype) @ self._t)

    def __lt__(self, o):
        return bool(self._t < _unwrap(o))

    def __float__(self):
        return float(self._t)

    def __format__(self, spec):
        return format(float(self._t), spec)

    def __array__(self, dtype=None, copy=None):
        a = self._t.detach().numpy()
        return a.astype(dtype) if dtype is not None else a


class Variable(Tensor):
    def __init__(self, initial):
        super().__init__(_t(initial).detach().clone().requires_grad_(True))
        self._gathered = False

    def value(self):
        return Tensor(self._t.detach().clone())

    def assign(self, new):
        with torch.no_grad():
            self._t.copy_(_unwrap(new))
