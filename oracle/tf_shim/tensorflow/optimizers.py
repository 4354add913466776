"""Keras OptimizerV2 update rules restated in torch (the NOT-pinned part; see ../README.md).  Dense form for
ordinary variables, the *_sparse form (m*b1 + g*(1-b1)) for variables read through tf.gather."""
import torch


class _Base:
    defaults = {}

    def __init__(self, **kwargs):
        unknown = set(kwargs) - set(self.defaults)
        if unknown:
            raise TypeError(f"unexpected keyword arguments {sorted(unknown)}")
        self.hp = dict(self.defaults, **kwargs)
        self.iterations = 0
        self.slots = {}

    def apply_gradients(self, grads_and_vars):
        self.iterations += 1
        with torch.no_grad():
            for g, v in grads_and_vars:
                if g is None:
                    continue
                self._update(v, g._t, getattr(v, "_gathered", False))


class SGD(_Base):
    defaults = dict(learning_rate=0.01)

    def _update(self, v, g, sparse):
        v._t -= self.hp["learning_rate"] * g


class Adamax(_Base):
    defaults = dict(learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7)

    def _update(self, v, g, sparse):
        dt = v._t.dtype
        lr, b1, b2, eps = (torch.tensor(self.hp[k], dtype=dt) for k in ("learning_rate", "beta_1", "beta_2", "epsilon"))
        one = torch.tensor(1.0, dtype=dt)
        m, u = self.slots.setdefault(id(v), (torch.zeros_like(v._t), torch.zeros_like(v._t)))
        b1p = torch.pow(b1, torch.tensor(float(self.iterations), dtype=dt))
        if sparse:
            m.copy_(m * b1 + g * (one - b1))
        else:
            m.add_((g - m) * (one - b1))
        torch.maximum(u * b2, g.abs(), out=u)
        v._t -= (lr / (one - b1p)) * (m / (u + eps))


class Adam(_Base):
    defaults = dict(learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7)

    def _update(self, v, g, sparse):
        dt = v._t.dtype
        lr, b1, b2, eps = (torch.tensor(self.hp[k], dtype=dt) for k in ("learning_rate", "beta_1", "beta_2", "epsilon"))
        one = torch.tensor(1.0, dtype=dt)
        m, u = self.slots.setdefault(id(v), (torch.zeros_like(v._t), torch.zeros_like(v._t)))
        t = torch.tensor(float(self.iterations), dtype=dt)
        lr_t = lr * torch.sqrt(one - torch.pow(b2, t)) / (one - torch.pow(b1, t))
        if sparse:
            m.copy_(m * b1 + g * (one - b1))
            u.copy_(u * b2 + (g * g) * (one - b2))
        else:
            m.add_((g - m) * (one - b1))
            u.add_((g * g - u) * (one - b2))
        v._t -= (m * lr_t) / (u.sqrt() + eps)


class _Unavailable(_Base):
    def __init__(self, **kwargs):
        raise NotImplementedError("optimizer not restated in tf_shim")


Adadelta = Ftrl = Nadam = RMSprop = Adagrad = _Unavailable
