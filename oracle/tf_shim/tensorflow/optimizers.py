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
    defaults = dict(learning_rate=0.01, momentum=0.0, nesterov=False)

    def _update(self, v, g, sparse):
        lr, mom = self.hp["learning_rate"], self.hp["momentum"]
        if mom == 0:
            v._t -= lr * g
            return
        (m,) = self.slots.setdefault(id(v), (torch.zeros_like(v._t),))
        m.copy_(m * mom - lr * g)  # ApplyKerasMomentum
        v._t += (m * mom - lr * g) if self.hp["nesterov"] else m


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


def _hp(self, v, *names):
    return (torch.tensor(self.hp[k], dtype=v._t.dtype) for k in names)


class RMSprop(_Base):
    defaults = dict(learning_rate=0.001, rho=0.9, momentum=0.0, epsilon=1e-7)

    def _update(self, v, g, sparse):
        lr, rho, mom, eps = _hp(self, v, "learning_rate", "rho", "momentum", "epsilon")
        m, u = self.slots.setdefault(id(v), (torch.zeros_like(v._t), torch.zeros_like(v._t)))
        u.copy_(rho * u + (1 - rho) * (g * g))
        if self.hp["momentum"] > 0:  # fused ApplyRMSProp: epsilon inside the square root
            m.copy_(mom * m + lr * g / torch.sqrt(u + eps))
            v._t -= m
        else:
            v._t -= lr * g / (torch.sqrt(u) + eps)


class Adagrad(_Base):
    defaults = dict(learning_rate=0.001, initial_accumulator_value=0.1, epsilon=1e-7)

    def _update(self, v, g, sparse):
        lr, eps = _hp(self, v, "learning_rate", "epsilon")
        (u,) = self.slots.setdefault(id(v), (torch.full_like(v._t, self.hp["initial_accumulator_value"]),))
        u += g * g
        v._t -= lr * g / (torch.sqrt(u) + eps)


class Adadelta(_Base):
    defaults = dict(learning_rate=0.001, rho=0.95, epsilon=1e-7)

    def _update(self, v, g, sparse):
        lr, rho, eps = _hp(self, v, "learning_rate", "rho", "epsilon")
        m, u = self.slots.setdefault(id(v), (torch.zeros_like(v._t), torch.zeros_like(v._t)))
        m.copy_(m * rho + (g * g) * (1 - rho))
        upd = torch.sqrt(u + eps) / torch.sqrt(m + eps) * g
        u.copy_(u * rho + (upd * upd) * (1 - rho))
        v._t -= upd * lr


class Nadam(_Base):
    defaults = dict(learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7)

    def apply_gradients(self, grads_and_vars):
        # Nadam._prepare_local: the momentum-schedule product advances once per step, shared by all variables
        t = self.iterations + 1
        b1 = self.hp["beta_1"]
        self._u_t = b1 * (1.0 - 0.5 * 0.96 ** (0.004 * t))
        self._u_t1 = b1 * (1.0 - 0.5 * 0.96 ** (0.004 * (t + 1)))
        self._ms_new = getattr(self, "_m_schedule", 1.0) * self._u_t
        self._m_schedule = self._ms_new
        super().apply_gradients(grads_and_vars)

    def _update(self, v, g, sparse):
        dt = v._t.dtype
        lr, b1, b2, eps = _hp(self, v, "learning_rate", "beta_1", "beta_2", "epsilon")
        u_t, u_t1, ms_new = (torch.tensor(x, dtype=dt) for x in (self._u_t, self._u_t1, self._ms_new))
        ms_next = ms_new * u_t1
        m, u = self.slots.setdefault(id(v), (torch.zeros_like(v._t), torch.zeros_like(v._t)))
        g_prime = g / (1 - ms_new)
        m.copy_(b1 * m + (1 - b1) * g)
        m_prime = m / (1 - ms_next)
        u.copy_(b2 * u + (1 - b2) * (g * g))
        v_prime = u / (1 - torch.pow(b2, torch.tensor(float(self.iterations), dtype=dt)))
        m_bar = (1 - u_t) * g_prime + u_t1 * m_prime
        v._t -= lr * m_bar / (torch.sqrt(v_prime) + eps)


class Ftrl(_Base):
    defaults = dict(learning_rate=0.001, learning_rate_power=-0.5, initial_accumulator_value=0.1,
                    l1_regularization_strength=0.0, l2_regularization_strength=0.0)

    def _update(self, v, g, sparse):
        lr, l1, l2 = _hp(self, v, "learning_rate", "l1_regularization_strength", "l2_regularization_strength")
        lp = self.hp["learning_rate_power"]
        acc, lin = self.slots.setdefault(id(v), (torch.full_like(v._t, self.hp["initial_accumulator_value"]),
                                                 torch.zeros_like(v._t)))
        pw = torch.sqrt if lp == -0.5 else (lambda x: torch.pow(x, -lp))
        acc_new = acc + g * g
        lin += g - (pw(acc_new) - pw(acc)) / lr * v._t
        quadratic = pw(acc_new) / lr + 2 * l2
        acc.copy_(acc_new)
        v._t.copy_(torch.where(lin.abs() > l1, (torch.sign(lin) * l1 - lin) / quadratic, torch.zeros_like(lin)))




class LAMB(_Base):
    """tensorflow_addons.optimizers.LAMB (_resource_apply_dense / _sparse compute the same expressions): Adam moments with bias
    correction and a per-variable trust ratio ||w|| / ||update||."""

    defaults = dict(learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-6, weight_decay=0.0)

    def _update(self, v, g, sparse):
        dt = v._t.dtype
        lr, b1, b2, eps, wd = (torch.tensor(self.hp[k], dtype=dt) for k in ("learning_rate", "beta_1", "beta_2", "epsilon", "weight_decay"))
        one = torch.tensor(1.0, dtype=dt)
        t = torch.tensor(float(self.iterations), dtype=dt)
        m, u = self.slots.setdefault(id(v), (torch.zeros_like(v._t), torch.zeros_like(v._t)))
        m.copy_(m * b1 + g * (one - b1))
        u.copy_(u * b2 + (g * g) * (one - b2))
        upd = (m / (one - torch.pow(b1, t))) / (torch.sqrt(u / (one - torch.pow(b2, t))) + eps)
        if float(wd) != 0.0:
            upd = upd + wd * v._t
        w_norm, g_norm = torch.linalg.vector_norm(v._t), torch.linalg.vector_norm(upd)
        ratio = (w_norm / g_norm) if (float(w_norm) > 0 and float(g_norm) > 0) else one
        v._t -= ratio * lr * upd
