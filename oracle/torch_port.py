"""torch-CPU op-for-op port of the reference's TensorFlow graph (TEST INFRASTRUCTURE / CPU baseline).

This is the "port" the bench reports as `cpu_baseline` (TensorFlow is not installable in this
image, so the reference's own CPU path cannot be timed).  It keeps the reference's cost structure on
purpose: the dense zero-padded [nvecs, ngrps, nbls, nfreqs] basis, the broadcast-multiply followed
by reduce over nvecs (calibration.py:1587-1590), four gathers per chunk (1594-1597), reverse-mode
autodiff for the gradient (664-666) and a Keras-v2 style optimizer step over every variable (667).
"""
import numpy as np
import torch

from .restatement import KERAS_DEFAULTS


def _t(x, dtype):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype)


class TorchProblem:
    """Holds the tensors of one integration and runs train steps the way calibration.py:663-668 does."""

    def __init__(self, g_r, g_i, fg_r, fg_i, data_r, data_i, wgts, fg_comps, corr_inds, optimizer="Adamax",
                 freeze_model=False, model_regularization=None, sky_model_r=None, sky_model_i=None,
                 dtype=torch.float32, **opt_kwargs):
        if optimizer not in KERAS_DEFAULTS:
            raise KeyError(optimizer)
        self.dtype = dtype
        self.opt_name = optimizer
        self.hp = dict(KERAS_DEFAULTS[optimizer], **opt_kwargs)
        self.freeze = freeze_model
        self.g_r = _t(g_r, dtype).requires_grad_(True)
        self.g_i = _t(g_i, dtype).requires_grad_(True)
        self.fg_r = [_t(x, dtype).requires_grad_(not freeze_model) for x in fg_r]
        self.fg_i = [_t(x, dtype).requires_grad_(not freeze_model) for x in fg_i]
        self.data_r = [_t(x, dtype) for x in data_r]
        self.data_i = [_t(x, dtype) for x in data_i]
        self.wgts = [_t(x, dtype) for x in wgts]
        self.comps = [_t(x, dtype) for x in fg_comps]
        self.a0 = [torch.as_tensor([[p[0] for p in grp] for grp in chunk], dtype=torch.long) for chunk in corr_inds]
        self.a1 = [torch.as_tensor([[p[1] for p in grp] for grp in chunk], dtype=torch.long) for chunk in corr_inds]
        self.reg = model_regularization == "sum"
        if self.reg:
            self.prior_r = torch.stack([(_t(s, dtype) * w).sum() for s, w in zip(sky_model_r, self.wgts)]).sum()
            self.prior_i = torch.stack([(_t(s, dtype) * w).sum() for s, w in zip(sky_model_i, self.wgts)]).sum()
        self.vars = [self.g_r, self.g_i] + ([] if freeze_model else self.fg_r + self.fg_i)
        self.sparse = [True, True] + [False] * (len(self.vars) - 2)
        self.slots = [(torch.zeros_like(v), torch.zeros_like(v)) for v in self.vars]
        self.iterations = 0

    def loss(self):
        chi, s_r, s_i = [], [], []
        for c in range(len(self.comps)):
            gr0, gr1 = self.g_r[self.a0[c]], self.g_r[self.a1[c]]
            gi0, gi1 = self.g_i[self.a0[c]], self.g_i[self.a1[c]]
            grgr, gigi, grgi, gigr = gr0 * gr1, gi0 * gi1, gr0 * gi1, gi0 * gr1
            vr = (self.fg_r[c] * self.comps[c]).sum(dim=0)
            vi = (self.fg_i[c] * self.comps[c]).sum(dim=0)
            m_r = (grgr + gigi) * vr + (grgi - gigr) * vi
            m_i = (gigr - grgi) * vr + (grgr + gigi) * vi
            chi.append((((self.data_r[c] - m_r) ** 2 + (self.data_i[c] - m_i) ** 2) * self.wgts[c]).sum())
            if self.reg:
                s_r.append((m_r * self.wgts[c]).sum())
                s_i.append((m_i * self.wgts[c]).sum())
        total = torch.stack(chi).sum()
        if self.reg:
            total = total + (torch.stack(s_r).sum() - self.prior_r) ** 2 + (torch.stack(s_i).sum() - self.prior_i) ** 2
        return total

    def grads(self):
        loss = self.loss()
        return loss.detach(), torch.autograd.grad(loss, self.vars)

    @torch.no_grad()
    def _apply(self, grads):
        self.iterations += 1
        t = self.iterations
        hp = self.hp
        one = torch.tensor(1.0, dtype=self.dtype)
        if self.opt_name == "SGD":
            for p, g in zip(self.vars, grads):
                p -= hp["learning_rate"] * g
            return
        lr, b1, b2, eps = (torch.tensor(hp[k], dtype=self.dtype) for k in ("learning_rate", "beta_1", "beta_2", "epsilon"))
        b1p = torch.pow(b1, torch.tensor(float(t), dtype=self.dtype))
        b2p = torch.pow(b2, torch.tensor(float(t), dtype=self.dtype))
        for p, g, (m, u), sp in zip(self.vars, grads, self.slots, self.sparse):
            if self.opt_name == "Adamax":
                if sp:
                    m.copy_(m * b1 + g * (one - b1))
                else:
                    m.add_((g - m) * (one - b1))
                torch.maximum(u * b2, g.abs(), out=u)
                p -= (lr / (one - b1p)) * (m / (u + eps))
            else:
                lr_t = lr * torch.sqrt(one - b2p) / (one - b1p)
                if sp:
                    m.copy_(m * b1 + g * (one - b1))
                    u.copy_(u * b2 + (g * g) * (one - b2))
                else:
                    m.add_((g - m) * (one - b1))
                    u.add_((g * g - u) * (one - b2))
                p -= (m * lr_t) / (u.sqrt() + eps)

    def train_step(self):
        loss, grads = self.grads()
        self._apply(grads)
        return loss

    def numpy_params(self):
        return (self.g_r.detach().numpy().copy(), self.g_i.detach().numpy().copy(),
                [x.detach().numpy().copy() for x in self.fg_r], [x.detach().numpy().copy() for x in self.fg_i])


def fit(g_r, g_i, fg_r, fg_i, data_r, data_i, wgts, fg_comps, corr_inds, use_min=False, tol=1e-14, maxsteps=10000,
        n_profile_steps=0, dtype=torch.float32, **kwargs):
    """Loop of calibration.py:681-732 around TorchProblem.train_step."""
    prob = TorchProblem(g_r, g_i, fg_r, fg_i, data_r, data_i, wgts, fg_comps, corr_inds, dtype=dtype, **kwargs)
    for _ in range(n_profile_steps + 1):
        prob.train_step()
    npdt = np.float32 if dtype == torch.float32 else np.float64
    history, min_loss, best = [], 9e99, None
    for step in range(maxsteps):
        history.append(npdt(prob.train_step().item()))
        if use_min and history[-1] < min_loss:
            min_loss = history[-1]
            best = prob.numpy_params()
        if step >= 1 and np.abs(history[-1] - history[-2]) < tol:
            break
    if not use_min:
        best = prob.numpy_params()
    return best + ({"loss": history},)
