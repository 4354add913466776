"""Python handle on a native fit plan (one GPU).  Thin: every method is one C-ABI call."""
import ctypes as C

import numpy as np

from . import _native as nat

OPTIMIZER_IDS = {"Adamax": 0, "Adam": 1, "SGD": 2, "RMSprop": 3, "Adagrad": 4, "Adadelta": 5, "Nadam": 6, "Ftrl": 7,
                 "LAMB": 8}
# tf.keras.optimizers defaults (TensorFlow >= 2.4 OptimizerV2), used when **opt_kwargs omits a field
KERAS_DEFAULTS = {
    "Adamax": dict(learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7),
    "Adam": dict(learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7),
    "SGD": dict(learning_rate=0.01, momentum=0.0, nesterov=False),
    "RMSprop": dict(learning_rate=0.001, rho=0.9, momentum=0.0, epsilon=1e-7),
    "Adagrad": dict(learning_rate=0.001, initial_accumulator_value=0.1, epsilon=1e-7),
    "Adadelta": dict(learning_rate=0.001, rho=0.95, epsilon=1e-7),
    "Nadam": dict(learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7),
    "Ftrl": dict(learning_rate=0.001, learning_rate_power=-0.5, initial_accumulator_value=0.1,
                 l1_regularization_strength=0.0, l2_regularization_strength=0.0),
    # tensorflow_addons.optimizers.LAMB (calibration.py:15, 26)
    "LAMB": dict(learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-6, weight_decay=0.0),
}
# names the reference's OPTIMIZERS dict accepts (calibration.py:17-27); all have a device implementation (LAMB: float32
# plans on one GPU; anything else raises the library's "unsupported" error, nothing is substituted)
REFERENCE_OPTIMIZERS = ("Adadelta", "Adam", "Adamax", "Ftrl", "Nadam", "SGD", "RMSprop", "Adagrad", "LAMB")


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _dtype_name(dt):
    return {np.dtype(np.float32): "float32", np.dtype(np.float64): "float64"}[np.dtype(dt)]


class FitPlan:
    """Device-resident basis + per-integration state.  Mirrors calibration.py:1143-1152 (create) and
    the per-integration calls of 1184-1300."""

    def __init__(self, layout, device=0, tile_freqs=0, basis_batch=2048, shared_basis=0):
        self._lib = nat.load()
        self.layout = layout
        # float32: the fused sm_100a kernels; float64 (precision=64): the generic device path
        self.dtype = np.dtype(getattr(layout, "dtype", np.float32))
        if self.dtype not in nat.DTYPE_IDS:
            raise TypeError(f"unsupported dtype {self.dtype}; float32 or float64")
        self._handle = C.c_void_p()
        desc = nat.PlanDesc(
            device=device, nants=layout.nants, nfreqs=layout.nfreqs, ngroups=layout.ngroups,
            group_ncomp=nat.iptr(layout.group_ncomp), group_nslots=nat.iptr(layout.group_nslots),
            slot_nbls=nat.iptr(layout.slot_nbls), bl_ant0=nat.iptr(layout.bl_ant0), bl_ant1=nat.iptr(layout.bl_ant1),
            tile_freqs=tile_freqs, dtype=nat.DTYPE_IDS[self.dtype],
            # groups with identical basis blocks: stored once, fitted by the shared-basis kernel (0 auto, 1 all, -1 never)
            group_class=nat.iptr(layout.group_class), shared_basis=int(shared_basis),
        )
        nat.check(self._lib.calb2_plan_create(C.byref(desc), C.byref(self._handle)))
        for g0 in range(0, layout.ngroups, basis_batch):
            blks = [np.ascontiguousarray(b, dtype=self.dtype) for b in layout.blocks[g0 : g0 + basis_batch]]
            ptrs = (C.c_void_p * len(blks))(*[b.ctypes.data if b.size else None for b in blks])
            nat.check(self._lib.calb2_plan_set_basis(self._handle, g0, len(blks), ptrs))
        info = nat.PlanInfo()
        nat.check(self._lib.calb2_plan_get_info(self._handle, C.byref(info)))
        self.info = {name: getattr(info, name) for name, _ in nat.PlanInfo._fields_}

    # -- lifetime
    def close(self):
        if self._handle:
            self._lib.calb2_plan_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- per integration
    def _arr(self, a):
        return np.ascontiguousarray(a, dtype=self.dtype)

    def _ptr(self, a):
        return nat.fptr(a, self.dtype)

    def set_integration(self, data_r, data_i, wgts):
        d_r, d_i, w = self._arr(data_r), self._arr(data_i), self._arr(wgts)
        assert d_r.shape == (self.layout.nbls, self.layout.nfreqs), d_r.shape
        nat.check(self._lib.calb2_set_integration(self._handle, self._ptr(d_r), self._ptr(d_i), self._ptr(w)))

    def set_gains(self, g_r, g_i):
        g_r, g_i = self._arr(g_r), self._arr(g_i)
        assert g_r.shape == (self.layout.nants, self.layout.nfreqs), g_r.shape
        nat.check(self._lib.calb2_set_gains(self._handle, self._ptr(g_r), self._ptr(g_i)))

    def set_coeffs(self, coef_r, coef_i):
        c_r, c_i = self._arr(coef_r), self._arr(coef_i)
        assert c_r.shape == (self.layout.ncoef,), c_r.shape
        nat.check(self._lib.calb2_set_coeffs(self._handle, self._ptr(c_r), self._ptr(c_i)))

    def init_coeffs(self, sky_r, sky_i):
        s_r, s_i = self._arr(sky_r), self._arr(sky_i)
        nat.check(self._lib.calb2_init_coeffs(self._handle, self._ptr(s_r), self._ptr(s_i)))

    def prior_sums(self, sky_r, sky_i):
        s_r, s_i = self._arr(sky_r), self._arr(sky_i)
        pr, pi = C.c_double(), C.c_double()
        nat.check(self._lib.calb2_prior_sums(self._handle, self._ptr(s_r), self._ptr(s_i), C.byref(pr), C.byref(pi)))
        return self.dtype.type(pr.value), self.dtype.type(pi.value)

    def apply_model_snr_weights(self):
        nat.check(self._lib.calb2_apply_model_snr_weights(self._handle))

    # -- the loop
    def fit(self, optimizer="Adamax", maxsteps=10000, tol=1e-14, use_min=False, freeze_model=False,
            model_regularization=None, prior_r_sum=0.0, prior_i_sum=0.0, n_profile_steps=0, steps_per_sync=0,
            use_graph=None, fuse_tail_update=False, **opt_kwargs):
        if optimizer not in REFERENCE_OPTIMIZERS:
            raise KeyError(optimizer)  # calibration.py:571: OPTIMIZERS[optimizer]
        if optimizer not in OPTIMIZER_IDS:
            raise NotImplementedError(f"optimizer {optimizer!r} has no device implementation yet")
        hp = dict(KERAS_DEFAULTS[optimizer])
        unknown = set(opt_kwargs) - set(hp)
        if unknown:
            raise TypeError(f"unexpected optimizer arguments for {optimizer}: {sorted(unknown)}")
        hp.update(opt_kwargs)
        opts = nat.FitOptions(
            optimizer=OPTIMIZER_IDS[optimizer], learning_rate=hp["learning_rate"], beta_1=hp.get("beta_1", 0.0),
            beta_2=hp.get("beta_2", 0.0), epsilon=hp.get("epsilon", 0.0), maxsteps=int(maxsteps), tol=float(tol),
            use_min=int(bool(use_min)), freeze_model=int(bool(freeze_model)),
            regularization=1 if model_regularization == "sum" else 0, prior_r_sum=float(prior_r_sum),
            prior_i_sum=float(prior_i_sum), n_profile_steps=int(n_profile_steps), steps_per_sync=int(steps_per_sync),
            use_graph=0 if use_graph is None else (1 if use_graph else -1), fuse_tail_update=int(bool(fuse_tail_update)),
            rho=hp.get("rho", 0.0), momentum=hp.get("momentum", 0.0),
            initial_accumulator_value=hp.get("initial_accumulator_value", 0.0),
            l1_regularization_strength=hp.get("l1_regularization_strength", 0.0),
            l2_regularization_strength=hp.get("l2_regularization_strength", 0.0),
            learning_rate_power=hp.get("learning_rate_power", 0.0), nesterov=int(bool(hp.get("nesterov", False))),
            weight_decay=hp.get("weight_decay", 0.0),
        )
        if optimizer == "LAMB":
            # the trust ratio is per tf.Variable: the reference keeps one coefficient variable per chunk (calibration.py:560-567)
            bounds = np.ascontiguousarray(self.layout.chunk_coef_bounds(), dtype=np.int64)
            nat.check(self._lib.calb2_plan_set_variables(self._handle, len(bounds) - 1,
                                                         bounds.ctypes.data_as(C.POINTER(C.c_int64))))
        hist = np.zeros(max(1, int(maxsteps)), dtype=self.dtype)
        res = nat.FitResult()
        nat.check(self._lib.calb2_fit(self._handle, C.byref(opts), self._ptr(hist), C.byref(res)))
        result = {name: getattr(res, name) for name, _ in nat.FitResult._fields_}
        return hist[: res.nsteps_recorded].copy(), result

    def loss_and_grads(self, model_regularization=None, prior_r_sum=0.0, prior_i_sum=0.0):
        lay = self.layout
        loss = C.c_double()
        dg_r = np.zeros((lay.nants, lay.nfreqs), dtype=self.dtype)
        dg_i = np.zeros_like(dg_r)
        dc_r = np.zeros(lay.ncoef, dtype=self.dtype)
        dc_i = np.zeros_like(dc_r)
        nat.check(self._lib.calb2_loss_and_grads(
            self._handle, 1 if model_regularization == "sum" else 0, float(prior_r_sum), float(prior_i_sum),
            C.byref(loss), self._ptr(dg_r), self._ptr(dg_i), self._ptr(dc_r), self._ptr(dc_i)))
        return self.dtype.type(loss.value), dg_r, dg_i, dc_r, dc_i

    # -- results
    def get_gains(self):
        g_r = np.zeros((self.layout.nants, self.layout.nfreqs), dtype=self.dtype)
        g_i = np.zeros_like(g_r)
        nat.check(self._lib.calb2_get_gains(self._handle, self._ptr(g_r), self._ptr(g_i)))
        return g_r, g_i

    def get_coeffs(self):
        c_r = np.zeros(self.layout.ncoef, dtype=self.dtype)
        c_i = np.zeros_like(c_r)
        nat.check(self._lib.calb2_get_coeffs(self._handle, self._ptr(c_r), self._ptr(c_i)))
        return c_r, c_i

    def get_model(self):
        m_r = np.zeros((self.layout.nbls, self.layout.nfreqs), dtype=self.dtype)
        m_i = np.zeros_like(m_r)
        nat.check(self._lib.calb2_get_model(self._handle, self._ptr(m_r), self._ptr(m_i)))
        return m_r, m_i

    def get_weights(self):
        w = np.zeros((self.layout.nbls, self.layout.nfreqs), dtype=self.dtype)
        nat.check(self._lib.calb2_get_weights(self._handle, self._ptr(w)))
        return w

    # -- multi-GPU
    def comm_init(self, unique_id, rank, nranks):
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        nat.check(self._lib.calb2_comm_init(self._handle, C.cast(buf, C.c_void_p), rank, nranks, nat.find_nccl().encode()))


    def peer_export(self):
        """64-byte cudaIpc handle of this rank's exchange buffer (to be all-gathered by the caller)."""
        buf = (C.c_char * 64)()
        nat.check(self._lib.calb2_comm_peer_export(self._handle, C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def peer_import(self, handles, rank, nranks):
        """`handles`: the exported handles of all ranks, concatenated in rank order (64 bytes each)."""
        blob = bytes(handles)
        assert len(blob) == 64 * nranks, len(blob)
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        nat.check(self._lib.calb2_comm_peer_import(self._handle, C.cast(buf, C.c_void_p), rank, nranks))
    def peer_close(self):
        """Last publish / wait round of the exchange, then unmap the peers' buffers (every rank must call it)."""
        nat.check(self._lib.calb2_comm_peer_close(self._handle))


def comm_init_peer(plan, rank, world):
    """Set up the NVLink peer-memory exchange with torch.distributed (any backend) as the out-of-band channel.
    Synchronise the ranks (e.g. `dist.barrier()`) before closing the plans: peers read each other's buffers."""
    import torch
    import torch.distributed as dist

    mine = torch.frombuffer(bytearray(plan.peer_export()), dtype=torch.uint8)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = mine.to(dev)
    allh = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine)
    plan.peer_import(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh), rank, world)
    dist.barrier()


def nccl_unique_id():
    buf = (C.c_char * 128)()
    nat.check(nat.load().calb2_comm_unique_id(C.cast(buf, C.c_void_p), nat.find_nccl().encode()))
    return bytes(buf)
