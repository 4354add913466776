"""Baseline-group sharding across GPUs.

Every rank owns a subset of the fitting groups -- their basis rows, data, weights, coefficients and coefficient
optimizer state -- balanced by cost (basis rows + per-baseline traffic), not by group count: every nranks-th group
(`cyclic`, default) or a contiguous range (`contiguous`).  Gains
and their optimizer state are replicated; per iteration the ranks all-reduce the [2, nants, nfreqs] gain
gradient and three scalars (chi^2 and the two regulariser sums).  Nothing else crosses ranks.
"""
import numpy as np

from .layout import RaggedLayout


def partition_groups(weights, nranks):
    """Contiguous ranges [g0, g1) per rank with near-equal total weight.  Deterministic."""
    w = np.asarray(weights, dtype=np.float64)
    csum = np.concatenate([[0.0], np.cumsum(w)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, nranks):
        target = total * r / nranks
        g = int(np.searchsorted(csum, target, side="left"))
        # pick the boundary closest to the target
        if g > 0 and abs(csum[g - 1] - target) <= abs(csum[min(g, len(w))] - target):
            g -= 1
        bounds.append(min(max(g, bounds[-1]), len(w)))
    bounds.append(len(w))
    return [(bounds[r], bounds[r + 1]) for r in range(nranks)]


class Shard:
    """One rank's view: a RaggedLayout of its groups plus the index arrays into the full flat vectors."""

    def __init__(self, full, groups):
        groups = np.asarray(groups, dtype=np.int64)
        self.groups = groups
        slot0 = np.concatenate([[0], np.cumsum(full.group_nslots)]).astype(np.int64)   # first slot of every group
        bl0 = np.concatenate([[0], np.cumsum(full.slot_nbls)]).astype(np.int64)        # first baseline of every slot
        slots = np.concatenate([np.arange(slot0[g], slot0[g + 1]) for g in groups]) if len(groups) else np.zeros(0, np.int64)
        self.bl_index = (np.concatenate([np.arange(bl0[s], bl0[s + 1]) for s in slots]) if len(slots)
                         else np.zeros(0, np.int64))
        self.coef_index = (np.concatenate([np.arange(full.group_coef0[g], full.group_coef0[g + 1]) for g in groups])
                           if len(groups) else np.zeros(0, np.int64))
        # contiguous shards keep the range view the first version of this class had
        self.contiguous = bool(len(groups) == 0 or np.array_equal(groups, np.arange(groups[0], groups[0] + len(groups))))
        if self.contiguous and len(groups):
            self.g0, self.g1 = int(groups[0]), int(groups[-1]) + 1
            self.bl0, self.bl1 = int(self.bl_index[0]), int(self.bl_index[-1]) + 1
            self.coef0, self.coef1 = int(full.group_coef0[self.g0]), int(full.group_coef0[self.g1])
        lay = RaggedLayout(full.nants, full.nfreqs, dtype=full.dtype)
        lay.group_ncomp = np.ascontiguousarray(full.group_ncomp[groups])
        lay.group_nslots = np.ascontiguousarray(full.group_nslots[groups])
        lay.slot_nbls = np.ascontiguousarray(full.slot_nbls[slots])
        lay.bl_ant0 = np.ascontiguousarray(full.bl_ant0[self.bl_index])
        lay.bl_ant1 = np.ascontiguousarray(full.bl_ant1[self.bl_index])
        lay.blocks = [full.blocks[g] for g in groups]
        lay.chunks = []  # chunk structure is a property of the full problem only
        lay._finalize()
        self.layout = lay

    def take_baselines(self, flat):
        return np.ascontiguousarray(flat[self.bl_index])

    def take_coeffs(self, flat):
        return np.ascontiguousarray(flat[self.coef_index])


def group_costs(layout):
    """Per-group cost in 4-byte words per channel: basis rows (ncomp per slot) plus the per-baseline traffic of the
    iteration -- data_r/data_i/weights read (3), z written (2) and read at both antennas (4), gains gathered (~1):
    about 10 words per baseline.  Balancing on basis bytes alone overloads the rank that owns the short baselines."""
    nbl_per_group = np.add.reduceat(layout.slot_nbls, np.concatenate([[0], np.cumsum(layout.group_nslots)[:-1]]))
    return layout.group_ncomp.astype(np.int64) * layout.group_nslots + 10 * nbl_per_group.astype(np.int64)


def make_shard(full_layout, rank, nranks, mode="cyclic"):
    """`cyclic` (default): rank r owns groups r, r + nranks, ... of the canonical order.  Neighbouring groups cost about
    the same, so the ranks' loads agree to within one group, and -- unlike contiguous ranges, where a few antennas own
    complete rows of baselines on one rank -- every antenna's baselines are spread evenly, which keeps the per-rank
    gain-gradient reduce short and equal (measured at 8 GPUs: 35-57 us -> see profiles/).  `contiguous`: ranges of
    near-equal cost."""
    if nranks == 1:
        return Shard(full_layout, np.arange(full_layout.ngroups))
    if mode == "contiguous":
        g0, g1 = partition_groups(group_costs(full_layout), nranks)[rank]
        return Shard(full_layout, np.arange(g0, g1))
    if mode != "cyclic":
        raise ValueError(f"unknown sharding mode {mode!r}")
    return Shard(full_layout, np.arange(rank, full_layout.ngroups, nranks))
