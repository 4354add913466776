"""Baseline-group sharding across GPUs.

Every rank owns a subset of the fitting groups -- their basis rows, data, weights, coefficients and coefficient
optimizer state -- balanced by cost (basis rows + per-baseline traffic), not by group count: whole runs of one basis
class (`class`, default), every nranks-th group (`cyclic`) or a contiguous range (`contiguous`).  Gains
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


def class_units(layout, nranks, run=64, min_members=4):
    """Sharding units that keep the shared-basis kernel's CTAs whole: the single-slot groups of one basis class, in runs of
    up to `run` (one CTA of calfit_shared.cuh takes up to 64 groups of one class through all channels, and skips work in
    blocks of 8 groups -- dealing single groups round-robin would leave every rank a ragged, mostly empty CTA per class).
    Classes too small to give every rank a full run are cut into runs of ceil(members / nranks), rounded up to 8.
    Every other group is its own unit.  Returns a list of int64 arrays of group indices (ascending inside a unit)."""
    cls = np.asarray(layout.group_class)
    single = np.asarray(layout.group_nslots) == 1
    units = []
    by_class = {}
    for g in range(layout.ngroups):
        if single[g] and cls[g] >= 0:
            by_class.setdefault(int(cls[g]), []).append(g)
        else:
            units.append(np.asarray([g], dtype=np.int64))
    for members in by_class.values():
        if len(members) < min_members:
            units.extend(np.asarray([g], dtype=np.int64) for g in members)
            continue
        step = min(run, max(8, 8 * int(np.ceil(len(members) / (8.0 * nranks)))))
        nunits = int(np.ceil(len(members) / step))
        for u in range(nunits):  # strided: a unit's baselines come from all over the antenna list
            units.append(np.asarray(members[u::nunits], dtype=np.int64))
    return units


def make_shard(full_layout, rank, nranks, mode="class"):
    """`class` (default): whole runs of 64 groups of one basis class (see class_units) are dealt to the ranks, most expensive
    first, each to the least loaded rank (deterministic LPT); groups without a shared basis are dealt the same way one by
    one.  Every rank then holds whole CTAs of the shared-basis kernel, the loads agree to within one unit, and -- as with
    `cyclic` -- every antenna's baselines end up spread over all ranks, which keeps the per-rank gain-gradient reduce
    short.  `cyclic`: rank r owns groups r, r + nranks, ... of the canonical order (the round-1 default; right for the
    streaming kernel).  `contiguous`: ranges of near-equal cost."""
    if nranks == 1:
        return Shard(full_layout, np.arange(full_layout.ngroups))
    if mode == "contiguous":
        g0, g1 = partition_groups(group_costs(full_layout), nranks)[rank]
        return Shard(full_layout, np.arange(g0, g1))
    if mode == "cyclic":
        return Shard(full_layout, np.arange(rank, full_layout.ngroups, nranks))
    if mode != "class":
        raise ValueError(f"unknown sharding mode {mode!r}")
    costs = group_costs(full_layout)
    units = class_units(full_layout, nranks)
    ucost = np.asarray([int(costs[u].sum()) for u in units], dtype=np.int64)
    order = np.lexsort((np.asarray([int(u[0]) for u in units]), -ucost))  # cost descending, first group ascending
    load = np.zeros(nranks, dtype=np.int64)
    mine = []
    for n in order:
        r = int(np.argmin(load))  # first minimum: deterministic
        load[r] += ucost[n]
        if r == rank:
            mine.append(units[n])
    groups = np.sort(np.concatenate(mine)) if mine else np.zeros(0, np.int64)
    return Shard(full_layout, groups)
