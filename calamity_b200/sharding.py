"""Baseline-group sharding across GPUs.

Every rank owns a contiguous range of fitting groups -- their basis rows, data, weights, coefficients and
coefficient optimizer state -- balanced by basis bytes (sum of ncomp * nslots), not by group count.  Gains
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
    """One rank's view: a RaggedLayout of its groups plus the index ranges into the full flat vectors."""

    def __init__(self, full, g0, g1):
        self.g0, self.g1 = g0, g1
        slot0 = int(np.sum(full.group_nslots[:g0]))
        slot1 = slot0 + int(np.sum(full.group_nslots[g0:g1]))
        self.bl0 = int(np.sum(full.slot_nbls[:slot0]))
        self.bl1 = self.bl0 + int(np.sum(full.slot_nbls[slot0:slot1]))
        self.coef0 = int(full.group_coef0[g0])
        self.coef1 = int(full.group_coef0[g1])
        lay = RaggedLayout(full.nants, full.nfreqs, dtype=full.dtype)
        lay.group_ncomp = full.group_ncomp[g0:g1]
        lay.group_nslots = full.group_nslots[g0:g1]
        lay.slot_nbls = full.slot_nbls[slot0:slot1]
        lay.bl_ant0 = full.bl_ant0[self.bl0 : self.bl1]
        lay.bl_ant1 = full.bl_ant1[self.bl0 : self.bl1]
        lay.blocks = full.blocks[g0:g1]
        lay.chunks = []  # chunk structure is a property of the full problem only
        lay._finalize()
        lay.group_ncomp = np.ascontiguousarray(lay.group_ncomp)
        lay.group_nslots = np.ascontiguousarray(lay.group_nslots)
        lay.slot_nbls = np.ascontiguousarray(lay.slot_nbls)
        lay.bl_ant0 = np.ascontiguousarray(lay.bl_ant0)
        lay.bl_ant1 = np.ascontiguousarray(lay.bl_ant1)
        self.layout = lay

    def take_baselines(self, flat):
        return np.ascontiguousarray(flat[self.bl0 : self.bl1])

    def take_coeffs(self, flat):
        return np.ascontiguousarray(flat[self.coef0 : self.coef1])


def group_costs(layout):
    """Per-group cost in 4-byte words per channel: basis rows (ncomp per slot) plus the per-baseline traffic of the
    iteration -- data_r/data_i/weights read (3), z written (2) and read at both antennas (4), gains gathered (~1):
    about 10 words per baseline.  Balancing on basis bytes alone overloads the rank that owns the short baselines."""
    nbl_per_group = np.add.reduceat(layout.slot_nbls, np.concatenate([[0], np.cumsum(layout.group_nslots)[:-1]]))
    return layout.group_ncomp.astype(np.int64) * layout.group_nslots + 10 * nbl_per_group.astype(np.int64)


def make_shard(full_layout, rank, nranks):
    ranges = partition_groups(group_costs(full_layout), nranks)
    g0, g1 = ranges[rank]
    return Shard(full_layout, g0, g1)
