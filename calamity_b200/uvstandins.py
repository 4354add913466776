"""Minimal duck-typed stand-ins for pyuvdata's UVData / UVCal / UVFlag.

pyuvdata is not installable in this image, so tests, the benchmark's API-level leg and users without
pyuvdata can drive the drop-in API with these.  They expose exactly the attribute / method surface the fit
path touches with pyuvdata's (old, "future_array_shapes=False") array shapes:
    UVData : data_array/flag_array/nsample_array [Nblts, 1, Nfreqs, Npols], time_array, ant_1/2_array, ...
    UVCal  : gain_array/flag_array [Nants, 1, Nfreqs, Ntimes, Njones], ant_array, jones_array, time_array
    UVFlag : weights_array/flag_array [Nblts, 1, Nfreqs, Npols]
Real pyuvdata objects work with `calamity_b200.calibration` as well: nothing there imports this module
unless pyuvdata is missing.
"""
import copy

import numpy as np

_POL_NUMS = {"xx": -5, "yy": -6, "xy": -7, "yx": -8, "pi": 1, "pq": 2, "pu": 3, "pv": 4,
             "rr": -1, "ll": -2, "rl": -3, "lr": -4}
_POL_STRS = {v: k for k, v in _POL_NUMS.items()}


def polstr2num(pol, x_orientation=None):
    """pyuvdata.utils.polstr2num for the strings the fit path meets ('xx', 'ee', 'Jxx', ...)."""
    if isinstance(pol, (list, tuple, np.ndarray)):
        return [polstr2num(p, x_orientation=x_orientation) for p in pol]
    key = str(pol).lower()
    if key.startswith("j"):
        key = key[1:]
    if set(key) <= {"e", "n"} and len(key) == 2:
        east_is_x = x_orientation is None or str(x_orientation).lower() in ("east", "e")
        table = {"e": "x", "n": "y"} if east_is_x else {"e": "y", "n": "x"}
        key = "".join(table[c] for c in key)
    if key not in _POL_NUMS:
        raise KeyError(f"unknown polarization string {pol!r}")
    return _POL_NUMS[key]


def polnum2str(num, x_orientation=None):
    return _POL_STRS[int(num)]


class MiniUVData:
    def __init__(self, antpos, freqs, times, antpairs, pols=("xx",), x_orientation="east", integration_time=10.0):
        """antpos: {antenna number: ENU xyz}; antpairs: list of (a, b); data zero-initialised."""
        self.antpos = {int(k): np.asarray(v, dtype=float) for k, v in antpos.items()}
        self.freq_array = np.asarray(freqs, dtype=float)[None, :]
        times = np.asarray(times, dtype=float)
        antpairs = [tuple(int(x) for x in ap) for ap in antpairs]
        self.ant_1_array = np.asarray([ap[0] for _ in times for ap in antpairs], dtype=int)
        self.ant_2_array = np.asarray([ap[1] for _ in times for ap in antpairs], dtype=int)
        self.time_array = np.repeat(times, len(antpairs))
        self.lst_array = self.time_array * 2 * np.pi % (2 * np.pi)
        self.polarization_array = np.asarray([polstr2num(p, x_orientation) for p in pols], dtype=int)
        self.x_orientation = x_orientation
        shape = (len(self.time_array), 1, self.freq_array.shape[1], len(pols))
        self.data_array = np.zeros(shape, dtype=np.complex128)
        self.flag_array = np.zeros(shape, dtype=bool)
        self.nsample_array = np.ones(shape, dtype=float)
        self.integration_time = np.full(len(self.time_array), float(integration_time))
        self.telescope_name = "synthetic"
        self.telescope_location = np.zeros(3)
        self.antenna_numbers = np.asarray(sorted(self.antpos), dtype=int)
        self.antenna_names = [f"ant{a}" for a in self.antenna_numbers]
        self.antenna_positions = np.asarray([self.antpos[a] for a in self.antenna_numbers])
        self.spw_array = np.asarray([0])
        self.Nspws = 1
        self.history = ""
        self._refresh()

    # ---- bookkeeping
    def _refresh(self):
        self.Nblts = len(self.time_array)
        self.Nfreqs = self.freq_array.shape[1]
        self.Npols = len(self.polarization_array)
        self.Ntimes = len(np.unique(self.time_array))
        pairs = self.get_antpairs()
        self.Nbls = len(pairs)
        ants = set(self.ant_1_array.tolist()) | set(self.ant_2_array.tolist())
        self.Nants_data = len(ants)
        self.Nants_telescope = len(self.antpos)
        self._pair_rows = {}
        for row, (a, b) in enumerate(zip(self.ant_1_array.tolist(), self.ant_2_array.tolist())):
            self._pair_rows.setdefault((a, b), []).append(row)

    def get_antpairs(self):
        seen, out = set(), []
        for ap in zip(self.ant_1_array.tolist(), self.ant_2_array.tolist()):
            if ap not in seen:
                seen.add(ap)
                out.append(ap)
        return out

    def get_pols(self):
        return [polnum2str(p, self.x_orientation) for p in self.polarization_array]

    def get_antpairpols(self):
        return [ap + (pol,) for ap in self.get_antpairs() for pol in self.get_pols()]

    def get_ENU_antpos(self, pick_data_ants=False):
        nums = sorted(set(self.ant_1_array.tolist()) | set(self.ant_2_array.tolist())) if pick_data_ants else sorted(self.antpos)
        return np.asarray([self.antpos[a] for a in nums]), np.asarray(nums)

    def antpair2ind(self, ant1, ant2=None):
        if ant2 is None:
            ant1, ant2 = ant1
        return np.asarray(self._pair_rows.get((int(ant1), int(ant2)), []), dtype=int)

    def _key2inds(self, key):
        """(rows as stored, rows stored conjugated, (pol index array, conjugate-pol index array))."""
        a, b = int(key[0]), int(key[1])
        fwd = self.antpair2ind(a, b)
        rev = self.antpair2ind(b, a) if a != b else np.asarray([], dtype=int)
        if len(fwd) == 0 and len(rev) == 0:
            raise KeyError(f"antenna pair {(a, b)} not found in data")
        if len(key) > 2:
            pnum = polstr2num(key[2], self.x_orientation)
            pidx = np.where(self.polarization_array == pnum)[0]
            if len(pidx) == 0:
                raise KeyError(f"polarization {key[2]} not found in data")
            pconj = pidx  # xx / yy conjugate to themselves
        else:
            pidx = pconj = np.arange(self.Npols)
        return fwd, rev, (pidx, pconj)

    def get_data(self, key):
        fwd, rev, (pidx, pconj) = self._key2inds(key)
        if len(fwd):
            return self.data_array[fwd, 0][:, :, pidx[0]]
        return np.conj(self.data_array[rev, 0][:, :, pconj[0]])

    def get_flags(self, key):
        fwd, rev, (pidx, pconj) = self._key2inds(key)
        rows, p = (fwd, pidx[0]) if len(fwd) else (rev, pconj[0])
        return self.flag_array[rows, 0][:, :, p]

    # ---- redundancy finder with pyuvdata's interface (groups of baseline numbers, conjugated to u > 0)
    @staticmethod
    def antnums_to_baseline(a, b):
        return 2048 * (int(a) + 1) + (int(b) + 1) + 2 ** 16

    @staticmethod
    def baseline_to_antnums(bl):
        bl = int(bl) - 2 ** 16
        return (bl // 2048 - 1, bl % 2048 - 1)

    def get_redundancies(self, tol=1.0, use_antpos=False, include_conjugates=False, include_autos=True, **kwargs):
        ants = sorted(self.antpos)
        vecs, pairs = [], []
        for n, i in enumerate(ants):
            for j in ants[n if include_autos else n + 1 :]:
                vec = self.antpos[j] - self.antpos[i]
                pair = (i, j)
                flip = vec[0] < -tol or (abs(vec[0]) <= tol and vec[1] < -tol) or (
                    abs(vec[0]) <= tol and abs(vec[1]) <= tol and vec[2] < -tol)
                if include_conjugates and flip:
                    vec, pair = -vec, (j, i)
                vecs.append(vec)
                pairs.append(pair)
        groups, seeds = [], []
        for vec, pair in zip(vecs, pairs):
            for n, c in enumerate(seeds):
                if np.linalg.norm(vec - c) <= tol:
                    groups[n].append(pair)
                    break
            else:
                groups.append([pair])
                seeds.append(np.array(vec))
        centers = [np.mean([self.antpos[j] - self.antpos[i] for (i, j) in grp], axis=0) for grp in groups]
        lengths = [float(np.linalg.norm(c)) for c in centers]
        order = np.argsort(lengths, kind="stable")
        bl_groups = [[self.antnums_to_baseline(*ap) for ap in groups[n]] for n in order]
        return bl_groups, [centers[n] for n in order], [lengths[n] for n in order], []

    @property
    def uvw_array(self):
        return np.asarray([self.antpos[b] - self.antpos[a] for a, b in zip(self.ant_1_array.tolist(), self.ant_2_array.tolist())])

    def select(self, bls=None, times=None, inplace=True):
        obj = self if inplace else copy.deepcopy(self)
        keep = np.ones(obj.Nblts, dtype=bool)
        if bls is not None:
            want = set()
            for ap in bls:
                want.add((int(ap[0]), int(ap[1])))
            have = np.asarray([((a, b) in want) or ((b, a) in want) for a, b in zip(obj.ant_1_array.tolist(), obj.ant_2_array.tolist())])
            keep &= have
        if times is not None:
            keep &= np.any(np.isclose(obj.time_array[:, None], np.asarray(times, dtype=float)[None, :], rtol=0, atol=1e-7), axis=1)
        for name in ("data_array", "flag_array", "nsample_array", "time_array", "lst_array", "ant_1_array", "ant_2_array",
                     "integration_time"):
            setattr(obj, name, getattr(obj, name)[keep])
        obj._refresh()
        return None if inplace else obj

    def __add__(self, other):
        out = copy.deepcopy(self)
        for name in ("data_array", "flag_array", "nsample_array", "time_array", "lst_array", "ant_1_array", "ant_2_array",
                     "integration_time"):
            setattr(out, name, np.concatenate([getattr(self, name), getattr(other, name)], axis=0))
        order = np.argsort(out.time_array, kind="stable")
        for name in ("data_array", "flag_array", "nsample_array", "time_array", "lst_array", "ant_1_array", "ant_2_array",
                     "integration_time"):
            setattr(out, name, getattr(out, name)[order])
        out._refresh()
        return out


class MiniUVCal:
    def __init__(self):
        self.gain_array = None
        self.flag_array = None
        self.quality_array = None
        self.ant_array = None
        self.jones_array = None
        self.time_array = None
        self.x_orientation = None

    def _ant_index(self, ant):
        return int(np.where(np.asarray(self.ant_array) == ant)[0][0])

    def _jones_index(self, jpol):
        num = jpol if isinstance(jpol, (int, np.integer)) else polstr2num(jpol, self.x_orientation)
        return int(np.where(np.asarray(self.jones_array) == num)[0][0])

    def get_gains(self, ant, jpol=None):
        j = 0 if jpol is None else self._jones_index(jpol)
        return self.gain_array[self._ant_index(ant), 0, :, :, j]

    def get_flags(self, ant, jpol=None):
        j = 0 if jpol is None else self._jones_index(jpol)
        return self.flag_array[self._ant_index(ant), 0, :, :, j]

    def select(self, times=None, inplace=True):
        obj = self if inplace else copy.deepcopy(self)
        if times is not None:
            keep = np.any(np.isclose(np.asarray(obj.time_array)[:, None], np.asarray(times, dtype=float)[None, :], rtol=0, atol=1e-7), axis=1)
            obj.time_array = np.asarray(obj.time_array)[keep]
            for name in ("gain_array", "flag_array", "quality_array"):
                if getattr(obj, name) is not None:
                    setattr(obj, name, getattr(obj, name)[:, :, :, keep])
            obj.Ntimes = len(obj.time_array)
        return None if inplace else obj

    def __add__(self, other):
        out = copy.deepcopy(self)
        out.time_array = np.concatenate([np.asarray(self.time_array), np.asarray(other.time_array)])
        order = np.argsort(out.time_array, kind="stable")
        out.time_array = out.time_array[order]
        for name in ("gain_array", "flag_array", "quality_array"):
            if getattr(self, name) is not None:
                setattr(out, name, np.concatenate([getattr(self, name), getattr(other, name)], axis=3)[:, :, :, order])
        out.Ntimes = len(out.time_array)
        return out


class MiniUVFlag:
    """UVFlag(uvdata, mode='flag'): per-baseline flags plus a weights_array the caller fills in."""

    def __init__(self, uvdata, mode="flag"):
        self.mode = mode
        self.flag_array = np.array(uvdata.flag_array, dtype=bool)
        self.weights_array = np.ones(uvdata.data_array.shape, dtype=float)
        self.time_array = np.array(uvdata.time_array)
        self.ant_1_array = np.array(uvdata.ant_1_array)
        self.ant_2_array = np.array(uvdata.ant_2_array)
        self.polarization_array = np.array(uvdata.polarization_array)
        self.x_orientation = uvdata.x_orientation
        self._pair_rows = {}
        for row, (a, b) in enumerate(zip(self.ant_1_array.tolist(), self.ant_2_array.tolist())):
            self._pair_rows.setdefault((a, b), []).append(row)

    def get_antpairs(self):
        return list(self._pair_rows.keys())

    def antpair2ind(self, ant1, ant2=None):
        if ant2 is None:
            ant1, ant2 = ant1
        return np.asarray(self._pair_rows.get((int(ant1), int(ant2)), []), dtype=int)
