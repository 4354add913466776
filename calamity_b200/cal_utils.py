"""Gain bookkeeping helpers (mirror of /root/reference/calamity/cal_utils.py:7-105).

`apply_gains` is vectorised over baselines and times instead of the reference's triple Python loop
(SURVEY.md section 8f rank 2); results are identical element for element.
"""
import copy

import numpy as np

try:  # real pyuvdata when present, duck-typed stand-ins otherwise
    from pyuvdata import UVCal as _UVCal
    from pyuvdata import utils as _uvutils

    _polstr2num = _uvutils.polstr2num
except Exception:  # pragma: no cover - pyuvdata is absent in this image
    from .uvstandins import MiniUVCal as _UVCal
    from .uvstandins import polstr2num as _polstr2num


def blank_uvcal_from_uvdata(uvdata):
    """Unity-gain, unflagged UVCal with the antennas / times / frequencies / Jones terms of `uvdata`
    (cal_utils.py:7-59; gain_convention is 'divide')."""
    cal = _UVCal()
    for src, dst in (("Nfreqs", "Nfreqs"), ("Npols", "Njones"), ("Ntimes", "Ntimes"), ("Nspws", "Nspws"),
                     ("telescope_name", "telescope_name"), ("telescope_location", "telescope_location"),
                     ("Nants_data", "Nants_data"), ("Nants_telescope", "Nants_telescope"),
                     ("antenna_names", "antenna_names"), ("antenna_numbers", "antenna_numbers"),
                     ("antenna_positions", "antenna_positions"), ("spw_array", "spw_array"),
                     ("freq_array", "freq_array"), ("polarization_array", "jones_array"),
                     ("x_orientation", "x_orientation")):
        setattr(cal, dst, getattr(uvdata, src, None))
    cal.history = ""
    cal.ant_array = np.asarray(list(set(uvdata.ant_1_array).union(set(uvdata.ant_2_array))))
    cal.time_array = np.unique(uvdata.time_array)
    cal.integration_time = np.mean(uvdata.integration_time)
    cal.lst_array = np.unique(uvdata.lst_array)
    cal.gain_convention = "divide"
    shape = (cal.Nants_data, cal.Nspws, cal.Nfreqs, cal.Ntimes, cal.Njones)
    cal.flag_array = np.zeros(shape, dtype=bool)
    cal.quality_array = np.zeros(shape, dtype=np.float64)
    cal.gain_array = np.ones(shape, dtype=np.complex128)
    cal.cal_style = "redundant"
    cal.cal_type = "gain"
    cal.time_range = (cal.time_array.min() - cal.integration_time / 2.0, cal.time_array.max() + cal.integration_time / 2.0)
    cal.channel_width = np.median(np.diff(cal.freq_array))
    return cal


def apply_gains(uvdata, gains, inverse=False):
    """Divide (or, with inverse=True, multiply) visibilities by g_i conj(g_j) and OR the antenna flags in
    (cal_utils.py:62-105).  Returns a calibrated deep copy."""
    out = copy.deepcopy(uvdata)
    ant_index = {int(a): n for n, a in enumerate(np.asarray(gains.ant_array).tolist())}
    a0 = np.asarray([ant_index[int(a)] for a in out.ant_1_array])
    a1 = np.asarray([ant_index[int(a)] for a in out.ant_2_array])
    gtimes = np.asarray(gains.time_array)
    tsel = np.asarray([np.where(np.isclose(gtimes, t, rtol=0.0, atol=1e-7))[0][0] for t in out.time_array])
    for pnum, pol in enumerate(uvdata.get_pols()):
        jnum = np.where(np.asarray(gains.jones_array) == _polstr2num(pol, x_orientation=gains.x_orientation))[0][0]
        g0 = gains.gain_array[a0, 0, :, tsel, jnum]  # [Nblts, Nfreqs]
        g1 = gains.gain_array[a1, 0, :, tsel, jnum]
        factor = g0 * np.conj(g1)
        if inverse:
            out.data_array[:, 0, :, pnum] = out.data_array[:, 0, :, pnum] * factor
        else:
            out.data_array[:, 0, :, pnum] = out.data_array[:, 0, :, pnum] / factor
        out.flag_array[:, 0, :, pnum] |= gains.flag_array[a0, 0, :, tsel, jnum] | gains.flag_array[a1, 0, :, tsel, jnum]
    return out
