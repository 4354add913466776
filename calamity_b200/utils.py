"""Small helpers shared by the drop-in modules (mirror of /root/reference/calamity/utils.py:1-37)."""
import numpy as np

try:  # tqdm is optional at run time; progress bars are cosmetic
    import tqdm as _tqdm

    def _bar(iterable, **kw):
        return _tqdm.tqdm(iterable, **kw)

    def _nbbar(iterable, **kw):
        try:
            import tqdm.notebook as tn

            return tn.tqdm(iterable, **kw)
        except Exception:
            return _tqdm.tqdm(iterable, **kw)

except Exception:  # pragma: no cover

    def _bar(iterable, **kw):
        return iterable

    _nbbar = _bar

PBARS = {True: _nbbar, False: _bar}


def echo(message, verbose=True):
    if verbose:
        print(message)


def select_baselines(uvdata, bllen_min=0.0, bllen_max=np.inf, bl_ew_min=0.0, ex_ants=None, select_ants=None):
    """Keep baselines inside a length window / EW-projection cut / antenna lists (utils.py:13-37). In place."""
    banned = set(ex_ants or [])
    pos, nums = uvdata.get_ENU_antpos(pick_data_ants=True)
    where = dict(zip(nums, pos))
    allowed = set(nums) if select_ants is None else set(select_ants)
    keep = []
    for a, b in uvdata.get_antpairs():
        vec = where[a] - where[b]
        ln = np.linalg.norm(vec)
        inside = bllen_min <= ln <= bllen_max and np.abs(vec[0]) > bl_ew_min
        if inside and not ({a, b} & banned) and {a, b} <= allowed:
            keep.append((a, b))
    uvdata.select(bls=keep, inplace=True)
