"""See README.md.  `activate()` puts the fake packages and /root/reference on sys.path and patches the NumPy
aliases (np.bool, np.float) the reference still uses."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"


def activate():
    import numpy as np

    for alias, real in (("bool", bool), ("float", float), ("int", int), ("complex", complex)):
        if not hasattr(np, alias):
            setattr(np, alias, real)
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import calamity  # the reference package itself
    from calamity import calibration, modeling, cal_utils, simple_cov  # noqa: F401

    return calamity
