"""Fake `pyuvdata`: the duck-typed stand-ins of calamity_b200.uvstandins under pyuvdata's names."""
from calamity_b200.uvstandins import MiniUVData as UVData  # noqa: F401
from calamity_b200.uvstandins import MiniUVCal as UVCal  # noqa: F401
from calamity_b200.uvstandins import MiniUVFlag as UVFlag  # noqa: F401
from . import utils  # noqa: F401
