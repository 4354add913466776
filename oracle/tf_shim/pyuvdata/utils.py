from calamity_b200.uvstandins import polstr2num, polnum2str  # noqa: F401
