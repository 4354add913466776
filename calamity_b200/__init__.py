"""calamity_b200: B200-native drop-in for CALAMITY's per-integration gain-and-foreground fit.

Public modules mirror the reference package (aewallwi/calamity): `calibration`, `modeling`, `cal_utils`,
`utils`.  The arithmetic of the fit loop runs in hand-written sm_100a kernels behind a C ABI
(include/calamity_b200.h, calamity_b200/csrc/); there is no CPU fallback.
"""
__version__ = "0.1.0"
