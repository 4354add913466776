"""tensorflow_addons stand-in for oracle/tf_shim: only `optimizers.LAMB`, the one symbol the reference touches
(calibration.py:15, 26).  The rule is oracle/tf_shim/tensorflow/optimizers.py:LAMB (tfa LAMB._resource_apply_dense)."""
from tensorflow.optimizers import LAMB as _LAMB


class _Optimizers:
    LAMB = _LAMB


optimizers = _Optimizers()
