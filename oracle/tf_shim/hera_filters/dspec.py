"""Fake `hera_filters.dspec`: dpss_operator restated over scipy (see calamity_b200.modeling.dpss_basis)."""
import numpy as np

from calamity_b200.modeling import dpss_basis


def dpss_operator(x, filter_centers, filter_half_widths, cache=None, eigenval_cutoff=None, **kwargs):
    x = np.asarray(x)
    cols = []
    for fc, fw, cut in zip(filter_centers, filter_half_widths, eigenval_cutoff):
        vecs = dpss_basis(x, fw, cut)
        xc = x[len(x) // 2]
        cols.append(vecs * np.exp(2j * np.pi * (x[:, None] - xc) * fc))
    amat = np.hstack(cols)
    return amat, [c.shape[1] for c in cols]
