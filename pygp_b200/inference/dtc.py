"""
Deterministic training conditional approximation (pygp/inference/dtc.py:20-199).

Same constructor and methods as the reference.  The device state is FITC's
(pygp_b200/csrc/fitc.cu) created through `pgp_dtc_create`, which switches
`_update`, `loglikelihood`, `posterior` and `_full_posterior` to the algebra of
dtc.py:54-199 (scalar noise scaling instead of FITC's per-point correction,
Rux = chol(Kuu + Kux Kux^T / sn2 + su2 I), mean through c^T a / sn2).
"""

from .fitc import FITC

__all__ = ['DTC']


class DTC(FITC):
    """Deterministic training conditional approximation to GP inference."""
    _dtc = True
