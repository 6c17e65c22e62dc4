# -*- coding: utf-8 -*-
"""rpsmf_b200 -- B200-native PSMF / rPSMF per-timestep filter.

Python surface of the reference kept for the filter hot path:

* ``robust_PSMF``, ``ProbabilisticSequentialMatrixFactorizer``   (ExperimentImpute/rPSMF.py, PSMF.py)
* ``PSMFIter``, ``PSMFRecursive``, ``rPSMFIter``, ``rPSMFIterMissing``, ``rPSMFRecursive``  (pypsmf/psmf)

All arithmetic runs in hand-written sm_100a CUDA (libpsmf_b200.so, C ABI in include/psmf_b200.h).
"""

from .engine import FilterEngine  # noqa: F401
from .impute import ProbabilisticSequentialMatrixFactorizer, robust_PSMF  # noqa: F401
from .psmf import PSMFIter, PSMFIterMissing, PSMFRecursive  # noqa: F401
from .rpsmf import rPSMFIter, rPSMFIterMissing, rPSMFRecursive  # noqa: F401

__version__ = "0.1.0"
