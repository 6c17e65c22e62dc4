# -*- coding: utf-8 -*-
"""rpsmf_b200 -- B200-native PSMF / rPSMF per-timestep filter.

Python surface of the reference kept for the filter hot path:

* ``robust_PSMF``, ``ProbabilisticSequentialMatrixFactorizer``   (ExperimentImpute/rPSMF.py, PSMF.py)
* ``PSMFIter``, ``PSMFRecursive``, ``rPSMFIter``, ``rPSMFIterMissing``, ``rPSMFRecursive``  (pypsmf/psmf)

All arithmetic runs in hand-written sm_100a CUDA (libpsmf_b200.so, C ABI in include/psmf_b200.h).
"""

from .engine import FilterEngine, count_nan, ingest, missing_segments, transpose_mask  # noqa: F401
from .experiment import prepare_missing, prepare_missing_device, run_impute_experiment  # noqa: F401
from .impute import ProbabilisticSequentialMatrixFactorizer, robust_PSMF  # noqa: F401
from .psmf import PSMFIter, PSMFIterMissing, PSMFRecursive  # noqa: F401
from .rpsmf import rPSMFIter, rPSMFIterMissing, rPSMFRecursive  # noqa: F401



def shard_rows(d, world_size, rank):
    """Rows [begin, end) of C, y_t, m_t owned by `rank` when one series is sharded over `world_size` GPUs.
    Shard boundaries fall on 32-row tile boundaries so that every shard keeps the 16-byte alignment the
    TMA-staged kernel needs."""
    tiles = (d + 31) // 32
    b = min(d, ((tiles * rank) // world_size) * 32)
    e = min(d, ((tiles * (rank + 1)) // world_size) * 32)
    return b, e


def shard_series(n_series, world_size, rank):
    """Series [begin, end) owned by `rank` when a batch of independent series (or restarts) is split over
    `world_size` GPUs: no communication between them (BASELINE.json configs[4])."""
    return (n_series * rank) // world_size, (n_series * (rank + 1)) // world_size


__version__ = "0.2.0"
