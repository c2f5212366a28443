"""hendrics_b200 -- B200-native (sm_100a) drop-in for the HENDRICS pulsar periodicity-search hot
path: the (f, fdot) trial-grid Z^2_n / epoch-folding search behind ``HENzsearch`` / ``HENefsearch``
(``hendrics/efsearch.py``), including ``--fast``.  See DESIGN.md and INTEGRATION.md.

The compute runs only in hand-written CUDA kernels (``csrc/``) behind the C ABI of
``include/hx_search.h``; there is no CPU fallback."""

from . import _cabi  # noqa: F401
from .search import epoch_folding_search, fold_profiles, z_n_search, z_n_search_unbinned  # noqa: F401

__version__ = "0.1.0"
