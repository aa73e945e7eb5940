"""Drop-in import path: ``import fcest`` resolves to the B200-native implementation (``fcest_b200``),
so user code written against OnnoKampman/FCEst (``from fcest.models.wishart_process import
SparseVariationalWishartProcess`` ...) runs unchanged on the hot path."""
from fcest_b200 import __version__  # noqa: F401

from . import helpers, models  # noqa: F401
