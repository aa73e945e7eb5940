from fcest_b200.models import (SlidingWindows, SparseVariationalWishartProcess, VariationalWishartProcess,
                               WishartProcessLikelihood)

from . import likelihoods, sliding_windows, wishart_process  # noqa: F401

__all__ = ["SlidingWindows", "SparseVariationalWishartProcess", "VariationalWishartProcess",
           "WishartProcessLikelihood"]
