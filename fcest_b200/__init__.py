"""fcest_b200 -- B200-native (sm_100a) hot path of OnnoKampman/FCEst behind FCEst's Python API.

Importing the package does not need a GPU; calling any operator does (there is no CPU fallback).
"""
from . import helpers, models
from .models import (SlidingWindows, SparseVariationalWishartProcess, VariationalWishartProcess,
                     WishartProcessLikelihood)

__version__ = "0.1.0"
__all__ = ["helpers", "models", "VariationalWishartProcess", "SparseVariationalWishartProcess",
           "WishartProcessLikelihood", "SlidingWindows"]
