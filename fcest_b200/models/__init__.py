from .likelihoods import WishartProcessLikelihood
from .sliding_windows import SlidingWindows
from .wishart_process import SparseVariationalWishartProcess, VariationalWishartProcess

__all__ = ["WishartProcessLikelihood", "VariationalWishartProcess", "SparseVariationalWishartProcess",
           "SlidingWindows"]
