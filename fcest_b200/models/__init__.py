from .likelihoods import WishartProcessLikelihood
from .wishart_process import SparseVariationalWishartProcess, VariationalWishartProcess

__all__ = ["WishartProcessLikelihood", "VariationalWishartProcess", "SparseVariationalWishartProcess"]
