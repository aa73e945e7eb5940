from fcest_b200.models.likelihoods import FactoredWishartProcessLikelihood, WishartProcessLikelihood  # noqa: F401
