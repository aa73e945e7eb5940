from fcest_b200.models.wishart_process import *  # noqa: F401,F403
from fcest_b200.models.wishart_process import SparseVariationalWishartProcess, VariationalWishartProcess  # noqa: F401
