from fcest_b200.helpers import *  # noqa: F401,F403

from . import array_operations, data, filtering, inference, summary_measures  # noqa: F401
