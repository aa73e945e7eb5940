from fcest_b200.helpers.array_operations import *  # noqa: F401,F403
