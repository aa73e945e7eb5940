from fcest_b200.helpers.data import *  # noqa: F401,F403
