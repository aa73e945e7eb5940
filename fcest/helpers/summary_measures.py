from fcest_b200.helpers.summary_measures import *  # noqa: F401,F403
