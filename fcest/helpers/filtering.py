from fcest_b200.helpers.filtering import *  # noqa: F401,F403
from fcest_b200.helpers.filtering import _butter_highpass, _compute_lower_frequency_cutoff  # noqa: F401
