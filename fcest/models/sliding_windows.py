from fcest_b200.models.sliding_windows import SlidingWindows  # noqa: F401
