from fcest_b200.helpers.inference import KerasAdam, run_adam  # noqa: F401
