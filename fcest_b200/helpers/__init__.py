from .array_operations import (_correlation_from_covariance, are_all_positive_definite,
                               convert_tensor_to_correlation, to_correlation_structure)
from .inference import run_adam

__all__ = ["run_adam", "to_correlation_structure", "_correlation_from_covariance",
           "convert_tensor_to_correlation", "are_all_positive_definite"]
