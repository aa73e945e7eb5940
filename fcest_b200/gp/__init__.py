from .kernels import Kernel, Matern52
from .parameter import Parameter

__all__ = ["Kernel", "Matern52", "Parameter"]
