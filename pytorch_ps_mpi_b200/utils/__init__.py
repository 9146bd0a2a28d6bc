"""Small helpers the reference keeps beside the optimizer (``/root/reference/ps.py:25-50``)."""
from .misc import _bytes_of, bytes_of, find_param, StepTimer, CudaStepTimer
from .clocks import ClockSampler

__all__ = ["_bytes_of", "bytes_of", "find_param", "StepTimer", "CudaStepTimer", "ClockSampler"]
