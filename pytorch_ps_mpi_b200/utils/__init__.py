"""Small helpers the reference keeps beside the optimizer (``/root/reference/ps.py:25-50``)."""
from .misc import _bytes_of, bytes_of, find_param, StepTimer, CudaStepTimer, summarize_timings, dump_chrome_trace
from .clocks import ClockSampler, NvmlClockSampler

__all__ = ["_bytes_of", "bytes_of", "find_param", "StepTimer", "CudaStepTimer", "ClockSampler", "NvmlClockSampler",
           "summarize_timings", "dump_chrome_trace"]
