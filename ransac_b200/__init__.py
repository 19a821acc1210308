"""ransac_b200 - B200 (sm_100a) hypothesize-and-verify engine behind the USAC plugin surface.

The package holds only what the hot path needs: csrc/ (CUDA kernels + C ABI -> libusac_gpu.so), the host-side mirror of
the reference's plugin interface (usac/, C++) and a thin Python layer over the C ABI used by tests and bench.py.
"""
from . import capi  # noqa: F401
from .api import GpuContext, UsacGpuError, nccl_unique_id  # noqa: F401
