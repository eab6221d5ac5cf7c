"""grb200: B200-native channelize + DMR-demod hot path behind the GNU Radio 3.5 block interface.

Thin Python mirror of the reference's operator API over libgr_cuda (include/gr_cuda.h).  PyTorch is
used by callers only for device memory, streams and torch.distributed; the compute is the hand-written
sm_100a kernels in csrc/.  No CPU fallback exists (grb200.lib.load raises if the library is missing).
"""
from . import firdes, synth  # noqa: F401
from .lib import GrCudaError, ORDER_GENERIC, ORDER_SSE, launches, load  # noqa: F401
