"""mppi_b200 -- host-side mirror of the reference controller interface
(controllers/mppi_*.py of SokhengDin/DNN-MPPI-MPC) over libmppi_b200.so, the hand-written
sm_100a MPPI kernels.  The library is loaded lazily by the first controller constructed;
there is no CPU fallback."""
from ._lib import LIB_PATH, MppiError, load  # noqa: F401
from .engine import MPPIEngine  # noqa: F401
