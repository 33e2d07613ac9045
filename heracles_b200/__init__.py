"""
heracles_b200 -- B200-native (sm_100a) backend for the catalogue -> HEALPix map
-> alm -> Cl hot path of Heracles; installed next to Heracles it is the
``heracles.cuda`` backend the reference's Field layer and CLI can drive
unchanged.

Public surface (mirrors the reference interfaces of this path):
  CudaHealpixMapper            <- heracles.healpy.HealpixMapper
  CudaDiscreteMapper           <- heracles.ducc.DiscreteMapper (pixel-free: catalogue -> alm, exact sums)
  alm2cl, angular_power_spectra <- heracles.twopoint
  transform                    <- heracles.mapping.transform (batched over maps)
  OverlappedTransform          the same transforms, run beside the catalogue mapping (second context + stream)
  dices.region_alms, dices.jackknife_cls <- heracles.dices.jackknife (batched region transforms)
  io.read_vmap                 <- heracles.io.read_vmap (mask maps / mask alm at up to nside 8192)

There is no CPU fallback: importing the kernels' library fails loudly if
``heracles_b200/lib/libheracles_cuda.so`` has not been built, and creating a
context fails without a CUDA device.
"""

from ._lib import Context, HeraclesCudaError, get_context, load  # noqa: F401
from .arrays import DeviceArray, update_metadata  # noqa: F401
from .mapper import CudaHealpixMapper  # noqa: F401
from .discrete import CudaDiscreteMapper  # noqa: F401
from .mapping import transform, transform_maps  # noqa: F401
from .twopoint import alm2cl, alm2lmax, angular_power_spectra  # noqa: F401
from .overlap import OverlappedTransform  # noqa: F401
from . import dices  # noqa: F401,E402
from . import io  # noqa: F401,E402

__all__ = [
    "CudaDiscreteMapper",
    "CudaHealpixMapper",
    "DeviceArray",
    "Context",
    "HeraclesCudaError",
    "OverlappedTransform",
    "alm2cl",
    "alm2lmax",
    "angular_power_spectra",
    "dices",
    "get_context",
    "io",
    "load",
    "update_metadata",
]
