"""eigensolver_b200 - B200-native dispersion-function sweep for the EIGENSOLVER
shooting solvers (slab / cylinder flux tubes with non-uniform density).

The numerical path lives in libeigensolver_b200.so (CUDA, sm_100a) behind the C
ABI of include/eigensolver_b200.h; this package is the thin host side.
"""
from ._lib import EsbError, LIB_PATH, load  # noqa: F401
from .solver import (CYLINDER_CORONAL, CYLINDER_FLOW_CORONAL, CYLINDER_PHOTOSPHERIC, AxialFlowMedium, GaussianAxialFlow, SLAB_CORONAL, SLAB_FLOW_CORONAL,  # noqa: F401
                     SLAB_PHOTOSPHERIC, DiscretisationWarning, DispersionSolver, ModelSpec, EpsteinDensity, FlowMedium, GaussianDensity, GaussianFlow, Medium,
                     PowerLawRotation, RootTable, bessel_ik_scaled, bessel_jy)
from .reference_api import ReferenceScript  # noqa: F401

__version__ = "0.1.0"
