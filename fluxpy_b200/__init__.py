"""fluxpy_b200 -- B200-native (sm_100a) form-factor assembly for fluxpy.

Drop-in for the ``flux.shape`` / ``flux.form_factors`` slice of the reference
(sampotter/fluxpy) that builds the visibility-gated view-factor CSR matrix.
Python here is host glue over the C ABI of ``libfluxb200.so``
(``include/fluxb200.h``); there is no CPU fallback.
"""
from . import config  # noqa: F401
from .form_factors import get_form_factor_matrix  # noqa: F401
from .shape import CudaTrimeshShapeModel, TrimeshShapeModel, trimesh_shape_models  # noqa: F401
from .device_csr import DeviceCsrSlab, get_form_factor_matrix_device  # noqa: F401
