"""Plug the CUDA backend into an importable fluxpy (the reference package `flux`).

``install()`` does what INTEGRATION.md section 1 shows: registers
``CudaTrimeshShapeModel`` in ``flux.shape.trimesh_shape_models`` (the
reference's own plugin list, src/flux/shape.py:424-427) and routes
``flux.form_factors.get_form_factor_matrix`` -- and the per-block assembly in
``flux.compressed_form_factors`` (:560, :685) -- to the fused CUDA path when the
shape model is a CUDA one.  Other shape models keep the reference's row loop.
"""
from . import form_factors as _ff
from .shape import CudaTrimeshShapeModel


def install(flux=None):
    if flux is None:
        import flux
    import flux.shape
    import flux.form_factors
    import flux.config
    if CudaTrimeshShapeModel not in flux.shape.trimesh_shape_models:
        flux.shape.trimesh_shape_models.append(CudaTrimeshShapeModel)
    flux.shape.CudaTrimeshShapeModel = CudaTrimeshShapeModel
    reference = getattr(flux.form_factors, '_reference_get_form_factor_matrix',
                        flux.form_factors.get_form_factor_matrix)

    def get_form_factor_matrix(shape_model, I=None, J=None, eps=None):
        if isinstance(shape_model, CudaTrimeshShapeModel):
            if eps is None:
                eps = flux.config.DEFAULT_EPS          # form_factors.py:16-17
            return _ff.get_form_factor_matrix(shape_model, I, J, eps)
        return reference(shape_model, I, J, eps)

    get_form_factor_matrix.__doc__ = reference.__doc__
    flux.form_factors._reference_get_form_factor_matrix = reference
    flux.form_factors.get_form_factor_matrix = get_form_factor_matrix
    try:
        import flux.compressed_form_factors as cff
        cff.get_form_factor_matrix = get_form_factor_matrix
    except ImportError:         # optional dependencies of that module may be missing
        pass
    return get_form_factor_matrix
