"""``target:`` names of the reference's sgm/modules/diffusionmodules/denoiser.py; defined in edm_glue.py."""
from .edm_glue import BoundDenoiser, Denoiser, DiscreteDenoiser  # noqa: F401
