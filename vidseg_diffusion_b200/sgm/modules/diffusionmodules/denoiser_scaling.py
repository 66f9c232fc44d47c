"""``target:`` names of the reference's sgm/modules/diffusionmodules/denoiser_scaling.py; defined in edm_glue.py."""
from .edm_glue import EDMScaling, EpsScaling, Preconditioner, VScaling, VScalingWithEDMcNoise  # noqa: F401
