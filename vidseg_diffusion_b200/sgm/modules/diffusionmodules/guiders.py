"""``target:`` names of the reference's sgm/modules/diffusionmodules/guiders.py; defined in edm_glue.py."""
from .edm_glue import Guidance as Guider  # noqa: F401
from .edm_glue import IdentityGuider, LinearPredictionGuider, VanillaCFG  # noqa: F401
