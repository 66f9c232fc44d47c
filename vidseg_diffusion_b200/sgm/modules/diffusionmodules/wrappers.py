"""``target:`` names of the reference's sgm/modules/diffusionmodules/wrappers.py; defined in edm_glue.py."""
from .edm_glue import IdentityWrapper, OpenAIWrapper  # noqa: F401

OPENAIUNETWRAPPER = "sgm.modules.diffusionmodules.wrappers.OpenAIWrapper"
