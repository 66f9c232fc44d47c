"""``target:`` names of the reference's sgm/modules/diffusionmodules/discretizer.py; defined in edm_glue.py."""
from .edm_glue import DDPMSubsampledSchedule as LegacyDDPMDiscretization  # noqa: F401
from .edm_glue import KarrasSchedule as EDMDiscretization  # noqa: F401
from .edm_glue import SigmaSchedule as Discretization  # noqa: F401
