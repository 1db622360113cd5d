"""geostatsmodels.jl_b200 -- B200-native interpolation hot path behind the GeoStatsModels.jl interface.

The directory name carries a dot, so it is imported under the alias ``gsm_b200``
(``__graft_entry__.load_package()`` registers it in ``sys.modules``).
"""
from . import _lib
from .api import *  # noqa: F401,F403
from .api import monomial_exponents  # noqa: F401
from ._lib import Context, GsmError, MultiContext, PinnedArray  # noqa: F401
from . import distributed  # noqa: F401,E402
