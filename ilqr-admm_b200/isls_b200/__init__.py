"""isls_b200 - B200-native batched mirror of the `isls` hot path (iLQR-ADMM)."""
from .projections import (Bound, ObstacleSets, SetConvexSOC, SetConvexSOCComponents, SetConvexSOCRows, bound,  # noqa: F401
                          project_bound)
from .utils import get_double_integrator_AB                # noqa: F401
from .isls import iSLS, PseudoHuberCost                    # noqa: F401
from .sls import SLS                                       # noqa: F401
from .admm import ADMM                                     # noqa: F401
from . import solver                                       # noqa: F401
