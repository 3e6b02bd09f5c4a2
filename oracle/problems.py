"""Problem data lives with the product (ilqr-admm_b200/isls_b200/configs.py: numpy arrays only, no solver logic);
this shim lets oracle-side code keep importing `oracle.problems`."""
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ilqr-admm_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
from isls_b200.configs import *          # noqa: F401,F403,E402
from isls_b200.configs import _seed      # noqa: F401,E402
