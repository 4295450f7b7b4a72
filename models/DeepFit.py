"""Path-compatible shim: the reference scripts do ``sys.path.insert(0, '../models'); import DeepFit``
(tutorial/compute_normals.py:9-12).  Everything lives in ``deepfit_b200.DeepFit``."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from deepfit_b200.DeepFit import *  # noqa: F401,F403,E402
from deepfit_b200.DeepFit import DeepFit, compute_principal_curvatures, fit_Wjet  # noqa: F401,E402
